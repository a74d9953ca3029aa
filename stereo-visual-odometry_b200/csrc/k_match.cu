// k_match.cu -- K6: brute-force Hamming matching of 256-bit descriptors.
//
// Replaces DescriptorMatcher("BruteForce-Hamming")->match(query, train) (tracking.cpp:539-544): one
// result per query, the LOWEST train index attaining the minimum popcount distance.
// One thread holds one query descriptor in 8 registers; the train set streams through shared memory in tiles of 128
// descriptors (broadcast reads, no bank conflicts); distance = 8 XOR + a carry-save adder tree + 4 POPC.  Strict '<' while scanning train indices
// in ascending order gives the first-minimum tie-break.
#include <cstdlib>
#include <vector>

#include "svo_common.cuh"

namespace {

constexpr int MT = 128;  // queries per CTA == train descriptors per tile

__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c) {  // IMAD (FMA pipe), not LEA / IADD3 (ALU pipe)
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__device__ __forceinline__ void match_core(const uint8_t* __restrict__ q, int nq, const uint8_t* __restrict__ t, int nt,
                                           int* __restrict__ idx_out, int* __restrict__ dist_out) {
    __shared__ uint4 tile[MT * 2];
    const int qi = blockIdx.x * MT + threadIdx.x;
    if (blockIdx.x * MT >= nq) return;  // CTA-uniform
    uint4 qa = make_uint4(0, 0, 0, 0), qb = qa;
    if (qi < nq) {
        qa = __ldg(reinterpret_cast<const uint4*>(q) + 2 * qi);
        qb = __ldg(reinterpret_cast<const uint4*>(q) + 2 * qi + 1);
    }
    uint32_t bestkey = 0xFFFFFFFFu;  // train sets are far below 2^22 descriptors (checked by the launchers)
    for (int t0 = 0; t0 < nt; t0 += MT) {
        const int cnt = min(MT, nt - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * 2; i += MT) tile[i] = __ldg(reinterpret_cast<const uint4*>(t) + 2 * t0 + i);
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
            const uint4 a = tile[2 * j], b = tile[2 * j + 1];
            // popcount of 256 bits with 4 POPC instead of 8 (POPC is a quarter-rate pipe): carry-save adders on the XOR
            // words, sum = a^b^c, carry = majority(a,b,c), one LOP3 each
            const uint32_t x0 = qa.x ^ a.x, x1 = qa.y ^ a.y, x2 = qa.z ^ a.z, x3 = qa.w ^ a.w;
            const uint32_t x4 = qb.x ^ b.x, x5 = qb.y ^ b.y, x6 = qb.z ^ b.z, x7 = qb.w ^ b.w;
            const uint32_t s0 = x0 ^ x1 ^ x2, c0 = (x0 & x1) | (x2 & (x0 ^ x1));
            const uint32_t s1 = x3 ^ x4 ^ x5, c1 = (x3 & x4) | (x5 & (x3 ^ x4));
            const uint32_t s2 = x6 ^ x7 ^ s0, c2 = (x6 & x7) | (s0 & (x6 ^ x7));
            const uint32_t s3 = c0 ^ c1 ^ c2, c3 = (c0 & c1) | (c2 & (c0 ^ c1));
            // the kernel is bound by the ALU pipe (XOR / LOP3 / POPC); the weighted sum and the (distance, index) key go to the
            // idle FMA pipe as integer multiply-adds, and one unsigned min replaces compare + two selects.  key = d << 22 | index:
            // the minimum key is the minimum distance with the LOWEST train index on ties == OpenCV's first-minimum rule
            const uint32_t d = mad_u32((uint32_t)__popc(c3), 4u, mad_u32((uint32_t)__popc(s3), 2u, (uint32_t)(__popc(s1) + __popc(s2))));
            bestkey = min(bestkey, mad_u32(d, 1u << 22, (uint32_t)(t0 + j)));
        }
    }
    if (qi < nq) {
        const bool any = bestkey != 0xFFFFFFFFu;
        idx_out[qi] = any ? (int)(bestkey & ((1u << 22) - 1u)) : -1;
        dist_out[qi] = any ? (int)(bestkey >> 22) : -1;
    }
}

__global__ void __launch_bounds__(MT) k_match_one(const uint8_t* q, int nq, const uint8_t* t, int nt, int* idx, int* dist) {
    match_core(q, nq, t, nt, idx, dist);
}

// frontend: problem p = 2 * f + kind for unit frame f of the batch; kind 0 = match(left_{f-1}, right_{f-1}),
// kind 1 = match(left_{f-1}, left_f).  Frame slots form a ring; left image slot = fs, right image slot = ring + fs.
__global__ void __launch_bounds__(MT) k_match_frontend(const uint8_t* __restrict__ desc, const int* __restrict__ nkp, int kp_cap,
                                                       int first, int ring, int skip_first, int* __restrict__ match) {
    const int f = blockIdx.y >> 1, kind = blockIdx.y & 1;
    const int cu = first + f, cur = cu >= ring ? cu - ring : cu, prev = cur == 0 ? ring - 1 : cur - 1;  // first, f < ring
    int* out = match + ((size_t)cur * 4 + 2 * kind) * kp_cap;
    if (f == 0 && skip_first) return;  // no predecessor: StereoInit_f2f only detects (tracking.cpp:78-92)
    const int qs = prev, ts = kind == 0 ? ring + prev : cur;
    match_core(desc + (size_t)qs * kp_cap * 32, nkp[qs], desc + (size_t)ts * kp_cap * 32, nkp[ts], out, out + kp_cap);
}


// ---------------------------------------------------------------------------------------------------------------
// K6 on the int8 tensor-core path.  The Hamming distance IS a contraction: d(q, t) = |q| + |t| - 2 * sum_k q_k t_k over {0, 1}^256.
// Descriptors are expanded once per image to one byte per bit (k_match_expand) in the register order of
// mma.sync.m16n8k32.u8.u8.s32 (SASS: IMMA.16832.U8.U8), so both the query (A) and the train (B) fragments are plain 16-byte loads:
//   entry (descriptor d, t = lane & 3) = 16 words; word 2 ks = bytes of bits 4 t .. 4 t + 3 of descriptor word ks, word 2 ks + 1 = bits
//   16 + 4 t .. 19 + 4 t -- exactly a0 / a2 (b0 / b1) of k-step ks for the thread with that t.
// A warp owns 16 queries (32 registers of A), a CTA 256; trains stream through shared memory in tiles of 64 (rows of 20 words:
// conflict-free LDS.128).  Per 16 x 8 block: 4 LDS.128, 8 IMMA, and the running minimum of key = d << 16 | train index by one
// 3-input min per row (the minimum key is the minimum distance with the LOWEST train index: OpenCV's first-minimum rule).
// (`mma.sync ... b1 .xor.popc` compiles for sm_100a but is emulated with 8 IMMA + ~140 instructions per m16n8k256: slower than POPC.)
// ---------------------------------------------------------------------------------------------------------------
constexpr int XW = 64;        // words per expanded descriptor (4 t-entries x 16 words)
constexpr int TT = 64;        // train descriptors per tile
constexpr int TSTRIDE = 20;   // words per (train, t) row in shared memory (16 + 4 pad)
constexpr int MQ = 256;       // queries per CTA (16 warps x 16)
constexpr int MTHR = 2 * MQ;  // threads per CTA

__device__ __forceinline__ uint32_t expand4(uint32_t nib) { return (nib * 0x00204081u) & 0x01010101u; }  // 4 bits -> 4 bytes of 0 / 1

// one thread per (descriptor, t): 16 expanded words + (t == 0) the descriptor's popcount
__global__ void __launch_bounds__(256) k_match_expand(const uint8_t* __restrict__ desc, const int* __restrict__ nkp, int kp_cap, uint32_t* __restrict__ xdesc,
                                                      int* __restrict__ xpop, const int* __restrict__ slots, int fixed_n) {
    const int slot = slots ? slots[blockIdx.y] : blockIdx.y;
    const int n = nkp ? nkp[slot] : fixed_n;
    const int d = blockIdx.x * 64 + (threadIdx.x >> 2), t = threadIdx.x & 3;
    if (d >= n) return;
    const uint4* src = reinterpret_cast<const uint4*>(desc + ((size_t)slot * kp_cap + d) * 32);
    const uint4 lo = __ldg(src), hi = __ldg(src + 1);
    const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    uint32_t o[16];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        o[2 * ks] = expand4((w[ks] >> (4 * t)) & 15u);
        o[2 * ks + 1] = expand4((w[ks] >> (16 + 4 * t)) & 15u);
    }
    uint4* dst = reinterpret_cast<uint4*>(xdesc + (((size_t)slot * kp_cap + d) * 4 + t) * 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
    if (t == 0) {
        int p = 0;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) p += __popc(w[ks]);
        xpop[(size_t)slot * kp_cap + d] = p;
    }
}

__device__ __forceinline__ void imma16832(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// xq / xt: expanded descriptors of the query / train set; pq / pt: their popcounts; nt < 65536
__device__ __forceinline__ void match_core_mma(const uint32_t* __restrict__ xq, const int* __restrict__ pq, int nq, const uint32_t* __restrict__ xt,
                                               const int* __restrict__ pt, int nt, int* __restrict__ idx_out, int* __restrict__ dist_out) {
    __shared__ __align__(16) uint32_t tile[2][TT * 4 * TSTRIDE];  // double-buffered: the next tile arrives (cp.async) while this one is multiplied
    __shared__ __align__(8) int ntkey[2][TT];
    if ((int)blockIdx.x * MQ >= nq) return;  // CTA-uniform
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int r0 = blockIdx.x * MQ + 16 * warp + g, r1 = r0 + 8;
    // stage 64 expanded trains (16 chunks of 16 bytes each) into rows of 20 words; descriptors beyond nt are zero-filled (src-size 0)
    // and get a huge key
    auto stage = [&](int t0, int buf) {
        for (int c = threadIdx.x; c < TT * 16; c += MTHR) {
            const int j = c >> 4, chunk = c & 15;
            const bool in = t0 + j < nt;
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tile[buf][(j * 4 + (chunk >> 2)) * TSTRIDE + 4 * (chunk & 3)]);
            const uint4* src = reinterpret_cast<const uint4*>(xt + (size_t)(in ? t0 + j : 0) * XW) + chunk;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(in ? 16 : 0) : "memory");
        }
        if (threadIdx.x < TT) ntkey[buf][threadIdx.x] = t0 + (int)threadIdx.x < nt ? ((__ldg(pt + t0 + threadIdx.x) << 16) | (t0 + (int)threadIdx.x)) : 0x3FFF0000;
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (nt > 0) stage(0, 0);
    uint32_t aq[2][16];
    int rowbase[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int r = h ? r1 : r0;
        if (r < nq) {
            const uint4* src = reinterpret_cast<const uint4*>(xq + ((size_t)r * 4 + t) * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint4 v = __ldg(src + i);
                aq[h][4 * i] = v.x;
                aq[h][4 * i + 1] = v.y;
                aq[h][4 * i + 2] = v.z;
                aq[h][4 * i + 3] = v.w;
            }
            rowbase[h] = __ldg(pq + r) << 16;
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) aq[h][i] = 0u;
            rowbase[h] = 0;
        }
    }
    int best0 = 0x7FFFFFFF, best1 = 0x7FFFFFFF;
    int buf = 0;
    for (int t0 = 0; t0 < nt; t0 += TT, buf ^= 1) {
        const bool more = t0 + TT < nt;
        if (more) stage(t0 + TT, buf ^ 1);  // (that buffer was released by the barrier at the end of the previous iteration)
        if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const uint32_t* tb = tile[buf];
        const int* nkb = ntkey[buf];
        // two 16 x 8 blocks at a time: their IMMA chains are independent, so the tensor pipe always has a second instruction to
        // issue while the first one's accumulator is in flight (one chain alone reached 55 % of the IMMA issue rate)
#pragma unroll 1
        for (int blk = 0; blk < TT / 8; blk += 2) {
            const uint4* bpA = reinterpret_cast<const uint4*>(tb + ((8 * blk + g) * 4 + t) * TSTRIDE);
            const uint4* bpB = bpA + 8 * 4 * TSTRIDE / 4;  // the next 8 train columns
            int cA[4] = {0, 0, 0, 0}, cB[4] = {0, 0, 0, 0};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint4 vA = bpA[i], vB = bpB[i];
                imma16832(cA, aq[0][4 * i], aq[1][4 * i], aq[0][4 * i + 1], aq[1][4 * i + 1], vA.x, vA.y);
                imma16832(cB, aq[0][4 * i], aq[1][4 * i], aq[0][4 * i + 1], aq[1][4 * i + 1], vB.x, vB.y);
                imma16832(cA, aq[0][4 * i + 2], aq[1][4 * i + 2], aq[0][4 * i + 3], aq[1][4 * i + 3], vA.z, vA.w);
                imma16832(cB, aq[0][4 * i + 2], aq[1][4 * i + 2], aq[0][4 * i + 3], aq[1][4 * i + 3], vB.z, vB.w);
            }
            // key = (|q| + |t| - 2 dot) << 16 | index, for train columns 2 t, 2 t + 1 of each block
            const int2 nkA = *reinterpret_cast<const int2*>(nkb + 8 * blk + 2 * t), nkB = *reinterpret_cast<const int2*>(nkb + 8 * blk + 8 + 2 * t);
            best0 = __vimin3_s32(best0, rowbase[0] + nkA.x - (cA[0] << 17), rowbase[0] + nkA.y - (cA[1] << 17));
            best1 = __vimin3_s32(best1, rowbase[1] + nkA.x - (cA[2] << 17), rowbase[1] + nkA.y - (cA[3] << 17));
            best0 = __vimin3_s32(best0, rowbase[0] + nkB.x - (cB[0] << 17), rowbase[0] + nkB.y - (cB[1] << 17));
            best1 = __vimin3_s32(best1, rowbase[1] + nkB.x - (cB[2] << 17), rowbase[1] + nkB.y - (cB[3] << 17));
        }
        __syncthreads();  // everybody is done with this buffer before the tile after next is staged into it
    }
    best0 = min(best0, __shfl_xor_sync(0xFFFFFFFFu, best0, 1));
    best0 = min(best0, __shfl_xor_sync(0xFFFFFFFFu, best0, 2));
    best1 = min(best1, __shfl_xor_sync(0xFFFFFFFFu, best1, 1));
    best1 = min(best1, __shfl_xor_sync(0xFFFFFFFFu, best1, 2));
    if (t == 0) {
        if (r0 < nq) {
            idx_out[r0] = nt > 0 ? (best0 & 0xFFFF) : -1;
            dist_out[r0] = nt > 0 ? (best0 >> 16) : -1;
        }
        if (r1 < nq) {
            idx_out[r1] = nt > 0 ? (best1 & 0xFFFF) : -1;
            dist_out[r1] = nt > 0 ? (best1 >> 16) : -1;
        }
    }
}

__global__ void __launch_bounds__(MTHR) k_match_mma_one(const uint32_t* xq, const int* pq, int nq, const uint32_t* xt, const int* pt, int nt, int* idx, int* dist) {
    match_core_mma(xq, pq, nq, xt, pt, nt, idx, dist);
}

__global__ void __launch_bounds__(MTHR) k_match_mma_frontend(const uint32_t* __restrict__ xdesc, const int* __restrict__ xpop, const int* __restrict__ nkp,
                                                            int kp_cap, int first, int ring, int skip_first, int* __restrict__ match) {
    const int f = blockIdx.y >> 1, kind = blockIdx.y & 1;
    const int cu = first + f, cur = cu >= ring ? cu - ring : cu, prev = cur == 0 ? ring - 1 : cur - 1;  // first, f < ring
    int* out = match + ((size_t)cur * 4 + 2 * kind) * kp_cap;
    if (f == 0 && skip_first) return;  // no predecessor: StereoInit_f2f only detects (tracking.cpp:78-92)
    const int qs = prev, ts = kind == 0 ? ring + prev : cur;
    match_core_mma(xdesc + (size_t)qs * kp_cap * XW, xpop + (size_t)qs * kp_cap, nkp[qs], xdesc + (size_t)ts * kp_cap * XW, xpop + (size_t)ts * kp_cap, nkp[ts],
                   out, out + kp_cap);
}

}  // namespace

int launch_match_frontend(svo_ctx* c, int first, int ring, int n_frames, int skip_first) {
    StageTimer t(c, 5);
    if (c->match_tc && c->kp_cap < 65536) {  // tcgen05 form: packed descriptors in, no expansion pass (k_match_tc.cu)
        int rc = launch_match_tc_frontend(c, first, ring, n_frames, skip_first);
        if (rc != SVO_OK) return rc;
    } else if (c->match_mma && c->kp_cap < 65536 && c->d_xdesc) {
        // expand the descriptors of the batch's frames (both sides; the carried frame was expanded by the batch that produced it)
        const SlotMap both{2, ring, first, ring};
        std::vector<int> slots(2 * n_frames);
        for (int b = 0; b < 2 * n_frames; ++b) slots[b] = both.slot(b);
        if ((size_t)2 * n_frames > c->xslots_cap) return c->fail(SVO_ERR_INVALID, "launch_match_frontend: slot list larger than configured");
        SVO_CUDA_TRY(c, cudaMemcpyAsync(c->d_xslots, slots.data(), sizeof(int) * slots.size(), cudaMemcpyHostToDevice, c->stream));
        k_match_expand<<<dim3((c->kp_cap + 63) / 64, 2 * n_frames), 256, 0, c->stream>>>(c->d_desc, c->d_nkp, c->kp_cap, c->d_xdesc, c->d_xpop, c->d_xslots, 0);
        k_match_mma_frontend<<<dim3((c->kp_cap + MQ - 1) / MQ, 2 * n_frames), MTHR, 0, c->stream>>>(c->d_xdesc, c->d_xpop, c->d_nkp, c->kp_cap, first, ring,
                                                                                                    skip_first, c->d_match);
        c->launches += 2;
    } else {
        dim3 grid((c->kp_cap + MT - 1) / MT, 2 * n_frames);
        k_match_frontend<<<grid, MT, 0, c->stream>>>(c->d_desc, c->d_nkp, c->kp_cap, first, ring, skip_first, c->d_match);
        ++c->launches;
    }
    SVO_CUDA_TRY(c, cudaGetLastError());
    return SVO_OK;
}

int svo_ensure_scratch(svo_ctx* c, size_t bytes) {
    if (c->scratch_bytes >= bytes) return SVO_OK;
    if (c->d_scratch) svo_dev_free(c->d_scratch);
    c->d_scratch = nullptr;
    c->scratch_bytes = 0;
    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_scratch, bytes));
    c->scratch_bytes = bytes;
    return SVO_OK;
}

extern "C" int svo_match_hamming(svo_ctx* c, const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* train_idx, int32_t* dist) {
    if (!c) return SVO_ERR_INVALID;
    if (nq < 0 || nt < 0 || (nq > 0 && (!q || !train_idx || !dist)) || (nt > 0 && !t)) return c->fail(SVO_ERR_INVALID, "svo_match_hamming: bad arguments");
    if (nq == 0) return SVO_OK;
    if (nt >= (1 << 22)) return c->fail(SVO_ERR_INVALID, "svo_match_hamming: train sets of 2^22 descriptors or more are not supported");
    cudaSetDevice(c->device);
    const size_t qb = svo_align((size_t)nq * 32, 256), tb = svo_align((size_t)std::max(nt, 1) * 32, 256), ob = svo_align((size_t)nq * 4, 256);
    // the tensor-core path needs the descriptors expanded to one byte per bit (256 B each) and their popcounts
    const bool tc = getenv("SVO_MATCH_POPC") == nullptr && getenv("SVO_MATCH_IMMA") == nullptr && nt < 65536;
    const bool mma = !tc && getenv("SVO_MATCH_POPC") == nullptr && nt > 0 && nt < 65536;
    const size_t xq = mma ? svo_align((size_t)nq * 256, 256) : 0, xt = mma ? svo_align((size_t)nt * 256, 256) : 0;
    const size_t pqb = mma ? svo_align((size_t)nq * 4, 256) : 0, ptb = mma ? svo_align((size_t)nt * 4, 256) : 0;
    int rc = svo_ensure_scratch(c, qb + tb + 2 * ob + xq + xt + pqb + ptb);
    if (rc != SVO_OK) return rc;
    uint8_t* dq = (uint8_t*)c->d_scratch;
    uint8_t* dt = dq + qb;
    int* di = (int*)(dt + tb);
    int* dd = (int*)((uint8_t*)di + ob);
    uint32_t* dxq = (uint32_t*)((uint8_t*)dd + ob);
    uint32_t* dxt = (uint32_t*)((uint8_t*)dxq + xq);
    int* dpq = (int*)((uint8_t*)dxt + xt);
    int* dpt = (int*)((uint8_t*)dpq + pqb);
    SVO_CUDA_TRY(c, cudaMemcpyAsync(dq, q, (size_t)nq * 32, cudaMemcpyHostToDevice, c->stream));
    if (nt > 0) SVO_CUDA_TRY(c, cudaMemcpyAsync(dt, t, (size_t)nt * 32, cudaMemcpyHostToDevice, c->stream));
    {
        StageTimer tm(c, 5);
        if (tc) {
            if ((rc = launch_match_tc_one(c, dq, nq, dt, nt, di, dd)) != SVO_OK) return rc;
        } else if (mma) {
            k_match_expand<<<dim3((nq + 63) / 64, 1), 256, 0, c->stream>>>(dq, nullptr, nq, dxq, dpq, nullptr, nq);
            k_match_expand<<<dim3((nt + 63) / 64, 1), 256, 0, c->stream>>>(dt, nullptr, nt, dxt, dpt, nullptr, nt);
            k_match_mma_one<<<(nq + MQ - 1) / MQ, MTHR, 0, c->stream>>>(dxq, dpq, nq, dxt, dpt, nt, di, dd);
            c->launches += 3;
        } else {
            k_match_one<<<(nq + MT - 1) / MT, MT, 0, c->stream>>>(dq, nq, dt, nt, di, dd);
            ++c->launches;
        }
    }
    SVO_CUDA_TRY(c, cudaGetLastError());
    SVO_CUDA_TRY(c, cudaMemcpyAsync(train_idx, di, (size_t)nq * 4, cudaMemcpyDeviceToHost, c->stream));
    SVO_CUDA_TRY(c, cudaMemcpyAsync(dist, dd, (size_t)nq * 4, cudaMemcpyDeviceToHost, c->stream));
    SVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SVO_OK;
}
