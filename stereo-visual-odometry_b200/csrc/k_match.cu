// k_match.cu -- K6: brute-force Hamming matching of 256-bit descriptors.
//
// Replaces DescriptorMatcher("BruteForce-Hamming")->match(query, train) (tracking.cpp:539-544): one
// result per query, the LOWEST train index attaining the minimum popcount distance.
// One thread holds one query descriptor in 8 registers; the train set streams through shared memory in tiles of 128
// descriptors (broadcast reads, no bank conflicts); distance = 8 XOR + a carry-save adder tree + 4 POPC.  Strict '<' while scanning train indices
// in ascending order gives the first-minimum tie-break.
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

}  // namespace

int launch_match_frontend(svo_ctx* c, int first, int ring, int n_frames, int skip_first) {
    StageTimer t(c, 5);
    dim3 grid((c->kp_cap + MT - 1) / MT, 2 * n_frames);
    k_match_frontend<<<grid, MT, 0, c->stream>>>(c->d_desc, c->d_nkp, c->kp_cap, first, ring, skip_first, c->d_match);
    ++c->launches;
    SVO_CUDA_TRY(c, cudaGetLastError());
    return SVO_OK;
}

int svo_ensure_scratch(svo_ctx* c, size_t bytes) {
    if (c->scratch_bytes >= bytes) return SVO_OK;
    if (c->d_scratch) cudaFree(c->d_scratch);
    c->d_scratch = nullptr;
    c->scratch_bytes = 0;
    SVO_CUDA_TRY(c, cudaMalloc(&c->d_scratch, bytes));
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
    int rc = svo_ensure_scratch(c, qb + tb + 2 * ob);
    if (rc != SVO_OK) return rc;
    uint8_t* dq = (uint8_t*)c->d_scratch;
    uint8_t* dt = dq + qb;
    int* di = (int*)(dt + tb);
    int* dd = (int*)((uint8_t*)di + ob);
    SVO_CUDA_TRY(c, cudaMemcpyAsync(dq, q, (size_t)nq * 32, cudaMemcpyHostToDevice, c->stream));
    if (nt > 0) SVO_CUDA_TRY(c, cudaMemcpyAsync(dt, t, (size_t)nt * 32, cudaMemcpyHostToDevice, c->stream));
    {
        StageTimer tm(c, 5);
        k_match_one<<<(nq + MT - 1) / MT, MT, 0, c->stream>>>(dq, nq, dt, nt, di, dd);
        ++c->launches;
    }
    SVO_CUDA_TRY(c, cudaGetLastError());
    SVO_CUDA_TRY(c, cudaMemcpyAsync(train_idx, di, (size_t)nq * 4, cudaMemcpyDeviceToHost, c->stream));
    SVO_CUDA_TRY(c, cudaMemcpyAsync(dist, dd, (size_t)nq * 4, cudaMemcpyDeviceToHost, c->stream));
    SVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SVO_OK;
}
