// k_match_tc.cu -- K6 on the 5th-generation tensor cores: brute-force Hamming matching as a 0/1 int8 GEMM (tcgen05.mma kind::i8,
// accumulators in TMEM), the row minimum taken straight out of TMEM.  Replaces DescriptorMatcher("BruteForce-Hamming")->match
// (tracking.cpp:539-544) like k_match.cu; the only GEMM-shaped stage of the front-end:
//     d(q, t) = |q| + |t| - 2 <q, t>   over {0,1}^256,   key = d << 16 | train index,   result = min key per query
// (the minimum key is the minimum distance with the LOWEST train index on ties = OpenCV's first-minimum rule, bit-exact).
// One CTA = 128 queries (the M = 128 rows of the UMMA = the 128 TMEM lanes); the train set streams through in tiles of 256 (N = 256).
// Both operands are expanded from the packed 32-byte descriptors to one byte per bit IN SHARED MEMORY, directly in the canonical
// K-major no-swizzle UMMA layout (8 x 16-byte core matrices: byte (row r, k) at (r / 8) * 2048 + (k / 16) * 128 + (r % 8) * 16 + k % 16),
// so nothing expanded ever touches HBM (the mma.sync form needed an expansion kernel and 256 bytes per descriptor in HBM).  One thread
// issues 8 MMAs (K = 32 each) per train tile and commits them to an mbarrier; all 8 warps then read the 128 x 256 int32 accumulators
// with tcgen05.ld (warp w: lanes 32 (w % 4) .., columns 128 (w / 4) ..) and fold them into the running key.  97 KB of shared memory and
// 256 TMEM columns per CTA: two CTAs per SM overlap each other's expand / MMA / epilogue phases.
// SASS: UTCIMMA x 8, LDTM.x32, UTCBAR (cuobjdump -sass).  tools/ubench/tcgen05_match.cu is the stand-alone form with its own checker.
#include <cstdlib>

#include "svo_common.cuh"

namespace {

constexpr int TM = 128, TN = 256, KB = 256;              // queries per CTA, train descriptors per tile, bytes (= bits) per descriptor
constexpr int A_BYTES = TM * KB, B_BYTES = TN * KB;      // 32 KB + 64 KB
constexpr int SMEM_BYTES = A_BYTES + B_BYTES + 1024;     // + alignment slack
constexpr uint32_t LBO = 128, SBO = 2048;                // K-direction / 8-row-group strides of the core matrices
constexpr uint32_t IDESC = (2u << 4) /* D = S32 */ | (0u << 7) /* A = U8 */ | (0u << 10) /* B = U8 */ | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {  // UMMA shared-memory descriptor, K-major, no swizzle, version 1
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(LBO >> 4) << 16) | ((uint64_t)(SBO >> 4) << 32) | (1ull << 46);
}

// one packed descriptor (8 words) -> 256 bytes of 0 / 1 at row r of an operand tile; returns its popcount
__device__ __forceinline__ int expand_row(uint8_t* tile, int r, const uint32_t* __restrict__ src, bool valid) {
    uint4 a = make_uint4(0, 0, 0, 0), b = a;
    if (valid) {
        a = __ldg(reinterpret_cast<const uint4*>(src));
        b = __ldg(reinterpret_cast<const uint4*>(src) + 1);
    }
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint8_t* row = tile + (r >> 3) * SBO + (r & 7) * 16;
    int pop = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        pop += __popc(w[i]);
        uint32_t o[8];
#pragma unroll
        for (int n = 0; n < 8; ++n) o[n] = (((w[i] >> (4 * n)) & 15u) * 0x00204081u) & 0x01010101u;  // 4 bits -> 4 bytes
        // bits 32 i .. 32 i + 31 = k chunks 2 i and 2 i + 1 (16 bytes each)
        *reinterpret_cast<uint4*>(row + (2 * i) * LBO) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4*>(row + (2 * i + 1) * LBO) = make_uint4(o[4], o[5], o[6], o[7]);
    }
    return pop;
}

// The whole match of one (query set, train set) pair for the CTA's 128 queries.  All 256 threads of the CTA must call it.
__device__ __forceinline__ void match_core_tc(const uint32_t* __restrict__ q, int nq, const uint32_t* __restrict__ t, int nt, int* __restrict__ idx_out,
                                              int* __restrict__ dist_out) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t s_mbar;
    __shared__ uint32_t s_tmem;
    // per train column of the tile: (|t| + 256) << 16 | train index; 0xFFFFFFFF past the end of the train set.  The key of (q, t) is
    // base - (<q, t> << 17) + (|q| << 16) - (256 << 16); the +256 keeps the partial key non-negative (2 <q, t> <= 2 |t|), |q| is added once
    // at the end (the minimum over t does not depend on it)
    __shared__ __align__(16) uint32_t s_tkey[TN];
    __shared__ uint32_t s_best[TM];
    uint8_t* sm = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = sm;
    uint8_t* sB = sm + A_BYTES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * TM;
    if (q0 >= nq) return;  // CTA-uniform, before anything is allocated

    if (warp == 0) {  // TMEM: 256 columns of 128 lanes x 32 bit = the accumulator tile
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&s_tmem)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    int qpop = 0;
    if (tid < TM) qpop = expand_row(sA, tid, q + (size_t)(q0 + tid) * 8, q0 + tid < nq);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s_tmem;
    // every thread keeps the popcount of "its" query row (lane quarter of its warp)
    const int myrow = 32 * (warp & 3) + lane;
    if (tid < TM) s_best[tid] = (uint32_t)qpop;
    __syncthreads();
    const int my_qpop = (int)s_best[myrow];
    __syncthreads();
    uint32_t best = 0xFFFFFFFFu;
    uint32_t parity = 0;
    for (int t0 = 0; t0 < nt; t0 += TN) {
        {
            const bool tv = t0 + tid < nt;
            const int tp = expand_row(sB, tid, t + (size_t)(t0 + tid) * 8, tv);
            s_tkey[tid] = tv ? ((uint32_t)(tp + 256) << 16) | (uint32_t)(t0 + tid) : 0xFFFFFFFFu;
        }
        asm volatile("fence.proxy.async.shared::cta;");  // the expanded operands (generic-proxy stores) become visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;");
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
#pragma unroll
            for (int k = 0; k < KB / 32; ++k) {  // K = 32 bytes per MMA = two core matrices along K
                const uint64_t da = make_desc(a0 + k * 2 * LBO), db = make_desc(b0 + k * 2 * LBO);
                const uint32_t acc = k > 0;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                    ::"r"(tmem), "l"(da), "l"(db), "r"(IDESC), "r"(acc), "r"(0u));
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_mbar)));
        }
        {  // wait for the accumulators; the spin is bounded so that a fault becomes a CUDA error instead of a hung device
            uint32_t done = 0;
            for (int spin = 0; spin < (1 << 26) && !done; ++spin)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(&s_mbar)), "r"(parity));
            if (!done) __trap();
            parity ^= 1;
        }
        asm volatile("tcgen05.fence::after_thread_sync;");
        // epilogue: this warp's 32 rows (lanes 32 (warp % 4) ..), columns 128 (warp / 4) .. + 127, 32 at a time
        const uint32_t tbase = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(128 * (warp >> 2));
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t v[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
                "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                  "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                  "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                  "=r"(v[31])
                : "r"(tbase + (uint32_t)c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;");
            const uint4* kb = reinterpret_cast<const uint4*>(s_tkey + 128 * (warp >> 2) + c0);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {  // per element: one IMAD (key = base - <q, t> * 2^17) and half a three-way minimum
                const uint4 b4 = kb[j >> 2];
                const uint32_t k0 = b4.x - (v[j] << 17), k1 = b4.y - (v[j + 1] << 17), k2 = b4.z - (v[j + 2] << 17), k3 = b4.w - (v[j + 3] << 17);
                best = __vimin3_u32(best, k0, k1);
                best = __vimin3_u32(best, k2, k3);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();  // the B tile and the accumulators may be overwritten
    }
    // the two column halves of a row sit in warps w and w + 4
    if (warp >= 4) s_best[myrow] = best;
    __syncthreads();
    if (warp < 4) {
        best = min(best, s_best[myrow]);
        const int qi = q0 + myrow;
        if (qi < nq) {
            const bool any = best != 0xFFFFFFFFu;
            idx_out[qi] = any ? (int)(best & 0xFFFFu) : -1;
            dist_out[qi] = any ? (int)(best >> 16) - 256 + my_qpop : -1;
        }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}


__global__ void __launch_bounds__(256, 2) k_match_tc_one(const uint32_t* __restrict__ q, int nq, const uint32_t* __restrict__ t, int nt,
                                                        int* __restrict__ idx, int* __restrict__ dist) {
    match_core_tc(q, nq, t, nt, idx, dist);
}

// front-end: problem blockIdx.y = (unit frame f, kind): kind 0 = match(left_{t-1}, right_{t-1}), kind 1 = match(left_{t-1}, left_t)
__global__ void __launch_bounds__(256, 2) k_match_tc_frontend(const uint8_t* __restrict__ desc, const int* __restrict__ nkp, int kp_cap, int first, int ring,
                                                             int skip_first, int* __restrict__ match) {
    const int f = blockIdx.y >> 1, kind = blockIdx.y & 1;
    const int cu = first + f, cur = cu >= ring ? cu - ring : cu, prev = cur == 0 ? ring - 1 : cur - 1;  // first, f < ring
    int* out = match + ((size_t)cur * 4 + 2 * kind) * kp_cap;
    if (f == 0 && skip_first) return;  // no predecessor: StereoInit_f2f only detects (tracking.cpp:78-92)
    const int qs = prev, ts = kind == 0 ? ring + prev : cur;
    match_core_tc(reinterpret_cast<const uint32_t*>(desc + (size_t)qs * kp_cap * 32), nkp[qs], reinterpret_cast<const uint32_t*>(desc + (size_t)ts * kp_cap * 32),
                  nkp[ts], out, out + kp_cap);
}

bool g_attr_set = false;
int ensure_attr(svo_ctx* c) {
    if (g_attr_set) return SVO_OK;
    SVO_CUDA_TRY(c, cudaFuncSetAttribute(k_match_tc_one, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    SVO_CUDA_TRY(c, cudaFuncSetAttribute(k_match_tc_frontend, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    g_attr_set = true;
    return SVO_OK;
}

}  // namespace

int launch_match_tc_frontend(svo_ctx* c, int first, int ring, int n_frames, int skip_first) {
    int rc = ensure_attr(c);
    if (rc != SVO_OK) return rc;
    k_match_tc_frontend<<<dim3((c->kp_cap + TM - 1) / TM, 2 * n_frames), 256, SMEM_BYTES, c->stream>>>(c->d_desc, c->d_nkp, c->kp_cap, first, ring, skip_first,
                                                                                                     c->d_match);
    ++c->launches;
    return SVO_OK;
}

int launch_match_tc_one(svo_ctx* c, const uint8_t* dq, int nq, const uint8_t* dt, int nt, int* di, int* dd) {
    int rc = ensure_attr(c);
    if (rc != SVO_OK) return rc;
    k_match_tc_one<<<(nq + TM - 1) / TM, 256, SMEM_BYTES, c->stream>>>(reinterpret_cast<const uint32_t*>(dq), nq, reinterpret_cast<const uint32_t*>(dt), nt, di, dd);
    ++c->launches;
    return SVO_OK;
}
