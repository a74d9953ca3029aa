// k_describe.cu -- K4 + K5: orientation and steered BRIEF for every selected keypoint, one warp per keypoint.
//
//   K4  IC_Angle (ORBextractor.cpp:21-48): integer moments m10, m01 over the 31x31 circular patch (row extents umax) of
//       the UNBLURRED level, reduced with warp shuffles, then cv::fastAtan2 in plain f32 (no FMA contraction).
//   K5  GaussianBlur(7x7, sigma 2, REFLECT_101) of the level (ORBextractor.cpp:1034-1035) evaluated only where the
//       descriptor samples it: the 43x43 source patch around the keypoint is staged in shared memory (reflecting at the
//       level's own edges), blurred separably in exact integer arithmetic (k = [18,34,48,56,48,34,18], (sum+32768)>>16),
//       and computeOrbDescriptor (ORBextractor.cpp:51-97) samples the 37x37 blurred patch with the 256 rotated pairs.
//   Epilogue (ORBextractor.cpp:797-800, 1044-1053): pt += minBorder, octave, size, pt *= scale, level-major output.
//
// cos/sin: the reference calls the float overloads (cosf/sinf, :56).  Here they are the correctly rounded floats of the
// double-precision result, which equals glibc's cosf/sinf except for rare 1-ulp cases; DESIGN.md discusses the bound.
#include "svo_common.cuh"

namespace {

constexpr int DW = 4;            // warps (keypoints) per CTA
constexpr int PR = 21;           // source patch radius: 18 (max rotated tap) + 3 (blur)
constexpr int PS = 2 * PR + 1;   // 43
constexpr int PP = 44;           // patch pitch
constexpr int BR = 18;           // blurred patch radius
constexpr int BS = 2 * BR + 1;   // 37
constexpr int BP = 40;           // blurred patch pitch

__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

__device__ __forceinline__ float fast_atan2_deg(float y, float x) {  // cv::fastAtan2
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale, p5 = 0.1555786518463281f * scale,
                p7 = -0.04432655554792128f * scale;
    const float ax = fabsf(x), ay = fabsf(y);
    const float eps = (float)2.2204460492503131e-16;
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0.f) a = __fsub_rn(180.f, a);
    if (y < 0.f) a = __fsub_rn(360.f, a);
    return a;
}

struct UmaxTable {
    int v[16];
};

__global__ void __launch_bounds__(DW * 32) k_describe(PyrGeom g, SlotMap map, UmaxTable umax, const int8_t* __restrict__ pattern,
                                                      const uint32_t* __restrict__ lkp_all, const int* __restrict__ nlkp,
                                                      svo_keypoint* __restrict__ kps_all, uint8_t* __restrict__ desc_all,
                                                      int* __restrict__ nkp, uint32_t lkp_per_slot, int kp_cap) {
    __shared__ uint8_t s_patch[DW][PS * PP];
    __shared__ uint16_t s_h[DW][PS * BP];   // horizontal pass: 43 rows x 37 cols
    __shared__ uint8_t s_blur[DW][BS * BP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = map.slot(blockIdx.y);
    const int kidx = blockIdx.x * DW + warp;

    // level of this keypoint: level-major concatenation
    int level = -1, local = 0, total = 0;
    {
        int acc = 0;
        for (int l = 0; l < g.nlevels; ++l) {
            const int nl = nlkp[slot * SVO_MAX_LEVELS + l];
            if (level < 0 && kidx < acc + nl) {
                level = l;
                local = kidx - acc;
            }
            acc += nl;
        }
        total = acc;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) nkp[slot] = total;
    if (level < 0) return;  // warp-uniform
    const LevelGeom& G = g.lvl[level];
    const uint32_t c = lkp_all[(size_t)slot * lkp_per_slot + G.lkp_off + local];
    const int kx = cand_x(c) + SVO_MINB, ky = cand_y(c) + SVO_MINB;  // level pixel coordinates (:797-798)
    const uint8_t* __restrict__ img = G.base + (size_t)slot * G.plane;

    uint8_t* patch = s_patch[warp];
    for (int i = lane; i < PS * PS; i += 32) {
        const int py = i / PS, px = i - py * PS;
        patch[py * PP + px] = __ldg(img + (size_t)reflect101(ky - PR + py, G.h) * G.pitch + reflect101(kx - PR + px, G.w));
    }
    __syncwarp();

    // ---- K4: intensity centroid.  lane u-15 handles column u (31 columns), rows limited by umax.
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int u = lane - 15;
        const int au = u < 0 ? -u : u;
        const uint8_t* ctr = patch + PR * PP + PR + u;
        m10 = u * ctr[0];
        int vs = 0;
#pragma unroll
        for (int v = 1; v <= 15; ++v) {
            if (au <= umax.v[v]) {
                const int plus = ctr[v * PP], minus = ctr[-v * PP];
                vs += v * (plus - minus);
                m10 += u * (plus + minus);
            }
        }
        m01 = vs;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xFFFFFFFFu, m10, o);
        m01 += __shfl_xor_sync(0xFFFFFFFFu, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // ---- K5: separable exact-integer blur of the 37x37 neighbourhood
    uint16_t* hb = s_h[warp];
    for (int i = lane; i < PS * BS; i += 32) {
        const int r = i / BS, x = i - r * BS;
        const uint8_t* p = patch + r * PP + x;  // taps x .. x+6 of the 43-wide row
        hb[r * BP + x] = (uint16_t)(18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3]);
    }
    __syncwarp();
    uint8_t* bl = s_blur[warp];
    for (int i = lane; i < BS * BS; i += 32) {
        const int y = i / BS, x = i - y * BS;
        const uint16_t* p = hb + y * BP + x;
        const uint32_t acc = 18u * (p[0] + p[6 * BP]) + 34u * (p[BP] + p[5 * BP]) + 48u * (p[2 * BP] + p[4 * BP]) + 56u * p[3 * BP];
        bl[y * BP + x] = (uint8_t)((acc + 32768u) >> 16);
    }
    __syncwarp();

    // ---- steered BRIEF: lane i computes descriptor byte i (pairs 8i .. 8i+7)
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    const float ang = __fmul_rn(angle, factorPI);
    const float a = (float)cos((double)ang), b = (float)sin((double)ang);
    const int8_t* pat = pattern + lane * 32;
    const uint8_t* ctr = bl + BR * BP + BR;
    int val = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float x0 = (float)pat[4 * k], y0 = (float)pat[4 * k + 1], x1 = (float)pat[4 * k + 2], y1 = (float)pat[4 * k + 3];
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int t0 = ctr[r0 * BP + c0], t1 = ctr[r1 * BP + c1];
        val |= (t0 < t1) << k;
    }
    desc_all[((size_t)slot * kp_cap + kidx) * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
        svo_keypoint kp;
        kp.x = (float)kx;
        kp.y = (float)ky;
        if (level != 0) {  // pt *= scale (:1044-1051)
            kp.x = __fmul_rn(kp.x, G.scale);
            kp.y = __fmul_rn(kp.y, G.scale);
        }
        kp.size = (float)G.patch_size;
        kp.angle = angle;
        kp.response = (float)cand_r(c);
        kp.octave = level;
        kp.class_id = -1;
        kps_all[(size_t)slot * kp_cap + kidx] = kp;
    }
}

}  // namespace

int launch_describe(svo_ctx* c, const SlotMap& map, int n_img) {
    StageTimer t(c, 4);
    UmaxTable um;
    for (int i = 0; i < 16; ++i) um.v[i] = c->umax[i];
    dim3 grid((c->kp_cap + DW - 1) / DW, n_img);
    k_describe<<<grid, DW * 32, 0, c->stream>>>(c->geom, map, um, c->d_pattern, c->d_lkp, c->d_nlkp, c->d_kps, c->d_desc, c->d_nkp,
                                                c->lkp_per_slot, c->kp_cap);
    ++c->launches;
    SVO_CUDA_TRY(c, cudaGetLastError());
    return SVO_OK;
}
