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
constexpr int PP = 52;           // patch pitch in bytes: 13 words (odd), so that a column of words is bank-conflict free; words 0..11 are read
constexpr int BR = 18;           // blurred patch radius
constexpr int HC = 40;           // columns of the horizontal pass (37 needed, 10 words x 4)
constexpr int HP = 46;           // pitch (u16 units, even) of the TRANSPOSED horizontal-pass array hT[col][row]

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

// Blur: out = (sum_i sum_j k_i k_j I(y+i-3, x+j-3) + 32768) >> 16 with k = [18,34,48,56,48,34,18] (exact integers).
//   horizontal pass for the whole 43 x 40 neighbourhood with IDP.4A on 32-bit words of the patch (4 outputs per step),
//   stored TRANSPOSED as u16 so that the 7 vertical taps of a sample are contiguous;
//   vertical pass ON DEMAND at the 512 rotated sample positions with IDP.2A (u16 x u8), 4 per sample.
__global__ void __launch_bounds__(DW * 32) k_describe(PyrGeom g, SlotMap map, UmaxTable umax, const int8_t* __restrict__ pattern,
                                                      const uint32_t* __restrict__ lkp_all, const int* __restrict__ nlkp,
                                                      svo_keypoint* __restrict__ kps_all, uint8_t* __restrict__ desc_all,
                                                      int* __restrict__ nkp, uint32_t lkp_per_slot, int kp_cap) {
    __shared__ __align__(16) uint8_t s_patch[DW][PS * PP];
    __shared__ __align__(16) uint16_t s_hT[DW][HC * HP];
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
    // this lane's 8 pattern pairs (32 signed bytes) as two 16-byte loads (byte loads would be 32 warp instructions with 32 scattered
    // sectors each); requested here, long before they are used, so that their L2 round trip hides behind the patch work
    uint32_t pw[8];
    {
        const uint4 p0 = __ldg(reinterpret_cast<const uint4*>(pattern + lane * 32)), p1 = __ldg(reinterpret_cast<const uint4*>(pattern + lane * 32) + 1);
        pw[0] = p0.x; pw[1] = p0.y; pw[2] = p0.z; pw[3] = p0.w;
        pw[4] = p1.x; pw[5] = p1.y; pw[6] = p1.z; pw[7] = p1.w;
    }
    const LevelGeom& G = g.lvl[level];
    const uint32_t c = lkp_all[(size_t)slot * lkp_per_slot + G.lkp_off + local];
    const int kx = cand_x(c) + SVO_MINB, ky = cand_y(c) + SVO_MINB;  // level pixel coordinates (:797-798)
    const uint8_t* __restrict__ img = G.base + (size_t)slot * G.plane;

    // ---- 43 x 43 patch -> shared memory (REFLECT_101 at the level's own edges; interior keypoints skip the reflection)
    uint8_t* patch = s_patch[warp];
    {
        const int x0 = kx - PR, y0 = ky - PR;
        const bool interior = x0 >= 0 && y0 >= 0 && x0 + PS <= G.w && y0 + PS <= G.h;  // warp-uniform
        const int c1 = lane + 32;                                                        // second column of this lane (< 43 for lanes < 11)
        if (interior) {
            // word-wise: smem word j of row r = image bytes x0 + 4j .. x0 + 4j + 3, cut out of two aligned global words
            const uint8_t* p = img + (size_t)y0 * G.pitch + (x0 & ~3);
            const unsigned sh = (unsigned)(x0 & 3) * 8u;
            // 473 words over 32 lanes = 15 rounds; unrolled in groups so that a group's loads are all in flight before the first is
            // used (rolled, every round waited for its own L2 round trip: 41 % of the kernel's stall samples sat on this line)
#ifndef DESC_UNROLL
#define DESC_UNROLL 15
#endif
#define SVO_STR_(x) #x
#define SVO_UNROLL_(n) _Pragma(SVO_STR_(unroll n))
            SVO_UNROLL_(DESC_UNROLL)
            for (int k = 0; k < (PS * 11 + 31) / 32; ++k) {
                const int it = lane + 32 * k;
                if (it < PS * 11) {
                    const int r = it / 11, j = it - r * 11;
                    const uint32_t* w = reinterpret_cast<const uint32_t*>(p + (size_t)r * G.pitch) + j;
                    reinterpret_cast<uint32_t*>(patch + r * PP)[j] = __funnelshift_r(__ldg(w), __ldg(w + 1), sh);
                }
            }
        } else {
            const int xa = reflect101(x0 + lane, G.w), xb = reflect101(x0 + min(c1, PS - 1), G.w);
            for (int r = 0; r < PS; ++r) {
                const uint8_t* p = img + (size_t)reflect101(y0 + r, G.h) * G.pitch;
                patch[r * PP + lane] = __ldg(p + xa);
                if (c1 < PS) patch[r * PP + c1] = __ldg(p + xb);
            }
        }
        // columns 43..47 of the patch are never initialised: they only feed horizontal outputs 37..39, which nothing samples
    }
    __syncwarp();

    // ---- K4: intensity centroid.  lane u+15 handles column u (31 columns), rows limited by umax.
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
    m10 = __reduce_add_sync(0xFFFFFFFFu, m10);
    m01 = __reduce_add_sync(0xFFFFFFFFu, m01);
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // ---- K5a: horizontal 7-tap pass, 4 outputs per item: item = (row r, word j), outputs columns 4j .. 4j+3
    uint16_t* hT = s_hT[warp];
    {
        const uint32_t K0 = 18u | (34u << 8) | (48u << 16) | (56u << 24), K1 = 48u | (34u << 8) | (18u << 16);
        // lanes take consecutive ROWS of one word column: the patch reads (pitch 13 words) and the transposed u16 stores
        // (consecutive rows of one output column) are then free of bank conflicts
        for (int it = lane; it < PS * (HC / 4); it += 32) {
            const int j = it / PS, r = it - j * PS;
            const uint32_t* w = reinterpret_cast<const uint32_t*>(patch + r * PP) + j;
            const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
            uint16_t* o = hT + (4 * j) * HP + r;
            o[0] = (uint16_t)(__dp4a(w0, K0, 0u) + __dp4a(w1, K1, 0u));
            o[HP] = (uint16_t)(__dp4a(__funnelshift_r(w0, w1, 8), K0, 0u) + __dp4a(__funnelshift_r(w1, w2, 8), K1, 0u));
            o[2 * HP] = (uint16_t)(__dp4a(__funnelshift_r(w0, w1, 16), K0, 0u) + __dp4a(__funnelshift_r(w1, w2, 16), K1, 0u));
            o[3 * HP] = (uint16_t)(__dp4a(__funnelshift_r(w0, w1, 24), K0, 0u) + __dp4a(__funnelshift_r(w1, w2, 24), K1, 0u));
        }
    }
    __syncwarp();

    // ---- steered BRIEF: lane i computes descriptor byte i (pairs 8i .. 8i+7), blurring its 16 samples on demand
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    const float ang = __fmul_rn(angle, factorPI);
    double sd, cd;
    sincos((double)ang, &sd, &cd);
    const float a = (float)cd, b = (float)sd;
    // vertical taps as packed u8 weights for even / odd first rows
    const uint32_t E0 = 18u | (34u << 8) | (48u << 16) | (56u << 24), E1 = 48u | (34u << 8) | (18u << 16);
    const uint32_t O0 = (18u << 8) | (34u << 16) | (48u << 24), O1 = 56u | (48u << 8) | (34u << 16) | (18u << 24);
    auto sample = [&](float px, float py) -> int {
        const int row = __float2int_rn(__fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a)));
        const int col = __float2int_rn(__fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b)));
        // blurred(row, col) = sum_i k_i h(row + 18 + i, col + 18) over source rows row+18 .. row+24 of column col+18
        const int r0 = row + BR, e = r0 & ~1;
        const uint32_t* w = reinterpret_cast<const uint32_t*>(hT + (col + BR) * HP + e);
        const bool odd = r0 & 1;
        const uint32_t B0 = odd ? O0 : E0, B1 = odd ? O1 : E1;
        uint32_t acc = __dp2a_lo(w[0], B0, 32768u);
        acc = __dp2a_hi(w[1], B0, acc);
        acc = __dp2a_lo(w[2], B1, acc);
        acc = __dp2a_hi(w[3], B1, acc);
        return (int)(acc >> 16);
    };
    int val = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t w = pw[k];  // (x0, y0, x1, y1) as signed bytes
        const int t0 = sample((float)(int8_t)(w & 0xFFu), (float)(int8_t)((w >> 8) & 0xFFu));
        const int t1 = sample((float)(int8_t)((w >> 16) & 0xFFu), (float)(int8_t)(w >> 24));
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
