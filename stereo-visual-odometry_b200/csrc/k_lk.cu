// k_lk.cu -- K7 + K8: pyramidal Lucas-Kanade tracking == cv::calcOpticalFlowPyrLK as called at tracking.cpp:593-618
// (winSize 21x21, maxLevel 3, TermCriteria(COUNT+EPS, 30, 0.01), flags 0, minEigThreshold 1e-3); SURVEY.md A.9.
//
//  K7  k_pyr_down   cv::pyrDown: 5x5 [1 4 6 4 1]^2 / 256, BORDER_REFLECT_101, round half up; levels 1..3 of each image.
//                   Built once per image (the reference rebuilds both pyramids inside each of its 4 calls).
//  K7b k_lk_pad     REFLECT_101 border (32 columns / 24 rows) around every LK plane, so that the tracker never branches on
//                   the image border when it loads pixels.
//  K8  k_lk_track   one warp per point, all pyramid levels in one launch.  Image neighbourhoods are staged per warp in
//                   shared memory as (row, row shifted by one byte) word pairs, so the two horizontal taps of the bilinear
//                   stencil are one aligned 16-bit load feeding IDP.2A.  Scharr gradients are derived on the fly (constant 0
//                   outside the image, exactly like OpenCV's padded derivative buffer -- no derivative image is ever written
//                   to HBM); the interpolated 21x21 template (I, Ix, Iy) lives in shared memory as pixel PAIRS (the two strips of a
//                   lane) and the Gauss-Newton iterations spread its 441 pixels over all 32 lanes, multiplying both gradients of a
//                   pair with the pair's differences in IDP.2A.  All window sums are exact integers (int64 across the warp),
//                   converted to float once.
//
// f32 arithmetic uses explicit round-to-nearest intrinsics (no FMA contraction) to follow OpenCV's scalar expressions.
#include <algorithm>
#include <cfloat>
#include <climits>

#include "svo_common.cuh"

namespace {

constexpr int WIN = SVO_LK_WIN;  // 21
constexpr int LW = 1;            // warps (points) per CTA

__device__ __forceinline__ int reflect101(int i, int n) {
    // |offset| <= WIN beyond the image and n > WIN on every level, so one reflection suffices
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

// horizontal [1 4 6 4 1] pass of one source row for 4 adjacent destination pixels x0..x0+3 (x0 % 4 == 0):
// taps of pixel i start at source byte 2*x0 - 2 + 2*i.  Interior: 4 aligned words + IDP.4A; border: per byte with REFLECT_101.
__device__ __forceinline__ void pyr_hpass4(const uint8_t* __restrict__ row, int x0, int sw, bool interior, int (&h)[4]) {
    if (interior) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(row + 2 * x0 - 4);
        const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3);
        const uint32_t K = 1u | (4u << 8) | (6u << 16) | (4u << 24);
        h[0] = (int)__dp4a(__funnelshift_r(w0, w1, 16), K, (w1 >> 16) & 0xFFu);
        h[1] = (int)__dp4a(w1, K, w2 & 0xFFu);
        h[2] = (int)__dp4a(__funnelshift_r(w1, w2, 16), K, (w2 >> 16) & 0xFFu);
        h[3] = (int)__dp4a(w2, K, w3 & 0xFFu);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = 2 * (x0 + i);
            h[i] = __ldg(row + reflect101(c - 2, sw)) + 4 * __ldg(row + reflect101(c - 1, sw)) + 6 * __ldg(row + reflect101(c, sw)) +
                   4 * __ldg(row + reflect101(c + 1, sw)) + __ldg(row + reflect101(c + 2, sw));
        }
    }
}

// cv::pyrDown: each thread produces 4 x 2 destination pixels (7 source rows, the middle 3 shared by both output rows).  The source is a
// PADDED plane (REFLECT_101 border materialised by k_lk_pad, which is also pyrDown's border mode), so every tap is a plain aligned
// load: no thread takes a border path (with an unpadded source the two border threads of every row made their warps run a ~200
// instruction byte-wise path: 16.7 of 32 threads active per instruction).
__global__ void __launch_bounds__(256) k_pyr_down(LkLevelGeom S, LkLevelGeom D, SlotMap map) {
    const int slot = map.slot(blockIdx.z);
    const int x0 = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4, y0 = (blockIdx.y * 4 + (threadIdx.x >> 6)) * 2;
    if (x0 >= D.w || y0 >= D.h) return;
    const uint8_t* __restrict__ src = S.base + (size_t)slot * S.plane;
    int hr[7][4];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const int sy = min(2 * y0 + i - 2, S.h + 1);  // rows beyond the second output row's taps are unused; -2 .. h + 1 lie in the border
        pyr_hpass4(src + (ptrdiff_t)sy * S.pitch, x0, S.w, true, hr[i]);
    }
    uint8_t* dst = D.base + (size_t)slot * D.plane + (size_t)y0 * D.pitch + x0;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        if (y0 + r >= D.h) break;
        uint32_t out = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int acc = hr[2 * r][i] + hr[2 * r + 4][i] + 4 * (hr[2 * r + 1][i] + hr[2 * r + 3][i]) + 6 * hr[2 * r + 2][i];
            out |= (uint32_t)((acc + 128) >> 8) << (8 * i);
        }
        *reinterpret_cast<uint32_t*>(dst + (size_t)r * D.pitch) = out;  // pitch % 64 == 0: the store stays inside the row
    }
}

struct LkParams {
    LkLevelGeom lv[SVO_LK_LEVELS];  // padded planes (REFLECT_101 border of SVO_LK_PADX / SVO_LK_PADY materialised)
    LkLevelGeom dv[SVO_LK_LEVELS];  // Scharr planes (Ix | Iy << 16 per pixel, 4 bytes; ZERO border: OpenCV's derivI), DERIV form only
    int levels;      // pyramid levels available (maxLevel + 1)
    int first, ring, skip_first;
    LkCall calls[4];  // the chain: the output list of call k is the source list of call k + 1
    int ncalls;
    int standalone;  // 1: src points are a plain xy list (list index call.src_list), n given by n_fixed
    int n_fixed;
    int max_iter;
    double eps2;
    float eps2_lo, eps2_hi;  // eps2 * (1 -/+ 1e-6): outside this band the float value of |delta|^2 decides like the double one
    float min_eig;
    int want_err;
    int pt_cap;
};

// K7b: REFLECT_101 border of the LK planes (what buildOpticalFlowPyramid's copyMakeBorder produces).  Level 0 is copied
// whole from the (unpadded) ORB level-0 plane; levels 1..3 are filled in place from their own interior.
struct LkPadParams {
    LkLevelGeom src0;               // unpadded level 0
    LkLevelGeom lv[SVO_LK_LEVELS];  // padded planes (base = interior origin)
    int row0[SVO_LK_LEVELS + 1];    // first CTA of every level (a CTA handles PAD_ROWS of the h + 2 * PADY padded rows)
    int levels;
    int copy0;  // 1: entry 0 is level 0 and is copied whole from src0 (the unpadded ORB plane); else every entry is filled in place
};

__device__ __forceinline__ int reflect101_any(int i, int n) {
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

constexpr int PAD_ROWS = 8;  // padded rows per CTA of k_lk_pad

__global__ void __launch_bounds__(128) k_lk_pad(LkPadParams P, SlotMap map) {
    // P.row0[] counts CTAs here: level l owns CTAs [row0[l], row0[l+1]), each PAD_ROWS consecutive padded rows
    int l = 0;
    while (l + 1 < P.levels && (int)blockIdx.x >= P.row0[l + 1]) ++l;
    const LkLevelGeom& D = P.lv[l];
    const int slot = map.slot(blockIdx.y);
    const int r0 = ((int)blockIdx.x - P.row0[l]) * PAD_ROWS - SVO_LK_PADY;  // first padded row, interior coordinates
    const int nrows = min(PAD_ROWS, D.h + SVO_LK_PADY - r0);
    uint8_t* dbase = D.base + (size_t)slot * D.plane;
    // source: level 0 comes from the (unpadded) ORB plane, the other levels from their own interior
    const bool from0 = l == 0 && P.copy0;
    const uint8_t* __restrict__ sbase = from0 ? P.src0.base + (size_t)slot * P.src0.plane : dbase;
    const int spitch = from0 ? P.src0.pitch : D.pitch;
    // 16 threads per row, 8 rows per CTA: no index divisions, 256 contiguous bytes per row and step
    const int rr = threadIdx.x >> 4, t16 = threadIdx.x & 15;
    if (rr >= nrows) return;
    const int yp = r0 + rr;
    const uint8_t* __restrict__ srow = sbase + (ptrdiff_t)reflect101_any(yp, D.h) * spitch;
    uint8_t* drow = dbase + (ptrdiff_t)yp * D.pitch;
    if (from0 || yp < 0 || yp >= D.h) {
        // interior columns (all rows of level 0, the top / bottom border rows of the others): 16-byte copies (rows are 16-byte
        // aligned), four independent loads in flight per thread, then the ragged end byte-wise
        const int nv = D.w >> 4;
        for (int c0 = t16; c0 < nv; c0 += 64) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (c0 + 16 * u < nv) v[u] = __ldg(reinterpret_cast<const uint4*>(srow) + c0 + 16 * u);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (c0 + 16 * u < nv) reinterpret_cast<uint4*>(drow)[c0 + 16 * u] = v[u];
        }
        for (int x = (nv << 4) + t16; x < D.w; x += 16) drow[x] = __ldg(srow + x);
    }
    // left / right pads
    for (int k = t16; k < 2 * SVO_LK_PADX; k += 16) {
        const int x = k < SVO_LK_PADX ? k - SVO_LK_PADX : D.w + k - SVO_LK_PADX;
        drow[x] = __ldg(srow + reflect101_any(x, D.w));
    }
}

// K7c: Scharr derivatives of an LK level (cv::calcScharrDeriv as calcOpticalFlowPyrLK runs it per level): Ix | Iy << 16 per pixel.
//   V = 3 (I(x, y-1) + I(x, y+1)) + 10 I(x, y), W = I(x, y+1) - I(x, y-1);  Ix = V(x+1) - V(x-1),  Iy = 3 (W(x-1) + W(x+1)) + 10 W(x)
// Taps beyond the image read the REFLECT_101 border of the padded image plane; the derivative plane's own border stays ZERO
// (OpenCV pads derivI with BORDER_CONSTANT), which is what makes windows across the image border need no special case.
// One thread = 4 pixels of one row: 9 aligned words in, one 16-byte store out.
__global__ void __launch_bounds__(256) k_lk_deriv(LkLevelGeom S, LkLevelGeom D, SlotMap map) {
    const int slot = map.slot(blockIdx.z);
    const int x0 = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4, y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x0 >= S.w || y >= S.h) return;
    const uint8_t* __restrict__ src = S.base + (size_t)slot * S.plane + (ptrdiff_t)y * S.pitch + x0;  // 4-byte aligned: PADX % 4 == 0
    int v[6], w[6];  // columns x0 - 1 .. x0 + 4
    {
        uint32_t r[3][3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(src + (ptrdiff_t)(k - 1) * S.pitch);
            r[k][0] = __ldg(p - 1);
            r[k][1] = __ldg(p);
            r[k][2] = __ldg(p + 1);
        }
        auto px = [&](int k, int c) -> int {  // c = -1 .. 4
            return c < 0 ? (int)(r[k][0] >> 24) : (c < 4 ? (int)((r[k][1] >> (8 * c)) & 0xFFu) : (int)(r[k][2] & 0xFFu));
        };
#pragma unroll
        for (int c = -1; c <= 4; ++c) {
            const int up = px(0, c), mid = px(1, c), dn = px(2, c);
            v[c + 1] = 3 * (up + dn) + 10 * mid;
            w[c + 1] = dn - up;
        }
    }
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ix = v[i + 2] - v[i], iy = 3 * (w[i] + w[i + 2]) + 10 * w[i + 1];
        o[i] = x0 + i < S.w ? (((uint32_t)ix & 0xFFFFu) | ((uint32_t)iy << 16)) : 0u;  // the border of the plane stays zero
    }
    *reinterpret_cast<uint4*>(D.base + (size_t)slot * D.plane + (ptrdiff_t)y * D.pitch + (size_t)x0 * 4) = make_uint4(o[0], o[1], o[2], o[3]);
}

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// exact warp-wide sum of one int32 per lane as a 64-bit value: two hardware reductions (REDUX.SUM) on the 16-bit halves
__device__ __forceinline__ long long warp_sum64(int v) {
    const int lo = __reduce_add_sync(0xFFFFFFFFu, v & 0xFFFF);  // <= 32 * 65535
    const int hi = __reduce_add_sync(0xFFFFFFFFu, v >> 16);     // signed halves, |sum| <= 32 * 32768
    return ((long long)hi << 16) + (long long)lo;
}

__device__ __forceinline__ void weights(float a, float b, int& w00, int& w01, int& w10, int& w11) {
    const float one_a = __fsub_rn(1.f, a), one_b = __fsub_rn(1.f, b);
    w00 = __float2int_rn(__fmul_rn(__fmul_rn(one_a, one_b), 16384.f));
    w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, one_b), 16384.f));
    w10 = __float2int_rn(__fmul_rn(__fmul_rn(one_a, b), 16384.f));
    w11 = 16384 - w00 - w01 - w10;
}

// d = c + a.lo16 * b.byte0 + a.hi16 * b.byte1  (a: two signed 16-bit weights, b: two unsigned pixels) -- one IDP.2A
__device__ __forceinline__ int dp2(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// ---- per-warp shared-memory tile of an image neighbourhood ------------------------------------------------------------
// 24 rows x 64 bytes.  A row holds 8 pairs of words (T_w, T2_w): T is the image row starting at the neighbourhood's first
// column (the byte alignment of the global address is removed while staging), T2 the same row one byte later.  The two
// horizontally adjacent pixels (c, c+1) of the bilinear stencil are therefore always ONE aligned 16-bit shared-memory load
// (from T for even c, from T2 for odd c) that feeds IDP.2A directly -- no shuffles, no byte permutes, no address arithmetic
// per row (rows are an immediate offset apart).
// Rows of one 7-row segment are 64 bytes apart; every further segment is skewed by 16 bytes, so that the three segments a warp
// reads at the same time during the iterations (rows j, 7 + j, 14 + j of the same columns) fall into different banks.
constexpr int TROW = 64, TROWS = 24, TILE_BYTES = TROW * TROWS + 64;
__host__ __device__ constexpr int row_off(int r) { return r * TROW + (r / 7) * 16; }
constexpr int SEG = 7;           // iteration form: the window is cut into 21 columns x 3 segments of 7 rows = 63 strips; a lane
constexpr int NROUND = 2;        // owns strips (lane) and (lane + 32): 14 pixels, 2 x 8 pair loads (vertical reuse inside a strip)
// derivative tile of the setup (DERIV form): 22 rows x 22 words (Ix | Iy << 16 per pixel), rows 24 words apart, every 7-row segment
// skewed by 3 words so that the three segments a warp reads at once fall into different banks; it shares the image tile's memory
__host__ __device__ constexpr int drow_off(int r) { return r * 24 + (r / 7) * 3; }
constexpr int DTILE_BYTES = (drow_off(21) + 24) * 4;
// interpolated-image array of the setup (inside form): 23 rows x 23 words, every 7-row segment skewed by 10 words (7 * 23 + 10 = 11 mod 32:
// the strips (column x, rows 7 sg ..) a warp reads at once -- x = 0..10, sg = 0..2 -- fall into 32 different banks); it takes the
// image tile's place while the previous image's own tile sits in the (still unused) template area
__host__ __device__ constexpr int irow_off(int r) { return r * 23 + (r / 7) * 10; }
constexpr int ITILE_BYTES = (irow_off(22) + 23) * 4;
constexpr int TILE_MAX = DTILE_BYTES > ITILE_BYTES ? (DTILE_BYTES > TILE_BYTES ? DTILE_BYTES : TILE_BYTES) : (ITILE_BYTES > TILE_BYTES ? ITILE_BYTES : TILE_BYTES);
constexpr int TILE_AREA = ((TILE_MAX + 63) / 64) * 64;
static_assert(TILE_BYTES <= 32 * NROUND * SEG * 8, "the two-phase setup stages the previous image's tile in the template area");
constexpr int WARP_SMEM = TILE_AREA + 32 * NROUND * SEG * 8;

__device__ __forceinline__ int pair_off(int c) {  // byte offset in a tile row of the aligned u16 holding tile bytes (c, c+1)
    const int e = c & ~1;
    return ((e >> 2) << 3) + (e & 3) + ((c & 1) << 2);
}

__device__ __forceinline__ uint32_t lds_u16(const uint8_t* p) { return *reinterpret_cast<const unsigned short*>(p); }

// rows y0 .. y0+23, columns x0 .. x0+27 of the padded plane -> tile (every lane: one aligned word per 4 rows, 6 rounds)
__device__ __forceinline__ void stage_tile(uint8_t* tile, const uint8_t* __restrict__ img, int pitch, int x0, int y0, int lane) {
    const int a0 = x0 & 3, sh = 8 * a0;
    const uint8_t* __restrict__ p = img + ((ptrdiff_t)(y0 + (lane >> 3)) * pitch + (x0 - a0) + 4 * (lane & 7));
    const int rr = lane >> 3;
    uint8_t* dst = tile + (lane & 7) * 8;
    const ptrdiff_t step = (ptrdiff_t)4 * pitch;
#pragma unroll
    for (int k = 0; k < TROWS / 4; ++k) {
        const uint32_t W = __ldg(reinterpret_cast<const uint32_t*>(p + k * step));
        const uint32_t Wn = __shfl_down_sync(0xFFFFFFFFu, W, 1);  // next word of the same row (lane & 7 == 7: unused word)
        const int r = 4 * k + rr;                                   // (4k + rr) / 7 for rr < 4: one compare against a constant
        const int seg = (4 * k) / 7 + ((4 * k) % 7 + rr >= 7 ? 1 : 0);
        *reinterpret_cast<uint2*>(dst + r * TROW + seg * 16) = make_uint2(__funnelshift_r(W, Wn, sh), __funnelshift_rc(W, Wn, sh + 8));
    }
}

// element-wise stores into the template for the forms that produce one pixel at a time (e = byte address of entry (j, lane'), s = round)
__device__ __forceinline__ void tmpl_store_x(uint8_t* e, int s, uint32_t x) { *reinterpret_cast<uint32_t*>(e + 4 * s) = x; }
__device__ __forceinline__ void tmpl_store_g(uint8_t* e, int s, int ixv, int iyv) {
    *reinterpret_cast<unsigned short*>(e + 8 + 2 * s) = (unsigned short)ixv;
    *reinterpret_cast<unsigned short*>(e + 12 + 2 * s) = (unsigned short)iyv;
}

// d = c + a.lo16 * b.byte2 + a.hi16 * b.byte3 with SIGNED bytes (the high bytes of two differences) -- one IDP.2A.HI
__device__ __forceinline__ int dp2_hi_s8(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// One warp per point, all pyramid levels in one launch.
//  setup (previous image): the Scharr gradient and the bilinear interpolation -- both linear with integer weights -- are commuted
//    EXACTLY when every gradient tap lies inside the image: the staged 24 x 24 neighbourhood is interpolated once (Ib, un-descaled,
//    2 IDP.2A per value, lane = column) into shared memory and Scharr is applied to Ib in the iterations' strip layout with all 32
//    lanes.  Otherwise (window across the image border) the general form with OpenCV's constant-0 derivative outside the image
//    runs (lane = column, neighbours by shuffle).  The template (I, Ix, Iy) goes to shared memory in FLATTENED pixel order.
//  iterations (next image): the window is cut into 63 strips (21 columns x 3 segments of 7 rows), two per lane, walked side by side:
//    8 pair loads per strip (vertical reuse), and per PAIR of pixels one 128-bit template load, 4 IDP.2A + 2 shifts for the two
//    differences, one PRMT gathering their bytes and 4 IDP.2A for the four products with the packed gradient pairs.
// All window sums are exact integers; float expressions follow OpenCV's scalar order with _rn intrinsics.
template <bool DERIV>
__global__ void __launch_bounds__(LW * 32, 32) k_lk_track(LkParams P, const svo_keypoint* __restrict__ fast_kps, const int* __restrict__ nfast,
                                                      float* __restrict__ lk_pts, uint8_t* __restrict__ lk_status, float* __restrict__ lk_err) {
    __shared__ __align__(16) uint8_t s_mem[LW * WARP_SMEM];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.y;
    if (f == 0 && P.skip_first) return;
    // ring arithmetic without division: first, f < ring, so every index below is in [0, 2 * ring)
    const int cur_u = P.first + f, cur = cur_u >= P.ring ? cur_u - P.ring : cur_u;
    const int seed_fs = cur == 0 ? P.ring - 1 : cur - 1;  // seeds: FAST corners of the previous left image
    const int n = P.standalone ? P.n_fixed : nfast[seed_fs];
    const int pt_first = blockIdx.x * LW + warp;
    if (pt_first >= n) return;  // warp-uniform
    uint8_t* tile = s_mem + warp * WARP_SMEM;
    // template: entry (row-in-strip j, lane) = uint4 {x of round 0, x of round 1, Ix0 | Ix1 << 16, Iy0 | Iy1 << 16} at tmpl[j * 32 + lane],
    // x = 256 - 512 * I; strip g = 3 x + y / 7 of window pixel (x, y) belongs to lane g & 31 in round g >> 5.  The two rounds of a
    // lane sit side by side so that the iterations read both pixels with one LDS.128 and multiply BOTH gradients of the pair with
    // the pair's differences in IDP.2A (see the iteration loop).  Strip 63 (lane 31, round 1) does not exist: its gradient halves
    // are zero (written here and by every two-phase setup, never touched by the other forms), so it contributes nothing to the sums.
    uint4* tmpl = reinterpret_cast<uint4*>(tile + TILE_AREA);
    uint8_t* tmb = tile + TILE_AREA;
    const uint32_t tm32 = (uint32_t)__cvta_generic_to_shared(tmb);
    if (lane < SEG) tmpl[lane * 32 + 31] = make_uint4(0u, 0u, 0u, 0u);
    int offJ[NROUND];  // byte offset of the strip's first pair (column x, row 7 * seg) in a tile whose byte 0 is column x = 0
#pragma unroll
    for (int sr = 0; sr < NROUND; ++sr) {
        const int g = min(lane + 32 * sr, 62), x = g / 3, sg = g - 3 * x;
        offJ[sr] = row_off(SEG * sg) + pair_off(x);
    }
    int offB[NROUND];  // setup, inside form: word offset of the strip's first entry (column x, row 7 * seg) in the interpolated-image array
#pragma unroll
    for (int sr = 0; sr < NROUND; ++sr) {
        const int g = min(lane + 32 * sr, 62), x = g / 3, sg = g - 3 * x;
        offB[sr] = irow_off(SEG * sg) + x;
    }
    const int ccol = min(lane, 23);  // setup: tile column of this lane

    // the whole chain of calls for this point in one launch (tracking.cpp:593-618 chains four calls: each call's output list is
    // the next call's input, failed points included -- exactly what separate launches would read back from memory)
    // grid-stride over the points of the frame (one pass with the default grid of pt_cap CTAs per frame; see svo_ctx::lk_grid_x: the
    // 82 % of the 933 k one-warp CTAs of a batch that find no point cost nothing measurable)
    for (int pt = pt_first; pt < n; pt += gridDim.x * LW) {
    __syncwarp();  // the previous point's last shared-memory reads are done before its tile / template are overwritten
    float px = 0.f, py = 0.f;
    for (int ci = 0; ci < P.ncalls; ++ci) {
    const LkCall& call = P.calls[ci];
    const int pfs = call.prev_back ? seed_fs : cur, nfs = call.next_back ? seed_fs : cur;  // back is 0 or 1
    const int prev_slot = call.prev_side * P.ring + pfs, next_slot = call.next_side * P.ring + nfs;
    if (ci == 0) {
        if (!P.standalone && call.src_list == 0) {
            const svo_keypoint kp = fast_kps[(size_t)seed_fs * P.pt_cap + pt];
            px = kp.x;
            py = kp.y;
        } else {
            const float* src = lk_pts + ((size_t)cur * 5 + call.src_list) * P.pt_cap * 2;
            px = src[2 * pt];
            py = src[2 * pt + 1];
        }
    }
    float* dst = lk_pts + ((size_t)cur * 5 + call.dst_list) * P.pt_cap * 2;
    uint8_t* st_out = lk_status + ((size_t)cur * 4 + call.status_row) * P.pt_cap;

    const float half = 10.f;   // (win - 1) * 0.5
    float nx = 0.f, ny = 0.f;  // nextPts[pt] carried across levels
    int status = 1;
    float err = 0.f;
    const bool wlane = lane >= 1 && lane <= WIN;  // lanes that own a window column during the setup

    for (int level = P.levels - 1; level >= 0; --level) {
        const LkLevelGeom& G = P.lv[level];
        const uint8_t* __restrict__ I = G.base + (size_t)prev_slot * G.plane;
        const uint8_t* __restrict__ J = G.base + (size_t)next_slot * G.plane;
        const float sc = __int_as_float((127 - level) << 23);  // 1 / 2^level, exact (a float division costs ~15 instructions)
        float ppx = __fmul_rn(px, sc), ppy = __fmul_rn(py, sc);
        if (level == P.levels - 1) {
            nx = ppx;
            ny = ppy;
        } else {
            nx = __fmul_rn(nx, 2.f);
            ny = __fmul_rn(ny, 2.f);
        }
        ppx = __fsub_rn(ppx, half);
        ppy = __fsub_rn(ppy, half);
        const int ipx = __float2int_rd(ppx), ipy = __float2int_rd(ppy);
        if (ipx < -WIN || ipx >= G.w || ipy < -WIN || ipy >= G.h) {
            if (level == 0) {
                status = 0;
                err = 0.f;
            }
            continue;
        }
        int iw00, iw01, iw10, iw11;
        weights(__fsub_rn(ppx, (float)ipx), __fsub_rn(ppy, (float)ipy), iw00, iw01, iw10, iw11);
        uint32_t Wtop = __byte_perm((uint32_t)iw00, (uint32_t)iw01, 0x5410), Wbot = __byte_perm((uint32_t)iw10, (uint32_t)iw11, 0x5410);

        int a11 = 0, a12 = 0, a22 = 0;
        bool all_lanes = DERIV;
        if (DERIV) {
            // ---- MEASURED ALTERNATIVE (SVO_LK_DERIV=1; off by default): template from the image plane and a precomputed Scharr plane
            // (OpenCV's own order: derivI first, then the bilinear interpolation of I, Ix, Iy), in the iterations' strip layout: all 32
            // lanes, 14 pixels each, one code path for windows inside and across the image border (the Scharr plane has a zero border,
            // the image plane a REFLECT_101 one).  Parity-green (all GPU tests) and 18 % fewer SASS instructions, but the tracker is
            // co-limited by the L1 / shared-memory pipe: the 1.9 KB Scharr tile per level and call more than doubles its L1 / L2 traffic
            // (4.86 vs 4.82 ms) and the Scharr planes cost 0.24 ms per step to write -- 7.42 vs 7.14 ms per step, so the default stays the
            // commuted form below (Scharr of the interpolated image in registers, no derivative plane anywhere).
            // phase 1: I.  tile byte 0 = column ipx, row 0 = image row ipy (the same staging as the next image's tiles)
            __syncwarp();
            stage_tile(tile, I, G.pitch, ipx, ipy, lane);
            __syncwarp();
#pragma unroll
            for (int sr = 0; sr < NROUND; ++sr) {
                const uint8_t* tp = tile + offJ[sr];
                uint32_t pa = lds_u16(tp);
                const bool real = lane + 32 * sr < 63;
#pragma unroll
                for (int j = 0; j < SEG; ++j) {
                    const uint32_t pb = lds_u16(tp + (row_off(j + 1) - row_off(0)));
                    const int iv = dp2(Wbot, pb, dp2(Wtop, pa, 256)) >> 9;
                    if (real) tmpl_store_x(tmb + (j * 32 + lane) * 16, sr, (uint32_t)(256 - 512 * iv));
                    pa = pb;
                }
            }
            // phase 2: Ix, Iy.  22 x 22 words of the Scharr plane from (ipx, ipy) on; lane c loads column c
            __syncwarp();
            {
                const LkLevelGeom& GD = P.dv[level];
                const uint32_t* __restrict__ dsrc = reinterpret_cast<const uint32_t*>(GD.base + (size_t)prev_slot * GD.plane + (ptrdiff_t)ipy * GD.pitch) + ipx + lane;
                uint32_t* dt = reinterpret_cast<uint32_t*>(tile);
                if (lane < 22) {
#pragma unroll
                    for (int r = 0; r < 22; ++r) dt[drow_off(r) + lane] = __ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(dsrc) + (ptrdiff_t)r * GD.pitch));
                }
            }
            __syncwarp();
            {
                const uint32_t* dt = reinterpret_cast<const uint32_t*>(tile);
#pragma unroll
                for (int sr = 0; sr < NROUND; ++sr) {
                    const int g = min(lane + 32 * sr, 62), x = g / 3, sg = g - 3 * x;
                    const uint32_t* dp = dt + drow_off(SEG * sg) + x;
                    const bool real = lane + 32 * sr < 63;
                    int ptx = 0, pty = 0;  // top-row combinations of the previous tile row
#pragma unroll
                    for (int j = 0; j <= SEG; ++j) {
                        const int ro = drow_off(j);  // relative to the strip's first row: 24 words per row, row 7 carries the next segment's skew
                        const uint32_t wA = dp[ro], wB = dp[ro + 1];
                        const int ixA = (int)(short)(wA & 0xFFFFu), iyA = (int)wA >> 16, ixB = (int)(short)(wB & 0xFFFFu), iyB = (int)wB >> 16;
                        if (j >= 1) {
                            const int ixv = (ptx + ixA * iw10 + ixB * iw11 + (1 << 13)) >> 14;
                            const int iyv = (pty + iyA * iw10 + iyB * iw11 + (1 << 13)) >> 14;
                            if (real) {
                                tmpl_store_g(tmb + ((j - 1) * 32 + lane) * 16, sr, ixv, iyv);
                                a11 += ixv * ixv;
                                a12 += ixv * iyv;
                                a22 += iyv * iyv;
                            }
                        }
                        if (j < SEG) {
                            ptx = ixA * iw00 + ixB * iw01;
                            pty = iyA * iw00 + iyB * iw01;
                        }
                    }
                }
            }
        } else {
        // ---- previous image: tile byte 0 = column ipx - 1, tile row 0 = image row ipy - 1
        // every gradient tap (columns ipx .. ipx+21, rows ipy .. ipy+21) inside the image?  (warp-uniform)
        const bool inside = ipx >= 0 && ipx + WIN < G.w && ipy >= 0 && ipy + WIN < G.h;
        const bool two_phase = inside;
        uint8_t* itile = two_phase ? reinterpret_cast<uint8_t*>(tmpl) : tile;  // (two-phase form: the tile area receives Ib)
        __syncwarp();
        stage_tile(itile, I, G.pitch, ipx - 1, ipy - 1, lane);
        __syncwarp();
        if (two_phase) {
            all_lanes = true;
            // Inside form in two phases through shared memory (all 32 lanes busy in the second):
            //   A  lane = tile column c <= 22 walks down the rows: Ib(r, c) = interpolated image, un-descaled, + 256 (the rounding
            //      constant of I's descale; it cancels in both Scharr sums)                              -> ibw[irow_off(r) + c]
            //   B  the iterations' strip layout (lane = column x, 7 rows; two strips per lane): per row of Ib three loads (columns
            //      x, x+1, x+2), dH = R - L, S = 3 (L + R) + 10 C; per pixel Ix = 3 (dH(y) + dH(y+2)) + 10 dH(y+1),
            //      Iy = S(y+2) - S(y), I from C(y+1).  ~420 warp-instructions; the one-phase form it replaced (lane = column, neighbours by
            //      shuffle: 24 active lanes spending 23 rows on 21 rows of output) took ~540 (LK chain 4.79 -> 4.68 ms).
            int* ibw = reinterpret_cast<int*>(tile);
            {
                const uint8_t* tp = itile + pair_off(min(lane, 22));
                int* iw = ibw + min(lane, 22);
                int top = dp2(Wtop, lds_u16(tp), 256);
#pragma unroll
                for (int r = 1; r < TROWS; ++r) {
                    const uint32_t pr = lds_u16(tp + row_off(r));
                    const int ib = dp2(Wbot, pr, top);
                    top = dp2(Wtop, pr, 256);
                    if (lane < 23) iw[irow_off(r - 1)] = ib;
                }
            }
            __syncwarp();
            // both strips of the lane side by side: one STS.128 per template entry (the pair the iterations read with one LDS.128)
            {
                const int* ip0 = ibw + offB[0];
                const int* ip1 = ibw + offB[1];
                uint4* to = tmpl + lane;
                const int m1 = lane < 31 ? -1 : 0;  // strip 63 does not exist: zero gradient
                int dHa0 = 0, dHb0 = 0, Sa0 = 0, Sb0 = 0, Cb0 = 0, dHa1 = 0, dHb1 = 0, Sa1 = 0, Sb1 = 0, Cb1 = 0;
#pragma unroll
                for (int j = 0; j < SEG + 2; ++j) {
                    const int o = irow_off(j);  // relative to the strip's first row (7 sg): 23 words per row, rows 7 and 8 carry the next segment's skew
                    const int L0 = ip0[o], C0 = ip0[o + 1], R0 = ip0[o + 2];
                    const int L1 = ip1[o], C1 = ip1[o + 1], R1 = ip1[o + 2];
                    const int dH0 = R0 - L0, S0 = 3 * (L0 + R0) + 10 * C0;
                    const int dH1 = R1 - L1, S1 = 3 * (L1 + R1) + 10 * C1;
                    if (j >= 2) {  // window row 7 sg + j - 2
                        const int ixv0 = (3 * (dHa0 + dH0) + 10 * dHb0 + (1 << 13)) >> 14;
                        const int iyv0 = (S0 - Sa0 + (1 << 13)) >> 14;
                        const int ixv1 = ((3 * (dHa1 + dH1) + 10 * dHb1 + (1 << 13)) >> 14) & m1;
                        const int iyv1 = ((S1 - Sa1 + (1 << 13)) >> 14) & m1;
                        to[(j - 2) * 32] = make_uint4((uint32_t)(256 - (Cb0 & ~511)), (uint32_t)(256 - (Cb1 & ~511)),
                                                      __byte_perm((uint32_t)ixv0, (uint32_t)ixv1, 0x5410), __byte_perm((uint32_t)iyv0, (uint32_t)iyv1, 0x5410));
                        a11 += ixv0 * ixv0 + ixv1 * ixv1;
                        a12 += ixv0 * iyv0 + ixv1 * iyv1;
                        a22 += iyv0 * iyv0 + iyv1 * iyv1;
                    }
                    dHa0 = dHb0; dHb0 = dH0; Sa0 = Sb0; Sb0 = S0; Cb0 = C0;
                    dHa1 = dHb1; dHb1 = dH1; Sa1 = Sb1; Sb1 = S1; Cb1 = C1;
                }
            }
        } else {
            // general form: Scharr gradient per image position, zero outside the image (OpenCV's constant-0 padded
            // derivative), then the bilinear interpolation of I, Ix, Iy separately
            const int X = ipx - 1 + ccol;
            const bool xin = X >= 0 && X < G.w;
            const uint8_t* tb = tile + ((ccol >> 2) << 3) + (ccol & 3);  // T byte of this lane's column
            const uint8_t* tp = tile + pair_off(ccol);
            int p0 = tb[0], p1 = tb[row_off(1)];
            int top = 256, tgx = 0, tgxn = 0, tgy = 0, tgyn = 0;
            // one row of the general form: the Scharr gradient at (X, Y = ipy + r) (tile row r + 1) from tile rows r .. r + 2, and -- for
            // r >= 1 -- the template entry of window row r - 1 from the gradients / pixels of rows r - 1 and r
            auto grow = [&](int r, int p2, uint32_t pr, bool emit, uint32_t te) {  // emit: window row r - 1 goes to shared-memory address te
                const int V = 3 * (p0 + p2) + 10 * p1, Dv = p2 - p0;
                const int Vl = __shfl_up_sync(0xFFFFFFFFu, V, 1), Vr = __shfl_down_sync(0xFFFFFFFFu, V, 1);
                const int Dl = __shfl_up_sync(0xFFFFFFFFu, Dv, 1), Dr = __shfl_down_sync(0xFFFFFFFFu, Dv, 1);
                const int Y = ipy + r;
                const bool in = xin && Y >= 0 && Y < G.h;
                const int gx = in ? Vr - Vl : 0;
                const int gy = in ? 3 * (Dl + Dr) + 10 * Dv : 0;
                const int gxn = __shfl_down_sync(0xFFFFFFFFu, gx, 1), gyn = __shfl_down_sync(0xFFFFFFFFu, gy, 1);
                if (emit) {
                    const int ib = dp2(Wbot, pr, top);  // interpolated pixel + 256
                    const int ixv = descale(tgx * iw00 + tgxn * iw01 + gx * iw10 + gxn * iw11, 14);
                    const int iyv = descale(tgy * iw00 + tgyn * iw01 + gy * iw10 + gyn * iw11, 14);
                    if (wlane)  // (a plain 32-bit shared address: as a generic pointer the compiler re-derived the shared window base at every store)
                        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(te), "r"((uint32_t)(256 - (ib & ~511))),
                                     "r"(__byte_perm((uint32_t)ixv, (uint32_t)iyv, 0x5410))
                                     : "memory");
                    a11 += ixv * ixv;
                    a12 += ixv * iyv;
                    a22 += iyv * iyv;
                }
                top = dp2(Wtop, pr, 256);
                tgx = gx; tgxn = gxn; tgy = gy; tgyn = gyn;
                p0 = p1;
                p1 = p2;
            };
            grow(0, tb[row_off(2)], lds_u16(tp + row_off(1)), false, 0u);
            // rows 1 .. 21 = three segments of 7: the segment loop stays rolled (this form runs for ~1 setup in 5; unrolling all 22
            // rows made the kernel's executed code outgrow the instruction cache once the err pass ran too)
#pragma unroll 1
            for (int sg = 0; sg < 3; ++sg) {
                const uint8_t* tbs = tb + sg * (row_off(SEG) - row_off(0));
                const uint8_t* tps = tp + sg * (row_off(SEG) - row_off(0));
                // this lane's strip of the segment belongs to lane g & 31, round g >> 5: the pixel goes there as (x, Ix | Iy << 16) first
                const int g = 3 * (lane - 1) + sg;
                const uint32_t te = tm32 + (uint32_t)((g & 31) * 16 + (g >> 5) * 8);
#pragma unroll
                for (int jj = 0; jj < SEG; ++jj) grow(1 + SEG * sg + jj, tbs[row_off(3 + jj)], lds_u16(tps + row_off(2 + jj)), true, te + jj * 512);
            }
            // {x0, g0, x1, g1} -> {x0, x1, Ix0 | Ix1 << 16, Iy0 | Iy1 << 16}: every lane repacks its own seven entries
            __syncwarp();
#pragma unroll
            for (int j = 0; j < SEG; ++j) {
                uint4 v = tmpl[j * 32 + lane];
                if (lane == 31) v.w = 0u;  // strip 63 does not exist: zero gradient
                tmpl[j * 32 + lane] = make_uint4(v.x, v.z, __byte_perm(v.y, v.w, 0x5410), __byte_perm(v.y, v.w, 0x7632));
            }
        }
        }  // !DERIV
        const bool alane = all_lanes || wlane;  // strip forms: every lane holds window pixels
        const long long sA11 = warp_sum64(alane ? a11 : 0), sA12 = warp_sum64(alane ? a12 : 0), sA22 = warp_sum64(alane ? a22 : 0);
        const float FLT_SCALE = 1.f / (1 << 20);
        const float A11 = __fmul_rn((float)sA11, FLT_SCALE), A12 = __fmul_rn((float)sA12, FLT_SCALE), A22 = __fmul_rn((float)sA22, FLT_SCALE);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dA = __fsub_rn(A11, A22);
        const float minEig = __fdiv_rn(
            __fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(__fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12)))), (float)(2 * WIN * WIN));
        if (minEig < P.min_eig || D < FLT_EPSILON) {
            if (level == 0) status = 0;
            continue;
        }
        D = __fdiv_rn(1.f, D);
        float outx = nx, outy = ny;  // nextPts[pt] as stored before the loop (untouched if the loop exits at once)
        nx = __fsub_rn(nx, half);
        ny = __fsub_rn(ny, half);
        float pdx = 0.f, pdy = 0.f;
        const uint4* tl = tmpl + lane;
        int tx = INT_MIN, ty = INT_MIN;  // origin of the staged next-image tile (none yet: the tile holds the previous image)
        for (int j = 0; j < P.max_iter; ++j) {
            const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
            if ((unsigned)inx + (unsigned)WIN >= (unsigned)(G.w + WIN) || (unsigned)iny + (unsigned)WIN >= (unsigned)(G.h + WIN)) {  // outside [-WIN, size)
                if (level == 0) status = 0;
                break;
            }
            weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), iw00, iw01, iw10, iw11);
            Wtop = __byte_perm((uint32_t)iw00, (uint32_t)iw01, 0x5410);
            Wbot = __byte_perm((uint32_t)iw10, (uint32_t)iw11, 0x5410);
            if (inx != tx || iny != ty) {  // sub-pixel steps usually stay in the same integer cell: the staged tile is still valid
                __syncwarp();              // the previous tile (and, in the first iteration, the template stores) are complete
                stage_tile(tile, J, G.pitch, inx, iny, lane);
                __syncwarp();
                tx = inx;
                ty = iny;
            }
            // Both strips of the lane side by side.  A difference d (|d| <= 8160) is 256 * (signed byte 1) + (unsigned byte 0); one PRMT
            // gathers the four bytes of the pair's differences, and each gradient pair (Ix0 | Ix1 << 16 resp. Iy0 | Iy1 << 16, as the
            // template stores them) takes two IDP.2A: low bytes (unsigned) and high bytes (signed) -- 5 instructions for the four
            // products of a pixel pair instead of two sign extensions and two IMAD per pixel.  The sums are the same exact integers.
            int b1l = 0, b1h = 0, b2l = 0, b2h = 0;
            {
                const uint8_t* tp0 = tile + offJ[0];
                const uint8_t* tp1 = tile + offJ[1];
                uint32_t pa0 = lds_u16(tp0), pa1 = lds_u16(tp1);  // the pair of row j; row j + 1's pair is reused as the next row's upper pair
#pragma unroll
                for (int j = 0; j < SEG; ++j) {
                    const uint32_t pb0 = lds_u16(tp0 + (row_off(j + 1) - row_off(0))), pb1 = lds_u16(tp1 + (row_off(j + 1) - row_off(0)));
                    const uint4 t = tl[j * 32];
                    const int d0 = dp2(Wbot, pb0, dp2(Wtop, pa0, (int)t.x)) >> 9;  // descale(J interpolated, 9) - I
                    const int d1 = dp2(Wbot, pb1, dp2(Wtop, pa1, (int)t.y)) >> 9;
                    const uint32_t dd = __byte_perm((uint32_t)d0, (uint32_t)d1, 0x5140);  // (d0.b0, d1.b0, d0.b1, d1.b1)
                    b1l = dp2(t.z, dd, b1l);
                    b1h = dp2_hi_s8(t.z, dd, b1h);
                    b2l = dp2(t.w, dd, b2l);
                    b2h = dp2_hi_s8(t.w, dd, b2h);
                    pa0 = pb0;
                    pa1 = pb1;
                }
            }
            // |diff| <= 8160 (pixels carry 5 fraction bits) and |Ix|, |Iy| <= 4080 (Scharr of 8-bit pixels): the window sum can reach
            // 441 * 8160 * 4080 ~ 2^34 (sharp stripes one pixel out of register), but its two parts fit 32 bits -- low bytes
            // <= 441 * 255 * 4080 < 2^29, high bytes <= 441 * 32 * 4080 < 2^26 -- so the exact sum is four hardware reductions
            const int r1l = __reduce_add_sync(0xFFFFFFFFu, b1l), r1h = __reduce_add_sync(0xFFFFFFFFu, b1h);
            const int r2l = __reduce_add_sync(0xFFFFFFFFu, b2l), r2h = __reduce_add_sync(0xFFFFFFFFu, b2h);
            const long long sb1 = ((long long)r1h << 8) + (long long)r1l, sb2 = ((long long)r2h << 8) + (long long)r2l;
            const float b1f = __fmul_rn((float)sb1, FLT_SCALE), b2f = __fmul_rn((float)sb2, FLT_SCALE);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2f), __fmul_rn(A22, b1f)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1f), __fmul_rn(A11, b2f)), D);
            nx = __fadd_rn(nx, dx);
            ny = __fadd_rn(ny, dy);
            outx = __fadd_rn(nx, half);
            outy = __fadd_rn(ny, half);
            // OpenCV: delta.ddot(delta) <= epsilon in double.  The float value (three roundings, < 3e-7 relative) decides the same
            // way unless it lies within 1e-6 of the threshold; only then the double expression is evaluated (NaN ends up there too)
            const float d2 = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
            if (d2 < P.eps2_lo) break;
            if (!(d2 > P.eps2_hi) && (double)dx * (double)dx + (double)dy * (double)dy <= P.eps2) break;
            // (double)|v| < 0.01 for a float v  <=>  |v| <= 0.01f: 0.01f is the largest float below 0.01
            if (j > 0 && fabsf(__fadd_rn(dx, pdx)) <= 0.01f && fabsf(__fadd_rn(dy, pdy)) <= 0.01f) {
                outx = __fsub_rn(outx, __fmul_rn(dx, 0.5f));
                outy = __fsub_rn(outy, __fmul_rn(dy, 0.5f));
                break;
            }
            pdx = dx;
            pdy = dy;
        }
        nx = outx;  // what OpenCV leaves in nextPts[pt] for the next (finer) level
        ny = outy;
        if (status && level == 0) {  // final range check + error (the reference asks for err, so this block runs)
            const float fx = __fsub_rn(nx, half), fy = __fsub_rn(ny, half);
            const int inx = __float2int_rd(fx), iny = __float2int_rd(fy);
            if (inx < -WIN || inx >= G.w || iny < -WIN || iny >= G.h) {
                status = 0;
            } else if (P.want_err) {
                weights(__fsub_rn(fx, (float)inx), __fsub_rn(fy, (float)iny), iw00, iw01, iw10, iw11);
                Wtop = __byte_perm((uint32_t)iw00, (uint32_t)iw01, 0x5410);
                Wbot = __byte_perm((uint32_t)iw10, (uint32_t)iw11, 0x5410);
                if (inx != tx || iny != ty) {
                    __syncwarp();
                    stage_tile(tile, J, G.pitch, inx, iny, lane);
                    __syncwarp();
                }
                int e = 0;
                {   // both strips side by side like the iterations: one 64-bit load brings the pair's two x = 256 - 512 * I
                    const uint8_t* tp0 = tile + offJ[0];
                    const uint8_t* tp1 = tile + offJ[1];
                    uint32_t pa0 = lds_u16(tp0), pa1 = lds_u16(tp1);
                    int e1 = 0;
#pragma unroll
                    for (int j = 0; j < SEG; ++j) {
                        const uint32_t pb0 = lds_u16(tp0 + (row_off(j + 1) - row_off(0))), pb1 = lds_u16(tp1 + (row_off(j + 1) - row_off(0)));
                        const uint2 tx2 = *reinterpret_cast<const uint2*>(tl + j * 32);
                        e += abs(dp2(Wbot, pb0, dp2(Wtop, pa0, (int)tx2.x)) >> 9);
                        e1 += abs(dp2(Wbot, pb1, dp2(Wtop, pa1, (int)tx2.y)) >> 9);
                        pa0 = pb0;
                        pa1 = pb1;
                    }
                    if (lane < 31) e += e1;  // strip 63 does not exist
                }
                const int se = __reduce_add_sync(0xFFFFFFFFu, e);  // <= 441 * 8160
                err = __fdiv_rn((float)se, (float)(32 * WIN * WIN));
            }
        }
    }
    if (lane == 0) {
        dst[2 * pt] = nx;
        dst[2 * pt + 1] = ny;
        st_out[pt] = (uint8_t)status;
        if (P.want_err && lk_err) lk_err[((size_t)cur * 4 + call.status_row) * P.pt_cap + pt] = status ? err : 0.f;
    }
    px = nx;  // source point of the next call of the chain
    py = ny;
    }  // calls
    }  // points
}

}  // namespace

int launch_lk_pyramid(svo_ctx* c, const SlotMap& map, int n_img) {
    StageTimer t(c, 7);
    // level 0: padded copy of the ORB level-0 plane; then level l = pyrDown(padded level l - 1), and its own border right after
    // (the next level's taps read it)
    auto pad = [&](int l) {
        LkPadParams P{};
        P.src0 = c->lk[0];
        P.lv[0] = l == 0 ? c->lkp0 : c->lk[l];
        P.levels = 1;
        P.copy0 = l == 0;
        P.row0[0] = 0;
        P.row0[1] = (P.lv[0].h + 2 * SVO_LK_PADY + PAD_ROWS - 1) / PAD_ROWS;
        k_lk_pad<<<dim3(P.row0[1], n_img), 128, 0, c->stream>>>(P, map);
        ++c->launches;
    };
    pad(0);
    for (int l = 1; l < c->lk_levels; ++l) {
        const LkLevelGeom& S = l == 1 ? c->lkp0 : c->lk[l - 1];
        const LkLevelGeom& D = c->lk[l];
        dim3 grid((D.w + 255) / 256, (D.h + 7) / 8, n_img);
        k_pyr_down<<<grid, 256, 0, c->stream>>>(S, D, map);
        ++c->launches;
        pad(l);
    }
    LkPadParams P{};  // (geometry of all levels for the Scharr-plane alternative below)
    P.lv[0] = c->lkp0;
    for (int l = 1; l < c->lk_levels; ++l) P.lv[l] = c->lk[l];
    if (c->lk_deriv) {  // Scharr planes of all levels (the tracker's setup interpolates them)
        for (int l = 0; l < c->lk_levels; ++l) {
            const LkLevelGeom& S = P.lv[l];
            k_lk_deriv<<<dim3((S.w + 255) / 256, (S.h + 3) / 4, n_img), 256, 0, c->stream>>>(S, c->lkd[l], map);
            ++c->launches;
        }
    }
    SVO_CUDA_TRY(c, cudaGetLastError());
    return SVO_OK;
}

static void fill_params(svo_ctx* c, LkParams& P, int max_level, int max_iter, double eps, double min_eig) {
    P.lv[0] = c->lkp0;
    for (int l = 1; l < SVO_LK_LEVELS; ++l) P.lv[l] = c->lk[l];
    for (int l = 0; l < SVO_LK_LEVELS; ++l) P.dv[l] = c->lkd[l];
    P.levels = std::min(c->lk_levels, max_level + 1);
    P.max_iter = std::min(std::max(max_iter, 0), 100);
    eps = std::min(std::max(eps, 0.), 10.);
    P.eps2 = eps * eps;
    P.eps2_lo = (float)(P.eps2 * (1.0 - 1e-6));
    P.eps2_hi = (float)(P.eps2 * (1.0 + 1e-6));
    P.min_eig = (float)min_eig;
    P.pt_cap = c->pt_cap;
}

int launch_lk_chain(svo_ctx* c, int first, int ring, int n_frames, int skip_first, const LkCall* calls, int ncalls, int max_iter, double eps,
                    double min_eig, int want_err) {
    StageTimer t(c, 8);
    LkParams P{};
    fill_params(c, P, SVO_LK_LEVELS - 1, max_iter, eps, min_eig);
    P.first = first;
    P.ring = ring;
    P.skip_first = skip_first;
    P.ncalls = std::min(ncalls, 4);
    for (int k = 0; k < P.ncalls; ++k) {
        P.calls[k] = calls[k];
        if (k && calls[k].src_list != calls[k - 1].dst_list) return c->fail(SVO_ERR_INVALID, "launch_lk_chain: calls do not form a chain");
    }
    P.standalone = 0;
    P.want_err = want_err;
    dim3 grid(std::min((c->pt_cap + LW - 1) / LW, c->lk_grid_x), n_frames);
    auto kern = c->lk_deriv ? k_lk_track<true> : k_lk_track<false>;
    kern<<<grid, LW * 32, 0, c->stream>>>(P, c->d_fast_kps, c->d_nfast, c->d_lk_pts, c->d_lk_status, want_err ? c->d_lk_err : nullptr);
    ++c->launches;
    SVO_CUDA_TRY(c, cudaGetLastError());
    return SVO_OK;
}

// standalone call: prev image in slot 0, next image in slot 1 (ring 1), points in list 0 -> list 1, status row 0
int launch_lk_standalone(svo_ctx* c, int n, int max_level, int max_iter, double eps, double min_eig, int want_err) {
    StageTimer t(c, 8);
    LkParams P{};
    fill_params(c, P, max_level, max_iter, eps, min_eig);
    P.first = 0;
    P.ring = 1;
    P.skip_first = 0;
    P.calls[0] = LkCall{0, 0, 1, 0, 0, 1, 0};
    P.ncalls = 1;
    P.standalone = 1;
    P.n_fixed = n;
    P.want_err = want_err;
    dim3 grid((n + LW - 1) / LW, 1);
    auto kern = c->lk_deriv ? k_lk_track<true> : k_lk_track<false>;
    kern<<<grid, LW * 32, 0, c->stream>>>(P, c->d_fast_kps, c->d_nfast, c->d_lk_pts, c->d_lk_status, c->d_lk_err);
    ++c->launches;
    SVO_CUDA_TRY(c, cudaGetLastError());
    return SVO_OK;
}
