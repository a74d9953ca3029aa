// k_lk.cu -- K7 + K8: pyramidal Lucas-Kanade tracking == cv::calcOpticalFlowPyrLK as called at tracking.cpp:593-618
// (winSize 21x21, maxLevel 3, TermCriteria(COUNT+EPS, 30, 0.01), flags 0, minEigThreshold 1e-3); SURVEY.md A.9.
//
//  K7  k_pyr_down   cv::pyrDown: 5x5 [1 4 6 4 1]^2 / 256, BORDER_REFLECT_101, round half up; levels 1..3 of each image.
//                   Built once per image (the reference rebuilds both pyramids inside each of its 4 calls).
//  K8  k_lk_track   one warp per point, all pyramid levels in one launch, fully register-resident (no shared memory): lane l
//                   owns one image column of the 24x24 neighbourhood of the previous image (REFLECT_101 outside the
//                   image), derives the Scharr gradients on the fly with a sliding 3-row window + warp shuffles
//                   (constant 0 outside the image, exactly like OpenCV's padded derivative buffer -- no derivative image is
//                   ever written to HBM), keeps its column of the 21x21 interpolated window (I, Ix, Iy as int16) in
//                   registers and iterates Gauss-Newton on 22x22 neighbourhoods of the next image read straight from
//                   L1/L2.  All window sums are exact integers (int64 across the warp), converted to float once.
//
// f32 arithmetic uses explicit round-to-nearest intrinsics (no FMA contraction) to follow OpenCV's scalar expressions.
#include <cfloat>

#include "svo_common.cuh"

namespace {

constexpr int WIN = SVO_LK_WIN;  // 21
constexpr int LW = 8;            // warps (points) per CTA

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

// cv::pyrDown: each thread produces 4 x 2 destination pixels (7 source rows, the middle 3 shared by both output rows)
__global__ void __launch_bounds__(256) k_pyr_down(LkLevelGeom S, LkLevelGeom D, SlotMap map) {
    const int slot = map.slot(blockIdx.z);
    const int x0 = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4, y0 = (blockIdx.y * 4 + (threadIdx.x >> 6)) * 2;
    if (x0 >= D.w || y0 >= D.h) return;
    const uint8_t* __restrict__ src = S.base + (size_t)slot * S.plane;
    // all 4 pixels' taps inside the row and the 4 aligned words inside the pitch
    const bool interior = 2 * x0 - 4 >= 0 && 2 * x0 + 12 <= S.w;
    int hr[7][4];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const int sy = reflect101(min(2 * y0 + i - 2, S.h + 1), S.h);  // rows beyond the second output row's taps are unused
        pyr_hpass4(src + (size_t)sy * S.pitch, x0, S.w, interior, hr[i]);
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
    LkLevelGeom lv[SVO_LK_LEVELS];
    int levels;      // pyramid levels available (maxLevel + 1)
    int first, ring, skip_first;
    LkCall call;
    int standalone;  // 1: src points are a plain xy list (list index call.src_list), n given by n_fixed
    int n_fixed;
    int max_iter;
    double eps2;
    float min_eig;
    int want_err;
    int pt_cap;
};

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

// rows [y0, y0 + N) of image column x (REFLECT_101 outside the image); `interior` is warp-uniform: no reflection needed
template <int N>
__device__ __forceinline__ void load_column(const uint8_t* __restrict__ img, const LkLevelGeom& G, int x, int y0, bool interior, int (&out)[N]) {
    if (interior) {
        // 32-bit offsets from the (warp-uniform) image base: one integer add per row instead of 64-bit pointer arithmetic
        const unsigned o0 = (unsigned)y0 * (unsigned)G.pitch + (unsigned)x, pitch = (unsigned)G.pitch;
#pragma unroll
        for (int r = 0; r < N; ++r) out[r] = __ldg(img + (o0 + (unsigned)r * pitch));
    } else {
        const uint8_t* __restrict__ col = img + reflect101(x, G.w);
#pragma unroll
        for (int r = 0; r < N; ++r) out[r] = __ldg(col + (size_t)reflect101(y0 + r, G.h) * G.pitch);
    }
}

// One warp per point; lane l owns image column (ip.x - 1 + l) of the 24-column neighbourhood, so window column x lives in
// lane x + 1.  Everything is register-resident: the lane walks down its column (sliding 3-row window for the Scharr
// gradient), horizontal neighbours come from warp shuffles, and the interpolated window (I, Ix, Iy of 21 rows) stays in
// registers for the Gauss-Newton iterations.  No shared memory, no derivative image.
__global__ void __launch_bounds__(LW * 32, 3) k_lk_track(LkParams P, const svo_keypoint* __restrict__ fast_kps, const int* __restrict__ nfast,
                                                      float* __restrict__ lk_pts, uint8_t* __restrict__ lk_status, float* __restrict__ lk_err) {
    __shared__ short s_Iw[LW][WIN * 32];
    __shared__ uint32_t s_G[LW][WIN * 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.y;
    if (f == 0 && P.skip_first) return;
    // ring arithmetic without division: first, f < ring, so every index below is in [0, 2 * ring)
    const int cur_u = P.first + f, cur = cur_u >= P.ring ? cur_u - P.ring : cur_u;
    const int seed_fs = cur == 0 ? P.ring - 1 : cur - 1;  // seeds: FAST corners of the previous left image
    const int n = P.standalone ? P.n_fixed : nfast[seed_fs];
    const int pt = blockIdx.x * LW + warp;
    if (pt >= n) return;  // warp-uniform; most warps of the fixed-size grid leave here
    const int pfs = P.call.prev_back ? seed_fs : cur, nfs = P.call.next_back ? seed_fs : cur;  // back is 0 or 1
    const int prev_slot = P.call.prev_side * P.ring + pfs, next_slot = P.call.next_side * P.ring + nfs;

    float px, py;
    if (!P.standalone && P.call.src_list == 0) {
        const svo_keypoint kp = fast_kps[(size_t)seed_fs * P.pt_cap + pt];
        px = kp.x;
        py = kp.y;
    } else {
        const float* src = lk_pts + ((size_t)cur * 5 + P.call.src_list) * P.pt_cap * 2;
        px = src[2 * pt];
        py = src[2 * pt + 1];
    }
    float* dst = lk_pts + ((size_t)cur * 5 + P.call.dst_list) * P.pt_cap * 2;
    uint8_t* st_out = lk_status + ((size_t)cur * 4 + P.call.status_row) * P.pt_cap;

    const float half = 10.f;   // (win - 1) * 0.5
    float nx = 0.f, ny = 0.f;  // nextPts[pt] carried across levels
    int status = 1;
    float err = 0.f;
    const bool wlane = lane >= 1 && lane <= WIN;  // lanes that own a window column

    for (int level = P.levels - 1; level >= 0; --level) {
        const LkLevelGeom& G = P.lv[level];
        const uint8_t* __restrict__ I = G.base + (size_t)prev_slot * G.plane;
        const uint8_t* __restrict__ J = G.base + (size_t)next_slot * G.plane;
        const float sc = 1.f / (float)(1 << level);
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

        // ---- previous image: column X = ipx - 1 + lane, rows ipy - 1 .. ipy + 22 (REFLECT_101 outside the image)
        // the interpolated window lives in shared memory (lane-major, conflict-free): keeps the kernel at 3 CTAs / SM
        short* Iw = s_Iw[warp] + lane;      // Iw[y * 32]
        uint32_t* Gw = s_G[warp] + lane;    // (Ix, Iy) int16 pair at Gw[y * 32]
        int a11 = 0, a12 = 0, a22 = 0;
        const bool inside = ipx - 1 >= 0 && ipx + 22 < G.w && ipy - 1 >= 0 && ipy + 22 < G.h;  // warp-uniform
        if (inside) {
            // Whole neighbourhood inside the image: no reflection, no zeroed gradients.  The Scharr operator and the bilinear
            // interpolation are both linear with integer weights, so they commute exactly: interpolate the IMAGE once
            // (Ib, un-descaled) and apply Scharr to Ib -- 4 multiplies per position instead of 12.
            //   sum_ij w_ij Ix(x+i, y+j) = Vb(x+1, y) - Vb(x-1, y),          Vb = 3 Ib(y-1) + 10 Ib(y) + 3 Ib(y+1)
            //   sum_ij w_ij Iy(x+i, y+j) = 3 (Db(x-1, y) + Db(x+1, y)) + 10 Db(x, y),   Db = Ib(y+1) - Ib(y-1)
            int pix[24];
            load_column<24>(I, G, ipx - 1 + min(lane, 23), ipy - 1, true, pix);
            int pn0 = __shfl_down_sync(0xFFFFFFFFu, pix[0], 1), pn1 = __shfl_down_sync(0xFFFFFFFFu, pix[1], 1);
            int ib0 = pix[0] * iw00 + pn0 * iw01 + pix[1] * iw10 + pn1 * iw11;  // Ib row -1 (image row ipy - 1)
            int pn2 = __shfl_down_sync(0xFFFFFFFFu, pix[2], 1);
            int ib1 = pix[1] * iw00 + pn1 * iw01 + pix[2] * iw10 + pn2 * iw11;  // Ib row 0
            int pnr = pn2;
#pragma unroll
            for (int y = 0; y < WIN; ++y) {
                const int pnn = __shfl_down_sync(0xFFFFFFFFu, pix[y + 3], 1);
                const int ib2 = pix[y + 2] * iw00 + pnr * iw01 + pix[y + 3] * iw10 + pnn * iw11;  // Ib row y + 1
                pnr = pnn;
                const int Vb = 3 * (ib0 + ib2) + 10 * ib1, Db = ib2 - ib0;
                const int ixn = __shfl_down_sync(0xFFFFFFFFu, Vb, 1) - __shfl_up_sync(0xFFFFFFFFu, Vb, 1);
                const int iyn = 3 * (__shfl_up_sync(0xFFFFFFFFu, Db, 1) + __shfl_down_sync(0xFFFFFFFFu, Db, 1)) + 10 * Db;
                const int ixv = descale(ixn, 14), iyv = descale(iyn, 14);
                Iw[y * 32] = (short)descale(ib1, 9);
                Gw[y * 32] = ((uint32_t)ixv & 0xFFFFu) | ((uint32_t)iyv << 16);
                a11 += ixv * ixv;
                a12 += ixv * iyv;
                a22 += iyv * iyv;
                ib0 = ib1;
                ib1 = ib2;
            }
        } else {
            const int X = ipx - 1 + min(lane, 23);
            const bool xin = X >= 0 && X < G.w;
            int pix[24];
            load_column<24>(I, G, X, ipy - 1, false, pix);
            int tp = 0, tpn = 0, tgx = 0, tgxn = 0, tgy = 0, tgyn = 0;  // the row above (top taps of the bilinear stencil)
#pragma unroll
            for (int r = 0; r < 22; ++r) {
                // Scharr gradient at (X, Y = ipy + r); zero outside the image like OpenCV's constant-0 padded derivative
                const int V = 3 * (pix[r] + pix[r + 2]) + 10 * pix[r + 1], Dv = pix[r + 2] - pix[r];
                const int Vl = __shfl_up_sync(0xFFFFFFFFu, V, 1), Vr = __shfl_down_sync(0xFFFFFFFFu, V, 1);
                const int Dl = __shfl_up_sync(0xFFFFFFFFu, Dv, 1), Dr = __shfl_down_sync(0xFFFFFFFFu, Dv, 1);
                const int Y = ipy + r;
                const bool in = xin && Y >= 0 && Y < G.h;
                const int gx = in ? Vr - Vl : 0;
                const int gy = in ? 3 * (Dl + Dr) + 10 * Dv : 0;
                const int p = pix[r + 1];
                const int pn = __shfl_down_sync(0xFFFFFFFFu, p, 1);
                const int gxn = __shfl_down_sync(0xFFFFFFFFu, gx, 1), gyn = __shfl_down_sync(0xFFFFFFFFu, gy, 1);
                if (r >= 1) {
                    const int ival = descale(tp * iw00 + tpn * iw01 + p * iw10 + pn * iw11, 9);
                    const int ixv = descale(tgx * iw00 + tgxn * iw01 + gx * iw10 + gxn * iw11, 14);
                    const int iyv = descale(tgy * iw00 + tgyn * iw01 + gy * iw10 + gyn * iw11, 14);
                    Iw[(r - 1) * 32] = (short)ival;
                    Gw[(r - 1) * 32] = ((uint32_t)ixv & 0xFFFFu) | ((uint32_t)iyv << 16);
                    a11 += ixv * ixv;
                    a12 += ixv * iyv;
                    a22 += iyv * iyv;
                }
                tp = p; tpn = pn; tgx = gx; tgxn = gxn; tgy = gy; tgyn = gyn;
            }
        }
        const long long sA11 = warp_sum64(wlane ? a11 : 0), sA12 = warp_sum64(wlane ? a12 : 0), sA22 = warp_sum64(wlane ? a22 : 0);
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
        for (int j = 0; j < P.max_iter; ++j) {
            const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
            if (inx < -WIN || inx >= G.w || iny < -WIN || iny >= G.h) {
                if (level == 0) status = 0;
                break;
            }
            weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), iw00, iw01, iw10, iw11);
            int jp[22];
            load_column<22>(J, G, inx - 1 + min(lane, 23), iny, inx - 1 >= 0 && inx + 22 < G.w && iny >= 0 && iny + 21 < G.h, jp);
            // neighbour column of every row first (22 independent shuffles), then two accumulator pairs: short dependency chains
            int jn[22];
#pragma unroll
            for (int r = 0; r < 22; ++r) jn[r] = __shfl_down_sync(0xFFFFFFFFu, jp[r], 1);
            int b1 = 0, b2 = 0, b1o = 0, b2o = 0;
#pragma unroll
            for (int y = 0; y < WIN; ++y) {
                const int diff = descale(jp[y] * iw00 + jn[y] * iw01 + jp[y + 1] * iw10 + jn[y + 1] * iw11, 9) - Iw[y * 32];
                const uint32_t gw = Gw[y * 32];
                const int gxv = (int)(short)(gw & 0xFFFFu), gyv = (int)gw >> 16;
                if (y & 1) {
                    b1o += diff * gxv;
                    b2o += diff * gyv;
                } else {
                    b1 += diff * gxv;
                    b2 += diff * gyv;
                }
            }
            b1 += b1o;
            b2 += b2o;
            const long long sb1 = warp_sum64(wlane ? b1 : 0), sb2 = warp_sum64(wlane ? b2 : 0);
            const float b1f = __fmul_rn((float)sb1, FLT_SCALE), b2f = __fmul_rn((float)sb2, FLT_SCALE);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2f), __fmul_rn(A22, b1f)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1f), __fmul_rn(A11, b2f)), D);
            nx = __fadd_rn(nx, dx);
            ny = __fadd_rn(ny, dy);
            outx = __fadd_rn(nx, half);
            outy = __fadd_rn(ny, half);
            if ((double)dx * (double)dx + (double)dy * (double)dy <= P.eps2) break;
            if (j > 0 && (double)fabsf(__fadd_rn(dx, pdx)) < 0.01 && (double)fabsf(__fadd_rn(dy, pdy)) < 0.01) {
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
                const uint8_t* __restrict__ col = J + reflect101(inx - 1 + min(lane, 23), G.w);
                int e = 0;
                int top = __ldg(col + (size_t)reflect101(iny, G.h) * G.pitch), topn = __shfl_down_sync(0xFFFFFFFFu, top, 1);
#pragma unroll
                for (int y = 0; y < WIN; ++y) {
                    const int bot = __ldg(col + (size_t)reflect101(iny + y + 1, G.h) * G.pitch), botn = __shfl_down_sync(0xFFFFFFFFu, bot, 1);
                    e += abs(descale(top * iw00 + topn * iw01 + bot * iw10 + botn * iw11, 9) - Iw[y * 32]);
                    top = bot;
                    topn = botn;
                }
                const long long se = warp_sum64(wlane ? e : 0);
                err = __fdiv_rn((float)se, (float)(32 * WIN * WIN));
            }
        }
    }
    if (lane == 0) {
        dst[2 * pt] = nx;
        dst[2 * pt + 1] = ny;
        st_out[pt] = (uint8_t)status;
        if (P.want_err && lk_err) lk_err[pt] = status ? err : 0.f;
    }
}

}  // namespace

int launch_lk_pyramid(svo_ctx* c, const SlotMap& map, int n_img) {
    StageTimer t(c, 7);
    for (int l = 1; l < c->lk_levels; ++l) {
        const LkLevelGeom& S = c->lk[l - 1];
        const LkLevelGeom& D = c->lk[l];
        dim3 grid((D.w + 255) / 256, (D.h + 7) / 8, n_img);
        k_pyr_down<<<grid, 256, 0, c->stream>>>(S, D, map);
        ++c->launches;
    }
    SVO_CUDA_TRY(c, cudaGetLastError());
    return SVO_OK;
}

static void fill_params(svo_ctx* c, LkParams& P, int max_level, int max_iter, double eps, double min_eig) {
    for (int l = 0; l < SVO_LK_LEVELS; ++l) P.lv[l] = c->lk[l];
    P.levels = std::min(c->lk_levels, max_level + 1);
    P.max_iter = std::min(std::max(max_iter, 0), 100);
    eps = std::min(std::max(eps, 0.), 10.);
    P.eps2 = eps * eps;
    P.min_eig = (float)min_eig;
    P.pt_cap = c->pt_cap;
}

int launch_lk_track(svo_ctx* c, int first, int ring, int n_frames, int skip_first, const LkCall& call, int max_iter, double eps,
                    double min_eig) {
    StageTimer t(c, 8 + call.status_row);
    LkParams P{};
    fill_params(c, P, SVO_LK_LEVELS - 1, max_iter, eps, min_eig);
    P.first = first;
    P.ring = ring;
    P.skip_first = skip_first;
    P.call = call;
    P.standalone = 0;
    P.want_err = 0;
    dim3 grid((c->pt_cap + LW - 1) / LW, n_frames);
    k_lk_track<<<grid, LW * 32, 0, c->stream>>>(P, c->d_fast_kps, c->d_nfast, c->d_lk_pts, c->d_lk_status, nullptr);
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
    P.call = LkCall{0, 0, 1, 0, 0, 1, 0};
    P.standalone = 1;
    P.n_fixed = n;
    P.want_err = want_err;
    dim3 grid((n + LW - 1) / LW, 1);
    k_lk_track<<<grid, LW * 32, 0, c->stream>>>(P, c->d_fast_kps, c->d_nfast, c->d_lk_pts, c->d_lk_status, c->d_lk_err);
    ++c->launches;
    SVO_CUDA_TRY(c, cudaGetLastError());
    return SVO_OK;
}
