// k_lk.cu -- K7 + K8: pyramidal Lucas-Kanade tracking == cv::calcOpticalFlowPyrLK as called at tracking.cpp:593-618
// (winSize 21x21, maxLevel 3, TermCriteria(COUNT+EPS, 30, 0.01), flags 0, minEigThreshold 1e-3); SURVEY.md A.9.
//
//  K7  k_pyr_down   cv::pyrDown: 5x5 [1 4 6 4 1]^2 / 256, BORDER_REFLECT_101, round half up; levels 1..3 of each image.
//                   Built once per image (the reference rebuilds both pyramids inside each of its 4 calls).
//  K8  k_lk_track   one warp per point, all pyramid levels in one launch.  Per level the warp stages the 24x24 patch of
//                   the previous image in shared memory (REFLECT_101 outside the image), derives the Scharr gradients
//                   on the fly (constant 0 outside the image, exactly like OpenCV's padded derivative buffer -- no
//                   derivative image is ever written to HBM), keeps the 21x21 interpolated window (I, Ix, Iy as int16)
//                   in registers -- lane r owns window row r -- and iterates Gauss-Newton on 22x22 patches of the next
//                   image.  All window sums are exact integers (int64 across the warp), converted to float once.
//
// f32 arithmetic uses explicit round-to-nearest intrinsics (no FMA contraction) to follow OpenCV's scalar expressions.
#include <cfloat>

#include "svo_common.cuh"

namespace {

constexpr int WIN = SVO_LK_WIN;  // 21
constexpr int LW = 8;            // warps (points) per CTA
constexpr int IP = 36;           // byte pitch of the staged image patches (9 words: conflict-free row-per-lane access)
constexpr int DP = 23;           // word pitch of the gradient patch

__device__ __forceinline__ int reflect101(int i, int n) {
    // |offset| <= WIN beyond the image and n > WIN on every level, so one reflection suffices
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

__global__ void __launch_bounds__(256) k_pyr_down(LkLevelGeom S, LkLevelGeom D, SlotMap map) {
    const int slot = map.slot(blockIdx.z);
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= D.w || y >= D.h) return;
    const uint8_t* __restrict__ src = S.base + (size_t)slot * S.plane;
    int cx[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) cx[j] = reflect101(2 * x + j - 2, S.w);
    int acc = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const uint8_t* r = src + (size_t)reflect101(2 * y + i - 2, S.h) * S.pitch;
        const int row = __ldg(r + cx[0]) + 4 * __ldg(r + cx[1]) + 6 * __ldg(r + cx[2]) + 4 * __ldg(r + cx[3]) + __ldg(r + cx[4]);
        const int kw = (i == 0 || i == 4) ? 1 : (i == 2 ? 6 : 4);
        acc += kw * row;
    }
    D.base[(size_t)slot * D.plane + (size_t)y * D.pitch + x] = (uint8_t)((acc + 128) >> 8);
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

__device__ __forceinline__ long long warp_sum64(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

__device__ __forceinline__ void weights(float a, float b, int& w00, int& w01, int& w10, int& w11) {
    const float one_a = __fsub_rn(1.f, a), one_b = __fsub_rn(1.f, b);
    w00 = __float2int_rn(__fmul_rn(__fmul_rn(one_a, one_b), 16384.f));
    w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, one_b), 16384.f));
    w10 = __float2int_rn(__fmul_rn(__fmul_rn(one_a, b), 16384.f));
    w11 = 16384 - w00 - w01 - w10;
}

// stage rows [y0, y0+rows) x cols [x0, x0+cols) of a level image (REFLECT_101 outside) into shared memory
__device__ __forceinline__ void stage_patch(const uint8_t* __restrict__ img, int w, int h, int pitch, int x0, int y0, int cols, int rows,
                                            uint8_t* dst, int lane) {
    if (lane < cols) {
        const int cx = reflect101(x0 + lane, w);
        for (int r = 0; r < rows; ++r) dst[r * IP + lane] = __ldg(img + (size_t)reflect101(y0 + r, h) * pitch + cx);
    }
}

__global__ void __launch_bounds__(LW * 32) k_lk_track(LkParams P, const svo_keypoint* __restrict__ fast_kps, const int* __restrict__ nfast,
                                                      float* __restrict__ lk_pts, uint8_t* __restrict__ lk_status, float* __restrict__ lk_err) {
    __shared__ __align__(16) uint8_t s_I[LW][24 * IP];
    __shared__ __align__(16) uint32_t s_D[LW][22 * DP];
    __shared__ __align__(16) uint8_t s_J[LW][22 * IP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.y;
    if (f == 0 && P.skip_first) return;
    const int cur = (P.first + f) % P.ring;
    const int pfs = (P.first + f - P.call.prev_back + P.ring) % P.ring, nfs = (P.first + f - P.call.next_back + P.ring) % P.ring;
    const int prev_slot = P.call.prev_side * P.ring + pfs, next_slot = P.call.next_side * P.ring + nfs;
    const int seed_fs = (P.first + f - 1 + P.ring) % P.ring;  // seeds: FAST corners of the previous left image
    const int n = P.standalone ? P.n_fixed : nfast[seed_fs];
    const int pt = blockIdx.x * LW + warp;
    if (pt >= n) return;  // warp-uniform

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

    uint8_t* sI = s_I[warp];
    uint32_t* sD = s_D[warp];
    uint8_t* sJ = s_J[warp];
    const float half = 10.f;  // (win - 1) * 0.5
    float nx = 0.f, ny = 0.f;  // nextPts[pt] carried across levels
    int status = 1;
    float err = 0.f;

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

        // previous-image patch: rows/cols [ip-1, ip+22]
        __syncwarp();
        stage_patch(I, G.w, G.h, G.pitch, ipx - 1, ipy - 1, 24, 24, sI, lane);
        __syncwarp();
        // Scharr gradients at the 22x22 positions (ipx+c, ipy+r); 0 outside the image
        if (lane < 22) {
            const int X = ipx + lane;
            const bool xin = X >= 0 && X < G.w;
            for (int r = 0; r < 22; ++r) {
                const int Y = ipy + r;
                uint32_t packed = 0;
                if (xin && Y >= 0 && Y < G.h) {
                    const uint8_t* p = sI + (r + 1) * IP + lane + 1;
                    const int a0 = p[-IP - 1], a1 = p[-IP], a2 = p[-IP + 1];
                    const int b0 = p[-1], b2 = p[1];
                    const int c0 = p[IP - 1], c1 = p[IP], c2 = p[IP + 1];
                    const int dxv = 3 * (a2 + c2) + 10 * b2 - 3 * (a0 + c0) - 10 * b0;
                    const int dyv = 3 * (c0 - a0) + 10 * (c1 - a1) + 3 * (c2 - a2);
                    packed = ((uint32_t)dxv & 0xFFFFu) | ((uint32_t)dyv << 16);
                }
                sD[r * DP + lane] = packed;
            }
        }
        __syncwarp();

        // interpolated window: lane r owns row r
        short Iw[WIN];
        uint32_t dIw[WIN];  // lo: Ix, hi: Iy
        long long sA11 = 0, sA12 = 0, sA22 = 0;
        if (lane < WIN) {
            const uint8_t* r0 = sI + (lane + 1) * IP + 1;
            const uint8_t* r1 = r0 + IP;
            const uint32_t* d0 = sD + lane * DP;
            const uint32_t* d1 = d0 + DP;
            int a11 = 0, a12 = 0, a22 = 0;
#pragma unroll
            for (int x = 0; x < WIN; ++x) {
                const int ival = descale(r0[x] * iw00 + r0[x + 1] * iw01 + r1[x] * iw10 + r1[x + 1] * iw11, 9);
                const uint32_t q00 = d0[x], q01 = d0[x + 1], q10 = d1[x], q11 = d1[x + 1];
                const int ixv = descale((short)(q00 & 0xFFFF) * iw00 + (short)(q01 & 0xFFFF) * iw01 + (short)(q10 & 0xFFFF) * iw10 +
                                            (short)(q11 & 0xFFFF) * iw11, 14);
                const int iyv = descale((short)(q00 >> 16) * iw00 + (short)(q01 >> 16) * iw01 + (short)(q10 >> 16) * iw10 +
                                            (short)(q11 >> 16) * iw11, 14);
                Iw[x] = (short)ival;
                dIw[x] = ((uint32_t)ixv & 0xFFFFu) | ((uint32_t)iyv << 16);
                a11 += ixv * ixv;
                a12 += ixv * iyv;
                a22 += iyv * iyv;
            }
            sA11 = a11;
            sA12 = a12;
            sA22 = a22;
        } else {
#pragma unroll
            for (int x = 0; x < WIN; ++x) {
                Iw[x] = 0;
                dIw[x] = 0;
            }
        }
        sA11 = warp_sum64(sA11);
        sA12 = warp_sum64(sA12);
        sA22 = warp_sum64(sA22);
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
            __syncwarp();
            stage_patch(J, G.w, G.h, G.pitch, inx, iny, 22, 22, sJ, lane);
            __syncwarp();
            long long sb1 = 0, sb2 = 0;
            if (lane < WIN) {
                const uint8_t* r0 = sJ + lane * IP;
                const uint8_t* r1 = r0 + IP;
                int b1 = 0, b2 = 0;
#pragma unroll
                for (int x = 0; x < WIN; ++x) {
                    const int diff = descale(r0[x] * iw00 + r0[x + 1] * iw01 + r1[x] * iw10 + r1[x + 1] * iw11, 9) - Iw[x];
                    b1 += diff * (short)(dIw[x] & 0xFFFF);
                    b2 += diff * (short)(dIw[x] >> 16);
                }
                sb1 = b1;
                sb2 = b2;
            }
            sb1 = warp_sum64(sb1);
            sb2 = warp_sum64(sb2);
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
                __syncwarp();
                stage_patch(J, G.w, G.h, G.pitch, inx, iny, 22, 22, sJ, lane);
                __syncwarp();
                long long se = 0;
                if (lane < WIN) {
                    const uint8_t* r0 = sJ + lane * IP;
                    const uint8_t* r1 = r0 + IP;
                    int e = 0;
#pragma unroll
                    for (int x = 0; x < WIN; ++x) e += abs(descale(r0[x] * iw00 + r0[x + 1] * iw01 + r1[x] * iw10 + r1[x + 1] * iw11, 9) - Iw[x]);
                    se = e;
                }
                se = warp_sum64(se);
                err = __fdiv_rn(__fmul_rn((float)se, 1.f), (float)(32 * WIN * WIN));
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
        dim3 grid((D.w + 63) / 64, (D.h + 3) / 4, n_img);
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
