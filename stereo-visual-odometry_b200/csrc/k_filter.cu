// k_filter.cu -- K10 / K11: the Tracking post-filters on the device (first "next" row).
//
//  K10 k_orb_filter   the tail of Tracking::ORB_Robust_Find_MuliImage_MatchedFeatures (tracking.cpp:546-577): des_index =
//                     min(rows of the three descriptor sets) (quirk Q3), min over i < des_index of max(d1, d2), accept query i
//                     if both distances <= max(2 * min_dist, 30) and |y_left - y_right| < feature_match_error; the three
//                     matched point lists are written in query order (what the reference push_backs).
//  K11 k_lk_filter    Tracking::deleteBadmatchFeatures (tracking.cpp:623-660): drop a track if any of the five points has a
//                     negative coordinate, any of the four status bytes is 0, or either stereo pair violates the epipolar
//                     band; survivors of the five lists keep their order (the reference erases in place).
// One CTA per unit frame, ordered compaction with ballot + block prefix; all comparisons in the reference's types (distances
// are integer-valued floats, the band test is float-abs against the double feature_match_error_).
#include "svo_common.cuh"

namespace {

constexpr int NT = 256;

// ordered append of the flagged items of one 256-wide chunk; returns the chunk total.  All NT threads call.
__device__ __forceinline__ int chunk_positions(bool flag, int* warp_sums, int& pos) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, flag);
    if (lane == 0) warp_sums[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) {
        const int v = warp_sums[w];
        if (w < warp) before += v;
        total += v;
    }
    pos = before + __popc(m & ((1u << lane) - 1));
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(NT) k_orb_filter(const svo_keypoint* __restrict__ kps, const int* __restrict__ nkp, const int* __restrict__ match,
                                                   int kp_cap, int first, int ring, int skip_first, double fme, float* __restrict__ out_pts,
                                                   int* __restrict__ out_n) {
    __shared__ int s_min[NT / 32];
    __shared__ int s_ws[NT / 32];
    const int f = blockIdx.x;
    const int cur_u = first + f, cur = cur_u >= ring ? cur_u - ring : cur_u, prev = cur == 0 ? ring - 1 : cur - 1;
    float* o = out_pts + (size_t)cur * 3 * kp_cap * 2;
    if (f == 0 && skip_first) {
        if (threadIdx.x == 0) out_n[cur] = 0;
        return;
    }
    const int nll = nkp[prev], nlr = nkp[ring + prev], ncl = nkp[cur];
    const int des_index = min(min(nll, nlr), ncl);
    const int* m = match + (size_t)cur * 4 * kp_cap;  // rows: stereo idx, stereo dist, temporal idx, temporal dist
    int mn = 1 << 30;
    for (int i = threadIdx.x; i < des_index; i += NT) mn = min(mn, max(m[kp_cap + i], m[3 * kp_cap + i]));
    mn = __reduce_min_sync(0xFFFFFFFFu, mn);
    if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = mn;
    __syncthreads();
    mn = s_min[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) mn = min(mn, s_min[w]);
    // min_dist starts at 10000 in the reference (:549); distances are at most 256
    const int lim = max(2 * min(mn, 10000), 30);
    const svo_keypoint* kl = kps + (size_t)prev * kp_cap;
    const svo_keypoint* kr = kps + (size_t)(ring + prev) * kp_cap;
    const svo_keypoint* kc = kps + (size_t)cur * kp_cap;
    int base = 0;
    for (int i0 = 0; i0 < des_index; i0 += NT) {
        const int i = i0 + threadIdx.x;
        bool ok = false;
        int ti = 0, tj = 0;
        if (i < des_index) {
            ti = m[i];
            tj = m[2 * kp_cap + i];
            ok = m[kp_cap + i] <= lim && m[3 * kp_cap + i] <= lim && (double)fabsf(kl[i].y - kr[ti].y) < fme;
        }
        int pos;
        const int total = chunk_positions(ok, s_ws, pos);
        if (ok) {
            const int p = base + pos;
            o[2 * p] = kc[tj].x;                      // points_t2_left
            o[2 * p + 1] = kc[tj].y;
            o[(size_t)kp_cap * 2 + 2 * p] = kl[i].x;  // points_t1_left
            o[(size_t)kp_cap * 2 + 2 * p + 1] = kl[i].y;
            o[(size_t)kp_cap * 4 + 2 * p] = kr[ti].x;  // points_t1_right
            o[(size_t)kp_cap * 4 + 2 * p + 1] = kr[ti].y;
        }
        base += total;
    }
    if (threadIdx.x == 0) out_n[cur] = base;
}

__global__ void __launch_bounds__(NT) k_lk_filter(const svo_keypoint* __restrict__ fast_kps, const int* __restrict__ nfast,
                                                  const float* __restrict__ lk_pts, const uint8_t* __restrict__ lk_status, int pt_cap, int first,
                                                  int ring, int skip_first, double fme, float* __restrict__ out_pts, int* __restrict__ out_n) {
    __shared__ int s_ws[NT / 32];
    const int f = blockIdx.x;
    const int cur_u = first + f, cur = cur_u >= ring ? cur_u - ring : cur_u, prev = cur == 0 ? ring - 1 : cur - 1;
    float* o = out_pts + (size_t)cur * 5 * pt_cap * 2;
    if (f == 0 && skip_first) {
        if (threadIdx.x == 0) out_n[cur] = 0;
        return;
    }
    const int n = nfast[prev];
    const svo_keypoint* seeds = fast_kps + (size_t)prev * pt_cap;
    const float* L = lk_pts + (size_t)cur * 5 * pt_cap * 2;  // lists 1..4: p1 (t1 right), p2 (t2 right), p3 (t2 left), p0r
    const uint8_t* S = lk_status + (size_t)cur * 4 * pt_cap;
    int base = 0;
    for (int i0 = 0; i0 < n; i0 += NT) {
        const int i = i0 + threadIdx.x;
        bool keep = false;
        float p[5][2];
        if (i < n) {
            p[0][0] = seeds[i].x;
            p[0][1] = seeds[i].y;
#pragma unroll
            for (int k = 1; k < 5; ++k) {
                p[k][0] = L[((size_t)k * pt_cap + i) * 2];
                p[k][1] = L[((size_t)k * pt_cap + i) * 2 + 1];
            }
            bool outside = false;
#pragma unroll
            for (int k = 0; k < 5; ++k) outside = outside || p[k][0] < 0 || p[k][1] < 0;
            const bool bad = !S[i] || !S[pt_cap + i] || !S[2 * pt_cap + i] || !S[3 * pt_cap + i];
            const bool epi = (double)fabsf(p[0][1] - p[1][1]) > fme || (double)fabsf(p[2][1] - p[3][1]) > fme;
            keep = !(outside || bad || epi);
        }
        int pos;
        const int total = chunk_positions(keep, s_ws, pos);
        if (keep) {
            const int q = base + pos;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                o[((size_t)k * pt_cap + q) * 2] = p[k][0];
                o[((size_t)k * pt_cap + q) * 2 + 1] = p[k][1];
            }
        }
        base += total;
    }
    if (threadIdx.x == 0) out_n[cur] = base;
}

}  // namespace

int launch_orb_filter(svo_ctx* c, int first, int ring, int n_frames, int skip_first) {
    k_orb_filter<<<n_frames, NT, 0, c->stream>>>(c->d_kps, c->d_nkp, c->d_match, c->kp_cap, first, ring, skip_first, c->fe_match_error,
                                                 c->d_orb_pts, c->d_norb);
    ++c->launches;
    SVO_CUDA_TRY(c, cudaGetLastError());
    return SVO_OK;
}

int launch_lk_filter(svo_ctx* c, int first, int ring, int n_frames, int skip_first) {
    k_lk_filter<<<n_frames, NT, 0, c->stream>>>(c->d_fast_kps, c->d_nfast, c->d_lk_pts, c->d_lk_status, c->pt_cap, first, ring, skip_first,
                                                c->fe_match_error, c->d_lkf_pts, c->d_nlkf);
    ++c->launches;
    SVO_CUDA_TRY(c, cudaGetLastError());
    return SVO_OK;
}
