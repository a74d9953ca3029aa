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
#include <algorithm>

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

// ---- standalone forms (the drop-in lzb_vio::Tracking calls these, so the per-frame path and the batch path share K10 / K11) ----
// The kernels address a ring of frame slots; here a two-slot ring is laid out in the scratch buffer: slot 0 = the last frame,
// slot 1 = the current frame, right-hand images at slot 2 + s.
extern "C" int svo_orb_match_filter(svo_ctx* c, const svo_keypoint* last_left, int n_last_left, const svo_keypoint* last_right, int n_last_right,
                                    const svo_keypoint* cur_left, int n_cur_left, const int32_t* idx1, const int32_t* dist1, const int32_t* idx2,
                                    const int32_t* dist2, double feature_match_error, float* t2_left, float* t1_left, float* t1_right, int cap,
                                    int* n_out) {
    if (!c) return SVO_ERR_INVALID;
    if (!n_out) return c->fail(SVO_ERR_INVALID, "svo_orb_match_filter: n_out is null");
    *n_out = 0;
    if (n_last_left < 0 || n_last_right < 0 || n_cur_left < 0 || cap < 0) return c->fail(SVO_ERR_INVALID, "svo_orb_match_filter: negative size");
    const int des = std::min(std::min(n_last_left, n_last_right), n_cur_left);  // quirk Q3: only the first des_index queries are looked at
    if (des == 0) return SVO_OK;
    if (!last_left || !last_right || !cur_left || !idx1 || !dist1 || !idx2 || !dist2 || !t2_left || !t1_left || !t1_right)
        return c->fail(SVO_ERR_INVALID, "svo_orb_match_filter: null array");
    for (int i = 0; i < des; ++i)
        if (idx1[i] < 0 || idx1[i] >= n_last_right || idx2[i] < 0 || idx2[i] >= n_cur_left)
            return c->fail(SVO_ERR_INVALID, "svo_orb_match_filter: match index out of range");
    cudaSetDevice(c->device);
    const int K = std::max(std::max(n_last_left, n_last_right), n_cur_left);
    const size_t kb = svo_align((size_t)K * sizeof(svo_keypoint), 16), mb = svo_align((size_t)K * 4, 16), ob = svo_align((size_t)K * 8, 16);
    int rc = svo_ensure_scratch(c, 4 * kb + 8 * mb + 6 * ob + 64);
    if (rc != SVO_OK) return rc;
    uint8_t* base = (uint8_t*)c->d_scratch;
    svo_keypoint* d_kps = (svo_keypoint*)base;                 // [4][K]: last left, cur left, last right, -
    int* d_match = (int*)(base + 4 * kb);                       // [2 slots][4][K]; slot 1 is used
    float* d_out = (float*)(base + 4 * kb + 8 * mb);            // [2 slots][3][K][2]
    int* d_cnt = (int*)(base + 4 * kb + 8 * mb + 6 * ob);       // nkp[4], out_n[2]
    const int Kr = K;  // row length of every array below (the kernel's kp_cap)
    svo_keypoint* rows_k[3] = {d_kps, d_kps + Kr, d_kps + 2 * (size_t)Kr};
    SVO_CUDA_TRY(c, cudaMemcpyAsync(rows_k[0], last_left, sizeof(svo_keypoint) * n_last_left, cudaMemcpyHostToDevice, c->stream));
    SVO_CUDA_TRY(c, cudaMemcpyAsync(rows_k[1], cur_left, sizeof(svo_keypoint) * n_cur_left, cudaMemcpyHostToDevice, c->stream));
    SVO_CUDA_TRY(c, cudaMemcpyAsync(rows_k[2], last_right, sizeof(svo_keypoint) * n_last_right, cudaMemcpyHostToDevice, c->stream));
    int* m1 = d_match + (size_t)4 * Kr;  // slot 1
    const int32_t* src[4] = {idx1, dist1, idx2, dist2};
    for (int r = 0; r < 4; ++r) SVO_CUDA_TRY(c, cudaMemcpyAsync(m1 + (size_t)r * Kr, src[r], sizeof(int) * des, cudaMemcpyHostToDevice, c->stream));
    const int h_cnt[6] = {n_last_left, n_cur_left, n_last_right, 0, 0, 0};
    SVO_CUDA_TRY(c, cudaMemcpyAsync(d_cnt, h_cnt, sizeof h_cnt, cudaMemcpyHostToDevice, c->stream));
    k_orb_filter<<<1, NT, 0, c->stream>>>(d_kps, d_cnt, d_match, Kr, 1, 2, 0, feature_match_error, d_out, d_cnt + 4);
    ++c->launches;
    SVO_CUDA_TRY(c, cudaGetLastError());
    int n = 0;
    SVO_CUDA_TRY(c, cudaMemcpyAsync(&n, d_cnt + 5, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    *n_out = n;
    if (n > cap) return c->fail(SVO_ERR_CAPACITY, "svo_orb_match_filter: more accepted matches than cap");
    if (n > 0) {
        const float* o1 = d_out + (size_t)3 * Kr * 2;  // slot 1: t2_left, t1_left, t1_right
        SVO_CUDA_TRY(c, cudaMemcpyAsync(t2_left, o1, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, c->stream));
        SVO_CUDA_TRY(c, cudaMemcpyAsync(t1_left, o1 + (size_t)Kr * 2, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, c->stream));
        SVO_CUDA_TRY(c, cudaMemcpyAsync(t1_right, o1 + (size_t)Kr * 4, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, c->stream));
        SVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
    return SVO_OK;
}

extern "C" int svo_lk_chain_filter(svo_ctx* c, int n, float* p0, float* p1, float* p2, float* p3, float* p0r, const uint8_t* s0, const uint8_t* s1,
                                   const uint8_t* s2, const uint8_t* s3, double feature_match_error, int* n_out) {
    if (!c) return SVO_ERR_INVALID;
    if (!n_out) return c->fail(SVO_ERR_INVALID, "svo_lk_chain_filter: n_out is null");
    *n_out = 0;
    if (n < 0) return c->fail(SVO_ERR_INVALID, "svo_lk_chain_filter: negative size");
    if (n == 0) return SVO_OK;
    if (!p0 || !p1 || !p2 || !p3 || !p0r || !s0 || !s1 || !s2 || !s3) return c->fail(SVO_ERR_INVALID, "svo_lk_chain_filter: null array");
    cudaSetDevice(c->device);
    const int P = n;
    const size_t kb = svo_align((size_t)P * sizeof(svo_keypoint), 16), lb = svo_align((size_t)5 * P * 8, 16), sb = svo_align((size_t)4 * P, 16);
    int rc = svo_ensure_scratch(c, 2 * kb + 4 * lb + 2 * sb + 64);
    if (rc != SVO_OK) return rc;
    uint8_t* base = (uint8_t*)c->d_scratch;
    svo_keypoint* d_seeds = (svo_keypoint*)base;              // slot 0 = the last frame's seeds
    float* d_pts = (float*)(base + 2 * kb);                   // [2 slots][5][P][2]; slot 1 is used (list 0 unused)
    float* d_out = (float*)(base + 2 * kb + 2 * lb);          // [2 slots][5][P][2]
    uint8_t* d_st = base + 2 * kb + 4 * lb;                   // [2 slots][4][P]
    int* d_cnt = (int*)(base + 2 * kb + 4 * lb + 2 * sb);     // nfast[2], out_n[2]
    std::vector<svo_keypoint> seeds((size_t)n);
    for (int i = 0; i < n; ++i) seeds[i] = svo_keypoint{p0[2 * i], p0[2 * i + 1], 7.f, -1.f, 0.f, 0, -1};
    SVO_CUDA_TRY(c, cudaMemcpyAsync(d_seeds, seeds.data(), sizeof(svo_keypoint) * n, cudaMemcpyHostToDevice, c->stream));
    float* L1 = d_pts + (size_t)5 * P * 2;
    const float* lists[4] = {p1, p2, p3, p0r};
    for (int k = 0; k < 4; ++k) SVO_CUDA_TRY(c, cudaMemcpyAsync(L1 + (size_t)(k + 1) * P * 2, lists[k], sizeof(float) * 2 * n, cudaMemcpyHostToDevice, c->stream));
    const uint8_t* st[4] = {s0, s1, s2, s3};
    for (int k = 0; k < 4; ++k) SVO_CUDA_TRY(c, cudaMemcpyAsync(d_st + (size_t)4 * P + (size_t)k * P, st[k], (size_t)n, cudaMemcpyHostToDevice, c->stream));
    const int h_cnt[4] = {n, 0, 0, 0};
    SVO_CUDA_TRY(c, cudaMemcpyAsync(d_cnt, h_cnt, sizeof h_cnt, cudaMemcpyHostToDevice, c->stream));
    k_lk_filter<<<1, NT, 0, c->stream>>>(d_seeds, d_cnt, d_pts, d_st, P, 1, 2, 0, feature_match_error, d_out, d_cnt + 2);
    ++c->launches;
    SVO_CUDA_TRY(c, cudaGetLastError());
    int kept = 0;
    SVO_CUDA_TRY(c, cudaMemcpyAsync(&kept, d_cnt + 3, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    *n_out = kept;
    if (kept > 0) {  // the five lists are compacted in place, like the reference's erase loop
        const float* o1 = d_out + (size_t)5 * P * 2;
        float* dst[5] = {p0, p1, p2, p3, p0r};
        for (int k = 0; k < 5; ++k) SVO_CUDA_TRY(c, cudaMemcpyAsync(dst[k], o1 + (size_t)k * P * 2, sizeof(float) * 2 * kept, cudaMemcpyDeviceToHost, c->stream));
        SVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
    return SVO_OK;
}
