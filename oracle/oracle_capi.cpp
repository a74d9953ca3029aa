// oracle/oracle_capi.cpp -- CPU ORACLE (test infrastructure only, see oracle.h).  extern "C" surface for ctypes.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>

#include "oracle.h"

using namespace orc;

extern "C" {

void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstride, uint8_t* dst, int dw, int dh, size_t dstride) {
    resize_linear_u8(Image{src, sw, sh, sstride}, dst, dw, dh, dstride);
}

int orc_fast9_detect(const uint8_t* img, int w, int h, size_t stride, int threshold, int nms, KeyPoint* out, int cap) {
    std::vector<KeyPoint> v;
    fast9_detect(Image{img, w, h, stride}, threshold, nms != 0, v);
    const int n = (int)std::min<size_t>(v.size(), (size_t)cap);
    if (n) std::memcpy(out, v.data(), sizeof(KeyPoint) * n);
    return (int)v.size();
}

// score map with the 3-px frame zeroed: S(p) if S(p) >= min_score else 0 (tests of the CUDA score stage)
void orc_fast9_score_map(const uint8_t* img, int w, int h, size_t stride, int min_score, uint8_t* out) {
    std::memset(out, 0, (size_t)w * h);
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            int s = fast9_score(img + (size_t)y * stride + x, stride);
            out[(size_t)y * w + x] = (uint8_t)(s >= min_score ? s : 0);
        }
}

void orc_gaussian_blur7(const uint8_t* src, int w, int h, size_t stride, uint8_t* dst, size_t dstride) {
    gaussian_blur7_sigma2(Image{src, w, h, stride}, dst, dstride);
}

float orc_fast_atan2(float y, float x) { return fast_atan2(y, x); }
void orc_fast_atan2_array(const float* y, const float* x, float* out, int n) {
    for (int i = 0; i < n; ++i) out[i] = fast_atan2(y[i], x[i]);
}

void orc_match_hamming(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx, int32_t* dist) {
    match_hamming(q, nq, t, nt, idx, dist);
}

void orc_pyr_down(const uint8_t* src, int w, int h, size_t stride, uint8_t* dst, size_t dstride) {
    pyr_down(Image{src, w, h, stride}, dst, dstride);
}

void orc_lk_track(const uint8_t* prev, const uint8_t* next, int w, int h, size_t stride, const float* prev_xy, int n,
                  float* next_xy, uint8_t* status, float* err, int win, int max_level, int max_iter, double eps, double min_eig) {
    calc_optical_flow_pyr_lk(Image{prev, w, h, stride}, Image{next, w, h, stride}, prev_xy, n, next_xy, status, err, win,
                             max_level, max_iter, eps, min_eig);
}

void orc_orb_tables(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, float* scale, float* inv_scale,
                    int* features_per_level, int* umax16) {
    OrbTables t = make_tables(OrbParams{nfeatures, scaleFactor, nlevels, iniTh, minTh});
    for (int i = 0; i < nlevels; ++i) {
        scale[i] = t.scale[i];
        inv_scale[i] = t.inv_scale[i];
        features_per_level[i] = t.features_per_level[i];
    }
    for (int i = 0; i < 16; ++i) umax16[i] = t.umax[i];
}

const int8_t* orc_brief_pattern() { return brief_pattern(); }

int orc_distribute_octree(const KeyPoint* keys, int n, int minX, int maxX, int minY, int maxY, int N, KeyPoint* out, int cap,
                          int* stats2) {
    std::vector<KeyPoint> k(keys, keys + n);
    OctreeStats st;
    std::vector<KeyPoint> r = distribute_octree(k, minX, maxX, minY, maxY, N, &st);
    if (stats2) { stats2[0] = st.rounds; stats2[1] = st.break_in_tie_group; }
    const int m = (int)std::min<size_t>(r.size(), (size_t)cap);
    if (m) std::memcpy(out, r.data(), sizeof(KeyPoint) * m);
    return (int)r.size();
}

// tail of ORBextractor::operator() on caller-built intermediates (see orb_finish).  lv / bl: nlevels pointers to contiguous w[l] x h[l]
// images; cand: all levels' candidates concatenated, ncand[l] each.
int orc_orb_finish(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, const uint8_t* const* lv, const uint8_t* const* bl,
                   const int* w, const int* h, const KeyPoint* cand, const int* ncand, KeyPoint* kps, uint8_t* desc, int cap) {
    std::vector<Image> levels(nlevels), blurred(nlevels);
    std::vector<std::vector<KeyPoint>> c(nlevels);
    size_t off = 0;
    for (int l = 0; l < nlevels; ++l) {
        levels[l] = Image{lv[l], w[l], h[l], (size_t)w[l]};
        blurred[l] = Image{bl[l], w[l], h[l], (size_t)w[l]};
        c[l].assign(cand + off, cand + off + ncand[l]);
        off += (size_t)ncand[l];
    }
    std::vector<KeyPoint> out;
    std::vector<uint8_t> d;
    orb_finish(OrbParams{nfeatures, scaleFactor, nlevels, iniTh, minTh}, levels, blurred, c, out, d);
    const int n = (int)std::min<size_t>(out.size(), (size_t)cap);
    if (n) {
        std::memcpy(kps, out.data(), sizeof(KeyPoint) * n);
        std::memcpy(desc, d.data(), (size_t)32 * n);
    }
    return (int)out.size();
}

// ---- full extractor with intermediates kept in a handle ----
struct orc_orb_handle {
    OrbParams p;
    OrbResult r;
};

orc_orb_handle* orc_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
    orc_orb_handle* h = new orc_orb_handle();
    h->p = OrbParams{nfeatures, scaleFactor, nlevels, iniTh, minTh};
    return h;
}
void orc_orb_destroy(orc_orb_handle* h) { delete h; }

int orc_orb_extract(orc_orb_handle* h, const uint8_t* img, int w, int h_, size_t stride, KeyPoint* kps, uint8_t* desc, int cap) {
    orb_extract(h->p, Image{img, w, h_, stride}, h->r, true);
    const int n = (int)std::min<size_t>(h->r.keypoints.size(), (size_t)cap);
    if (n) {
        std::memcpy(kps, h->r.keypoints.data(), sizeof(KeyPoint) * n);
        std::memcpy(desc, h->r.descriptors.data(), (size_t)32 * n);
    }
    return (int)h->r.keypoints.size();
}
int orc_orb_level_size(orc_orb_handle* h, int level, int* w, int* hh) {
    if (level < 0 || level >= (int)h->r.pyramid.size()) return -1;
    *w = h->r.pyramid[level].w;
    *hh = h->r.pyramid[level].h;
    return 0;
}
int orc_orb_level_pixels(orc_orb_handle* h, int level, uint8_t* dst) {
    if (level < 0 || level >= (int)h->r.pyramid.size()) return -1;
    std::memcpy(dst, h->r.pyramid[level].buf.data(), h->r.pyramid[level].buf.size());
    return 0;
}
int orc_orb_level_candidates(orc_orb_handle* h, int level, KeyPoint* out, int cap) {
    if (level < 0 || level >= (int)h->r.candidates.size()) return -1;
    const auto& c = h->r.candidates[level];
    const int n = (int)std::min<size_t>(c.size(), (size_t)cap);
    if (n && out) std::memcpy(out, c.data(), sizeof(KeyPoint) * n);
    return (int)c.size();
}
int orc_orb_level_keypoints(orc_orb_handle* h, int level, KeyPoint* out, int cap) {
    if (level < 0 || level >= (int)h->r.level_kps.size()) return -1;
    const auto& c = h->r.level_kps[level];
    const int n = (int)std::min<size_t>(c.size(), (size_t)cap);
    if (n && out) std::memcpy(out, c.data(), sizeof(KeyPoint) * n);
    return (int)c.size();
}
int orc_orb_octree_stats(orc_orb_handle* h, int level, int* rounds, int* tie_break) {
    if (level < 0 || level >= (int)h->r.octree_stats.size()) return -1;
    *rounds = h->r.octree_stats[level].rounds;
    *tie_break = h->r.octree_stats[level].break_in_tie_group;
    return 0;
}

int orc_orb_match_filter(const int32_t* idx1, const int32_t* dist1, const int32_t* idx2, const int32_t* dist2, int nll, int nlr,
                         int ncl, const KeyPoint* ll, const KeyPoint* lr, double fme, int32_t* accepted) {
    return orb_match_filter(idx1, dist1, idx2, dist2, nll, nlr, ncl, ll, lr, fme, accepted);
}
int orc_lk_chain_filter(int n, const float* p0, const float* p1, const float* p2, const float* p3, const float* p0r,
                        const uint8_t* s0, const uint8_t* s1, const uint8_t* s2, const uint8_t* s3, double fme, uint8_t* keep) {
    return lk_chain_filter(n, p0, p1, p2, p3, p0r, s0, s1, s2, s3, fme, keep);
}

void orc_triangulate_points(const double* P1, const double* P2, const float* pts1, const float* pts2, int n, float* out4, float* out3) {
    triangulate_points(P1, P2, pts1, pts2, n, out4, out3);
}

// ---------------------------------------------------------------------------------------------------------------
// Whole front-end over a sequence, for the CPU baseline ("port"): per frame t: ORB on L_t and R_t, match(L_{t-1},R_{t-1}),
// match(L_{t-1},L_t); FAST on L_t and the 4-call LK chain over (t-1, t).  Frames are split over `threads` workers in
// contiguous chunks overlapping by one frame (same sharding rule as the GPU path).  mode bit0 = ORB, bit1 = LK, bit2 = the pose
// step on every filtered list (triangulatePoints + convertPointsFromHomogeneous + solvePnPRansac(500, 0.5 px, 0.99), KITTI
// sequence-00 projection matrices; tracking.cpp:190-200).
// Returns a checksum so the work cannot be optimised away.
// ---------------------------------------------------------------------------------------------------------------
// triangulate (t1_left, t1_right), then the pose from those 3-D points and their pixels in left_t
static uint64_t pose_step(const float* t1_left, const float* t1_right, const float* t2_left, int n) {
    static const double P1[12] = {718.856, 0, 607.1928, 0, 0, 718.856, 185.2157, 0, 0, 0, 1, 0};
    static const double P2[12] = {718.856, 0, 607.1928, -386.1448, 0, 718.856, 185.2157, 0, 0, 0, 1, 0};
    static const double K[9] = {718.856, 0, 607.1928, 0, 718.856, 185.2157, 0, 0, 1};
    std::vector<float> x4((size_t)4 * n), x3((size_t)3 * n);
    triangulate_points(P1, P2, t1_left, t1_right, n, x4.data(), x3.data());
    std::vector<uint8_t> mask((size_t)n);
    double out[15];
    int info[3];
    solve_pnp_ransac(K, x3.data(), t2_left, n, 500, 0.5f, (double)0.99f, mask.data(), out, info);
    return (uint64_t)info[0] * 31u + (uint64_t)info[2];
}

uint64_t orc_frontend_sequence(const uint8_t* left, const uint8_t* right, int n_frames, int w, int h, int nfeatures,
                               float scaleFactor, int nlevels, int iniTh, int minTh, int mode, int threads) {
    threads = std::max(1, std::min(threads, std::max(1, n_frames - 1)));
    std::atomic<uint64_t> checksum{0};
    const size_t plane = (size_t)w * h;
    auto worker = [&](int f0, int f1) {  // processes units t in (f0, f1]; frame f0 is the warm-up "last frame"
        OrbParams p{nfeatures, scaleFactor, nlevels, iniTh, minTh};
        OrbResult lastL, lastR, curL, curR;
        std::vector<KeyPoint> last_fast, cur_fast;
        uint64_t cs = 0;
        for (int t = f0; t <= f1; ++t) {
            Image L{left + plane * t, w, h, (size_t)w}, R{right + plane * t, w, h, (size_t)w};
            if (mode & 1) {
                orb_extract(p, L, curL, false);
                orb_extract(p, R, curR, false);
                if (t > f0) {
                    const int nq = (int)lastL.keypoints.size();
                    std::vector<int32_t> i1(nq), d1(nq), i2(nq), d2(nq), acc((size_t)3 * nq + 3);
                    match_hamming(lastL.descriptors.data(), nq, lastR.descriptors.data(), (int)lastR.keypoints.size(), i1.data(), d1.data());
                    match_hamming(lastL.descriptors.data(), nq, curL.descriptors.data(), (int)curL.keypoints.size(), i2.data(), d2.data());
                    int na = orb_match_filter(i1.data(), d1.data(), i2.data(), d2.data(), nq, (int)lastR.keypoints.size(),
                                              (int)curL.keypoints.size(), lastL.keypoints.data(), lastR.keypoints.data(), 3.0, acc.data());
                    cs += (uint64_t)na * 1000003u;
                    if ((mode & 4) && na > 0) {
                        std::vector<float> a1((size_t)2 * na), a2((size_t)2 * na), a3((size_t)2 * na);
                        for (int i = 0; i < na; ++i) {
                            const KeyPoint &kl = lastL.keypoints[acc[3 * i]], &kr = lastR.keypoints[acc[3 * i + 1]], &kc = curL.keypoints[acc[3 * i + 2]];
                            a1[2 * i] = kl.x, a1[2 * i + 1] = kl.y, a2[2 * i] = kr.x, a2[2 * i + 1] = kr.y, a3[2 * i] = kc.x, a3[2 * i + 1] = kc.y;
                        }
                        cs += pose_step(a1.data(), a2.data(), a3.data(), na);
                    }
                    for (int i = 0; i < nq; ++i) cs += (uint64_t)(i1[i] + 3 * i2[i] + d1[i]);
                }
                std::swap(lastL, curL);
                std::swap(lastR, curR);
            }
            if (mode & 2) {
                fast9_detect(L, 20, true, cur_fast);
                if (t > f0 && !last_fast.empty()) {
                    const int n = (int)last_fast.size();
                    std::vector<float> p0((size_t)2 * n), p1((size_t)2 * n), p2((size_t)2 * n), p3((size_t)2 * n), p0r((size_t)2 * n), err(n);
                    std::vector<uint8_t> s0(n), s1(n), s2(n), s3(n), keep(n);
                    for (int i = 0; i < n; ++i) { p0[2 * i] = last_fast[i].x; p0[2 * i + 1] = last_fast[i].y; }
                    Image pL{left + plane * (t - 1), w, h, (size_t)w}, pR{right + plane * (t - 1), w, h, (size_t)w};
                    calc_optical_flow_pyr_lk(pL, pR, p0.data(), n, p1.data(), s0.data(), err.data(), 21, 3, 30, 0.01, 0.001);
                    calc_optical_flow_pyr_lk(pR, R, p1.data(), n, p2.data(), s1.data(), err.data(), 21, 3, 30, 0.01, 0.001);
                    calc_optical_flow_pyr_lk(R, L, p2.data(), n, p3.data(), s2.data(), err.data(), 21, 3, 30, 0.01, 0.001);
                    calc_optical_flow_pyr_lk(L, pL, p3.data(), n, p0r.data(), s3.data(), err.data(), 21, 3, 30, 0.01, 0.001);
                    const int nk = lk_chain_filter(n, p0.data(), p1.data(), p2.data(), p3.data(), p0r.data(), s0.data(), s1.data(), s2.data(),
                                                   s3.data(), 3.0, keep.data());
                    cs += (uint64_t)nk * 7919u;
                    if ((mode & 4) && nk > 0) {
                        std::vector<float> a1((size_t)2 * nk), a2((size_t)2 * nk), a3((size_t)2 * nk);
                        for (int i = 0, j = 0; i < n; ++i)
                            if (keep[i]) {
                                a1[2 * j] = p0[2 * i], a1[2 * j + 1] = p0[2 * i + 1], a2[2 * j] = p1[2 * i], a2[2 * j + 1] = p1[2 * i + 1];
                                a3[2 * j] = p3[2 * i], a3[2 * j + 1] = p3[2 * i + 1];
                                ++j;
                            }
                        cs += pose_step(a1.data(), a2.data(), a3.data(), nk);
                    }
                }
                std::swap(last_fast, cur_fast);
            }
        }
        checksum += cs;
    };
    const int units = n_frames - 1;
    std::vector<std::thread> pool;
    for (int k = 0; k < threads; ++k) {
        const int u0 = (int)((int64_t)units * k / threads), u1 = (int)((int64_t)units * (k + 1) / threads);
        if (u1 > u0) pool.emplace_back(worker, u0, u1);
    }
    for (auto& th : pool) th.join();
    return checksum.load();
}

}  // extern "C"
