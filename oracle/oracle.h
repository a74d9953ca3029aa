// oracle/oracle.h -- CPU ORACLE for the stereo front-end. TEST INFRASTRUCTURE ONLY.
//
// This is a from-scratch CPU restatement of the hot path of liuzhenboo/Stereo-Visual-Odometry
// (ORBextractor::operator(), the two brute-force Hamming matches, the four-call pyramidal LK chain and the
// post-filters around them).  It is the checker for the CUDA path: only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may build, link, import or execute it.  Nothing under
// stereo-visual-odometry_b200/ links or loads it; the product path has no CPU fallback.
//
// PARITY PIN STATUS:  the reference ships no tests, fixtures or golden vectors and its build
// needs OpenCV C++ 3.x, Sophus, glog, g2o and Pangolin, none of which exist in this image.  Two pins replace them:
//  (1) every third-party (OpenCV) primitive restated here is checked bit-for-bit against cv2 4.13.0
//      (tests/test_oracle_cpu.py, fixtures under tests/golden/cv2_primitives.npz made by oracle/make_golden.py);
//  (2) the reference's OWN src/ORBextractor.cpp is compiled, unmodified and where it lies, against the header shim in
//      oracle/ref_shim/ (OpenCV containers written out, OpenCV algorithms forwarded to the primitives of (1)) into
//      oracle/_ref/liborb_ref.so (`make -C oracle ref`).  Under a monotone allocator the compiled reference and the
//      restatement in oracle_orb.cpp agree bit for bit -- keypoints in order with every field, descriptors, pyramid --
//      on every configuration of tests/test_reference_cpu.py, and the CUDA path agrees with both
//      (tests/test_orb_gpu.py).  tests/golden/ref_orbextractor.npz holds outputs of that library
//      (oracle/make_golden_ref.py).  Under the system malloc the octree's pair<int, ExtractorNode*> sort
//      (ORBextractor.cpp:636) makes ~1-3 % of the keypoints depend on the heap's history (measured, same test file):
//      parity of the SELECTED keypoints is therefore defined under the monotone allocator (see distribute_octree()).
//  (3) the three front-end member functions of src/tracking.cpp (:534-660: the ORB association filter, the four-call LK
//      chain, deleteBadmatchFeatures) are cut out of the reference file at build time into oracle/_ref/ and compiled inside
//      oracle/ref_tracking_capi.cpp (tracking.cpp as a whole needs Eigen / Sophus / g2o); orb_match_filter() and
//      lk_chain_filter() below equal that compiled code exactly (tests/test_reference_cpu.py).
//
// All citations "ORBextractor.cpp:N" / "tracking.cpp:N" are into /root/reference/src/.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace orc {

// == cv::KeyPoint layout (28 bytes): pt.x, pt.y, size, angle, response, octave, class_id
struct KeyPoint {
    float x, y, size, angle, response;
    int32_t octave, class_id;
};

struct Image {  // non-owning view, 8-bit single channel
    const uint8_t* data;
    int w, h;
    size_t stride;
};

struct OwnedImage {
    std::vector<uint8_t> buf;
    int w = 0, h = 0;
    Image view() const { return Image{buf.data(), w, h, (size_t)w}; }
    void create(int w_, int h_) { w = w_; h = h_; buf.assign((size_t)w_ * h_, 0); }
};

// ---- OpenCV primitive restatements ----
int cv_round(float v);    // cvRound: round-half-to-even
int cv_round(double v);
void resize_linear_u8(const Image& src, uint8_t* dst, int dw, int dh, size_t dstride);            // cv::resize INTER_LINEAR 8UC1
int fast9_score(const uint8_t* p, size_t stride);                                               // cornerScore<16> with threshold 0
void fast9_detect(const Image& roi, int threshold, bool nms, std::vector<KeyPoint>& out);       // cv::FAST(TYPE_9_16)
void gaussian_blur7_sigma2(const Image& src, uint8_t* dst, size_t dstride);                     // cv::GaussianBlur(7x7, 2, 2, REFLECT_101)
float fast_atan2(float y, float x);                                                             // cv::fastAtan2 (degrees)
void match_hamming(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* train_idx, int32_t* dist);  // BFMatcher(NORM_HAMMING).match
void pyr_down(const Image& src, uint8_t* dst, size_t dstride);                                  // cv::pyrDown, dst ((w+1)/2,(h+1)/2)
// cv::calcOpticalFlowPyrLK(prev, next, prevPts, nextPts, status, err, Size(win,win), maxLevel,
//                          TermCriteria(COUNT+EPS, maxIter, eps), flags=0, minEigThreshold)
void calc_optical_flow_pyr_lk(const Image& prev, const Image& next, const float* prev_xy, int n, float* next_xy,
                              uint8_t* status, float* err, int win, int max_level, int max_iter, double eps,
                              double min_eig);

// ---- reference restatement: ORBextractor ----
struct OrbParams {
    int nfeatures;
    float scaleFactor;
    int nlevels, iniThFAST, minThFAST;
};

struct OrbTables {  // ORBextractor::ORBextractor, ORBextractor.cpp:359-418
    std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
    std::vector<int> features_per_level;
    std::vector<int> umax;
};
OrbTables make_tables(const OrbParams& p);
const int8_t* brief_pattern();  // 256 x 4 (x0,y0,x1,y1), the standard ORB learned pattern

struct OctreeStats {
    int rounds = 0;
    int break_in_tie_group = 0;  // the final break of round B landed inside a group of equal-count nodes
};

// DistributeOctTree (ORBextractor.cpp:487-715); keys are (x,y,response) relative to minBorder.
std::vector<KeyPoint> distribute_octree(const std::vector<KeyPoint>& keys, int minX, int maxX, int minY, int maxY,
                                        int N, OctreeStats* stats = nullptr);

struct OrbResult {
    std::vector<OwnedImage> pyramid;                // levels without border
    std::vector<std::vector<KeyPoint>> candidates;  // per level, output of the cell loop (coords relative to minBorder)
    std::vector<std::vector<KeyPoint>> level_kps;   // per level after octree + angle (level coords)
    std::vector<KeyPoint> keypoints;                // final (scaled to level 0)
    std::vector<uint8_t> descriptors;               // N x 32
    std::vector<OctreeStats> octree_stats;
};
// ORBextractor::operator() (ORBextractor.cpp:990-1055)
void orb_extract(const OrbParams& p, const Image& img, OrbResult& out, bool keep_intermediates = true);

// tail of operator() (octree, IC_Angle, steered BRIEF, rescale) on externally built levels / blurs / candidates (CPU baseline variant
// with real-OpenCV image primitives, oracle/ref_pipeline.py)
void orb_finish(const OrbParams& p, const std::vector<Image>& levels, const std::vector<Image>& blurred,
                const std::vector<std::vector<KeyPoint>>& candidates, std::vector<KeyPoint>& keypoints, std::vector<uint8_t>& descriptors);

// ---- reference restatement: Tracking front-end ----
// Tracking::ORB_Robust_Find_MuliImage_MatchedFeatures (tracking.cpp:534-581). Returns #accepted; fills index triples.
int orb_match_filter(const int32_t* idx1, const int32_t* dist1, const int32_t* idx2, const int32_t* dist2, int n_last_left,
                     int n_last_right, int n_cur_left, const KeyPoint* last_left, const KeyPoint* last_right,
                     double feature_match_error, int32_t* accepted /* n x 3: (lastL idx, lastR idx, curL idx) */);
// Tracking::deleteBadmatchFeatures (tracking.cpp:623-660): keep[i]=1 if point i survives.
int lk_chain_filter(int n, const float* p0, const float* p1, const float* p2, const float* p3, const float* p0r,
                    const uint8_t* s0, const uint8_t* s1, const uint8_t* s2, const uint8_t* s3, double feature_match_error,
                    uint8_t* keep);

// cv::triangulatePoints + cv::convertPointsFromHomogeneous (tracking.cpp:190-192, 292-294).  P1, P2: 3x4 row-major; out4: n x 4
// (unit right singular vector of the smallest singular value, sign arbitrary), out3: n x 3 (float division like OpenCV).
void triangulate_points(const double* P1, const double* P2, const float* pts1, const float* pts2, int n, float* out4, float* out3);

// ---- pose from points (oracle_pnp.cpp; its header states what is pinned against cv2 4.13 and what cannot be) ----
// cv::solvePnP(SOLVEPNP_EPNP) on n >= 5 points; cv::solvePnP(SOLVEPNP_ITERATIVE, useExtrinsicGuess = true) on the points with
// mask != 0 (param = [rvec, tvec] in / out; returns LM iterations); the whole cv::solvePnPRansac call of tracking.cpp:486-488:
// out15 = rvec, tvec, R (row-major); info3 = inliers, winning hypothesis, hypotheses evaluated; returns solvePnPRansac's bool.
bool epnp(const double K[9], const float* pts3d, const float* pts2d, int n, double R[9], double t[3]);
int refine_lm(const double K[9], const float* p3, const float* p2, int n, const uint8_t* mask, double param[6]);
int solve_pnp_ransac(const double K[9], const float* p3, const float* p2, int n, int max_iters, float reproj_err, double confidence, uint8_t* mask,
                     double out15[15], int info3[3]);

}  // namespace orc
