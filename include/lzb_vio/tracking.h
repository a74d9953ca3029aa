// include/lzb_vio/tracking.h -- the front-end half of the reference's Tracking (include/lzb_vio/tracking.h:29-153).
//
// Same method names and argument order as the reference for every method on the hot path:
//   Detect_MyORBFeatures / Detect_OpenCVFASTFeatures                      (tracking.h:88-89,  tracking.cpp:94-113, 502-532)
//   ORB_Robust_Find_MuliImage_MatchedFeatures                             (tracking.h:57-59,  tracking.cpp:534-581)
//   LK_Robust_Find_MuliImage_MatchedFeatures / deleteBadmatchFeatures     (tracking.h:52-56, 91-95, tracking.cpp:583-660)
//   OpenCV_EstimatePose_PnP                                               (tracking.h:76-80,  tracking.cpp:464-501)
// and the AddFrame / StereoInit_f2f / Track state machine around them (tracking.cpp:49-128).  Triangulation and the pose
// estimate are device calls too (TriangulateMatched = the cv::triangulatePoints + convertPointsFromHomogeneous pair at
// tracking.cpp:190-192; OpenCV_EstimatePose_PnP with the reference's own signature).  What follows them (pose gating, frame_pose_
// bookkeeping, display: tracking.cpp:200-249, 300-344) is the reference's own host code and is reached through pose_backend_,
// which the application sets.
#pragma once
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "lzb_vio/ORBextractor.h"
#include "lzb_vio/feature.h"
#include "lzb_vio/frame.h"
#include "lzb_vio/parameter.h"

namespace lzb_vio {

class System;
class Sensors;

enum class TrackingStatus { INITING, TRACKING_GOOD, LOST };

class Tracking {
public:
    typedef std::shared_ptr<Tracking> Ptr;
    // host pose back-end: (matched_t1_left, matched_t1_right, matched_t2_left) -> pose accepted?
    typedef std::function<bool(const std::vector<cv::Point2f>&, const std::vector<cv::Point2f>&, const std::vector<cv::Point2f>&)> PoseBackend;

    Tracking(System* system, Parameter::Ptr parameter, std::shared_ptr<Sensors> sensors);
    ~Tracking();
    void Set_vo(System* vo) { system_ = vo; }
    bool AddFrame(Frame::Ptr frame);
    TrackingStatus GetStatus() const { return status_; }

    PoseBackend pose_backend_;  // unset: the front-end result alone decides AddFrame's return value

    // the methods below are private in the reference; they are public here so that the parity tests can drive them
    bool StereoInit_f2f();
    bool Track();
    bool LK_StereoF2F_PnP_Track();
    bool ORB_StereoF2F_PnP_Track();
    int LK_Robust_Find_MuliImage_MatchedFeatures(std::vector<cv::Point2f>& points_t2_left, std::vector<cv::Point2f>& points_t2_right,
                                                 std::vector<cv::Point2f>& points_t1_left, std::vector<cv::Point2f>& points_t1_right);
    // positional semantics of the .cpp (quirk Q4): (t2_left, t1_left, t1_right)
    int ORB_Robust_Find_MuliImage_MatchedFeatures(std::vector<cv::Point2f>& points_t2_left, std::vector<cv::Point2f>& points_t1_left,
                                                  std::vector<cv::Point2f>& points_t1_right);
    int Detect_MyORBFeatures();
    int Detect_OpenCVFASTFeatures();
    // cv::triangulatePoints(projMatrl, projMatrr, t1_left, t1_right, points4D) + cv::convertPointsFromHomogeneous(points4D.t(),
    // points3D) (tracking.cpp:190-192, 292-294): projection matrices 3 x 4 CV_64F, result n x 1 CV_32FC3
    void TriangulateMatched(cv::Mat& projMatrl, cv::Mat& projMatrr, std::vector<cv::Point2f>& points_t1_left,
                            std::vector<cv::Point2f>& points_t1_right, cv::Mat& points3D_t1);
    // same signature and outputs as the reference (tracking.cpp:464-501): rotation 3 x 3 CV_64F, translation 3 x 1 CV_64F, the
    // return value is the inlier-rate test; last_inliers_ keeps solvePnPRansac's `inliers` (indices)
    bool OpenCV_EstimatePose_PnP(cv::Mat& projMatrl, std::vector<cv::Point2f>& pointsLeft_t2, cv::Mat& points3D_t1, cv::Mat& rotation,
                                 cv::Mat& translation);
    std::vector<int> last_inliers_;
    // Sensors::projMatr1_ / projMatr2_ (sensors.cpp:8-24; 3 x 4 CV_64F).  The reference reads them from sensors_; Sensors (YAML /
    // calibration files) is outside this library, so the application hands the two matrices over once.  With them set and no
    // pose_backend_, *_StereoF2F_PnP_Track run the reference's whole tail except the display: triangulation, OpenCV_EstimatePose_PnP,
    // the Euler-angle / step-length gates and frame_pose_ = frame_pose_ * [R t; 0 1]^-1 (tracking.cpp:186-246, 288-341).
    void SetProjection(const cv::Mat& projMatrl, const cv::Mat& projMatrr);
    cv::Vec3f rotationMatrixToEulerAngles(cv::Mat& R);  // tracking.cpp:440-463
    bool EstimateAndAccumulatePose(std::vector<cv::Point2f>& matched_t1_left, std::vector<cv::Point2f>& matched_t1_right,
                                   std::vector<cv::Point2f>& matched_t2_left, bool lk_mode);
    cv::Mat frame_pose_;  // 4 x 4 CV_64F, identity at start (tracking.h:117)
    double Px_ = 0, Py_ = 0, Pz_ = 0;
    void deleteBadmatchFeatures(std::vector<cv::Point2f>& points0, std::vector<cv::Point2f>& points1, std::vector<cv::Point2f>& points2,
                                std::vector<cv::Point2f>& points3, std::vector<cv::Point2f>& points0_return,
                                std::vector<unsigned char>& status0, std::vector<unsigned char>& status1,
                                std::vector<unsigned char>& status2, std::vector<unsigned char>& status3);

    Frame::Ptr current_frame_ = nullptr;
    Frame::Ptr last_frame_ = nullptr;
    ORBextractor* mpORBextractorLeft_ = nullptr;
    // results of the last association, for inspection
    std::vector<cv::DMatch> last_match1_, last_match2_;

private:
    void Readparameter();
    TrackingStatus status_ = TrackingStatus::INITING;
    System* system_ = nullptr;
    Parameter::Ptr parameter_;
    std::shared_ptr<Sensors> sensors_;
    std::string track_mode_;
    double feature_match_error_ = 10;
    int num_features_tracking_ = 50;
    double inlier_rate_ = 0.01;
    int iterationsCount_ = 500;
    float reprojectionError_ = 0.5f, confidence_ = 0.99f;
    double minmove_ = 0.05, maxmove_ = 10;
    cv::Mat projMatrl_, projMatrr_;
    int nFeatures_ = 2000, nLevels_ = 8, fIniThFAST_ = 20, fMinThFAST_ = 7;
    float fScaleFactor_ = 1.2f;
};

}  // namespace lzb_vio
