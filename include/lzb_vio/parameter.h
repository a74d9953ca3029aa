// include/lzb_vio/parameter.h -- the fields of the reference's Parameter (include/lzb_vio/parameter.h:20-57) that reach
// the front-end (src/parameter.cpp:53-70).  The YAML reader (cv::FileStorage, config.cpp) stays host-side and unchanged.
#pragma once
#include <memory>
#include <string>

namespace lzb_vio {

struct Parameter {
    typedef std::shared_ptr<Parameter> Ptr;
    std::string track_mode_ = "ORB_stereof2f_pnp";  // config/default.yaml:75
    double feature_match_error_ = 3;                 // config/default.yaml:66
    int num_features_tracking_ = 50;
    double inlier_rate_ = 0.01;         // config/default.yaml:77
    int iterationsCount_ = 500;         // :80
    float reprojectionError_ = 0.5f;    // :81
    float confidence_ = 0.99f;          // :82
    double minmove_ = 0.05;             // :87
    double maxmove_ = 10;               // :88
    int nFeatures_ = 2000;        // config/default.yaml:93
    float fScaleFactor_ = 1.2f;   // :92
    int nLevels_ = 8;             // :91
    int fIniThFAST_ = 20;         // :90
    int fMinThFAST_ = 7;          // :89
};

}  // namespace lzb_vio
