// oracle/oracle_orb.cpp -- CPU ORACLE (test infrastructure only, see oracle.h).
// Restatement of the reference's own arithmetic in ORBextractor.cpp: scale/quota tables, the per-cell FAST grid
// with threshold fallback, the octree distribution, IC_Angle, the steered BRIEF descriptor and operator().
#include <algorithm>
#include <cmath>
#include <cstring>
#include <list>

#include "oracle.h"

namespace orc {

static const int8_t kBriefPattern[1024] = {
#include "brief_pattern.inc"
};
const int8_t* brief_pattern() { return kBriefPattern; }

static const int kPatchSize = 31, kHalfPatch = 15, kEdge = 19;  // ORBextractor.cpp:17-19

// ORBextractor::ORBextractor, ORBextractor.cpp:359-418
OrbTables make_tables(const OrbParams& p) {
    OrbTables t;
    const int n = p.nlevels;
    const double scaleFactorD = (double)p.scaleFactor;  // the member is a double initialised from the float arg
    t.scale.assign(n, 1.f);
    t.sigma2.assign(n, 1.f);
    for (int i = 1; i < n; ++i) {
        t.scale[i] = (float)(t.scale[i - 1] * scaleFactorD);
        t.sigma2[i] = t.scale[i] * t.scale[i];
    }
    t.inv_scale.resize(n);
    t.inv_sigma2.resize(n);
    for (int i = 0; i < n; ++i) {
        t.inv_scale[i] = 1.0f / t.scale[i];
        t.inv_sigma2[i] = 1.0f / t.sigma2[i];
    }
    t.features_per_level.assign(n, 0);
    const float factor = (float)(1.0f / scaleFactorD);
    float desired = p.nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)n));
    int sum = 0;
    for (int l = 0; l < n - 1; ++l) {
        t.features_per_level[l] = cv_round(desired);
        sum += t.features_per_level[l];
        desired *= factor;
    }
    t.features_per_level[n - 1] = std::max(p.nfeatures - sum, 0);

    t.umax.assign(kHalfPatch + 1, 0);
    const int vmax = (int)std::floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
    const int vmin = (int)std::ceil(kHalfPatch * std::sqrt(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (int v = 0; v <= vmax; ++v) t.umax[v] = cv_round(std::sqrt(hp2 - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (t.umax[v0] == t.umax[v0 + 1]) ++v0;
        t.umax[v] = v0;
        ++v0;
    }
    return t;
}

// ---------------------------------------------------------------------------------------------------------------
// Octree.  ExtractorNode::DivideNode (ORBextractor.cpp:430-485) + ORBextractor::DistributeOctTree (:487-715).
//
// CANONICALISATION: the reference sorts pair<int, ExtractorNode*> (:636), i.e. ties in
// key count are ordered by heap address, which is not a function of the input.  Here every node carries a
// creation stamp and ties are ordered by stamp -- the reference's behaviour under a monotone (bump) allocator.
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct Node {
    int x0, x1, y0, y1;     // UL.x, UR.x, UL.y, BL.y
    std::vector<int> keys;  // indices into the input, original order preserved
    bool final_ = false;    // bNoMore
    uint32_t stamp = 0;
    std::list<Node>::iterator self;
};

struct Quad {
    Node c[4];
};

void divide(const Node& n, const std::vector<KeyPoint>& keys, Quad& q) {
    const int halfX = (int)std::ceil(static_cast<float>(n.x1 - n.x0) / 2);
    const int halfY = (int)std::ceil(static_cast<float>(n.y1 - n.y0) / 2);
    const int xm = n.x0 + halfX, ym = n.y0 + halfY;
    q.c[0].x0 = n.x0; q.c[0].x1 = xm;   q.c[0].y0 = n.y0; q.c[0].y1 = ym;
    q.c[1].x0 = xm;   q.c[1].x1 = n.x1; q.c[1].y0 = n.y0; q.c[1].y1 = ym;
    q.c[2].x0 = n.x0; q.c[2].x1 = xm;   q.c[2].y0 = ym;   q.c[2].y1 = n.y1;
    q.c[3].x0 = xm;   q.c[3].x1 = n.x1; q.c[3].y0 = ym;   q.c[3].y1 = n.y1;
    for (int k : n.keys) {
        const KeyPoint& kp = keys[k];
        const bool left = kp.x < (float)xm, top = kp.y < (float)ym;
        q.c[left ? (top ? 0 : 2) : (top ? 1 : 3)].keys.push_back(k);
    }
    for (auto& c : q.c)
        if (c.keys.size() == 1) c.final_ = true;
}
}  // namespace

std::vector<KeyPoint> distribute_octree(const std::vector<KeyPoint>& keys, int minX, int maxX, int minY, int maxY, int N,
                                        OctreeStats* stats) {
    struct Pending {
        int count;
        uint32_t stamp;
        std::list<Node>::iterator it;
        bool operator<(const Pending& o) const { return count != o.count ? count < o.count : stamp < o.stamp; }
    };
    const int nIni = (int)std::round(static_cast<float>(maxX - minX) / (maxY - minY));
    const float hX = static_cast<float>(maxX - minX) / nIni;
    std::list<Node> nodes;
    std::vector<Node*> roots(nIni);
    uint32_t next_stamp = 0;
    for (int i = 0; i < nIni; ++i) {
        Node r;
        r.x0 = (int)(hX * static_cast<float>(i));
        r.x1 = (int)(hX * static_cast<float>(i + 1));
        r.y0 = 0;
        r.y1 = maxY - minY;
        r.stamp = next_stamp++;
        nodes.push_back(r);
        roots[i] = &nodes.back();
    }
    // list order must stay "descending stamp": roots are pushed to the back, so give them descending stamps
    {
        uint32_t s = (uint32_t)nIni;
        for (auto& n : nodes) n.stamp = --s;
        next_stamp = (uint32_t)nIni;
    }
    for (size_t i = 0; i < keys.size(); ++i) roots[(size_t)(keys[i].x / hX)]->keys.push_back((int)i);
    for (auto it = nodes.begin(); it != nodes.end();) {
        if (it->keys.size() == 1) { it->final_ = true; ++it; }
        else if (it->keys.empty()) it = nodes.erase(it);
        else ++it;
    }

    bool finish = false;
    std::vector<Pending> pending;
    OctreeStats st;
    auto emit_children = [&](Quad& q, std::vector<Pending>& out, int* n_expand) {
        for (auto& c : q.c) {
            if (c.keys.empty()) continue;
            c.stamp = next_stamp++;
            nodes.push_front(std::move(c));
            if (nodes.front().keys.size() > 1) {
                if (n_expand) ++*n_expand;
                nodes.front().self = nodes.begin();
                out.push_back(Pending{(int)nodes.front().keys.size(), nodes.front().stamp, nodes.begin()});
            }
        }
    };
    while (!finish) {
        ++st.rounds;
        int prev = (int)nodes.size();
        int n_expand = 0;
        pending.clear();
        for (auto it = nodes.begin(); it != nodes.end();) {  // round A: split every multi-key node once (:556-615)
            if (it->final_) { ++it; continue; }
            Quad q;
            divide(*it, keys, q);
            emit_children(q, pending, &n_expand);
            it = nodes.erase(it);
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == prev) {
            finish = true;
        } else if ((int)nodes.size() + n_expand * 3 > N) {  // round B: biggest nodes first, stop at N (:624-689)
            while (!finish) {
                ++st.rounds;
                prev = (int)nodes.size();
                std::vector<Pending> todo = pending;
                pending.clear();
                std::sort(todo.begin(), todo.end());
                for (int j = (int)todo.size() - 1; j >= 0; --j) {
                    Quad q;
                    divide(*todo[j].it, keys, q);
                    emit_children(q, pending, nullptr);
                    nodes.erase(todo[j].it);
                    if ((int)nodes.size() >= N) {
                        if (j > 0 && todo[j - 1].count == todo[j].count) st.break_in_tie_group = 1;
                        break;
                    }
                }
                if ((int)nodes.size() >= N || (int)nodes.size() == prev) finish = true;
            }
        }
    }
    std::vector<KeyPoint> result;
    result.reserve(nodes.size());
    for (const auto& n : nodes) {  // best response per node, first wins ties (:696-712)
        int best = n.keys[0];
        float r = keys[best].response;
        for (size_t k = 1; k < n.keys.size(); ++k)
            if (keys[n.keys[k]].response > r) { best = n.keys[k]; r = keys[best].response; }
        result.push_back(keys[best]);
    }
    if (stats) *stats = st;
    return result;
}

// ---------------------------------------------------------------------------------------------------------------
// ORBextractor::ComputeKeyPointsOctTree cell loop (ORBextractor.cpp:717-782)
// ---------------------------------------------------------------------------------------------------------------
static void detect_cells(const Image& lvl, int iniTh, int minTh, std::vector<KeyPoint>& cand) {
    cand.clear();
    const float W = 30;
    const int minBX = kEdge - 3, minBY = minBX, maxBX = lvl.w - kEdge + 3, maxBY = lvl.h - kEdge + 3;
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    if (nCols <= 0 || nRows <= 0) return;
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    std::vector<KeyPoint> cell;
    for (int i = 0; i < nRows; ++i) {
        const float iniY = (float)(minBY + i * hCell);
        float maxY = iniY + hCell + 6;
        if (iniY >= maxBY - 3) continue;
        if (maxY > maxBY) maxY = (float)maxBY;
        for (int j = 0; j < nCols; ++j) {
            const float iniX = (float)(minBX + j * wCell);
            float maxX = iniX + wCell + 6;
            if (iniX >= maxBX - 6) continue;
            if (maxX > maxBX) maxX = (float)maxBX;
            Image roi{lvl.data + (size_t)(int)iniY * lvl.stride + (int)iniX, (int)maxX - (int)iniX, (int)maxY - (int)iniY, lvl.stride};
            fast9_detect(roi, iniTh, true, cell);
            if (cell.empty()) fast9_detect(roi, minTh, true, cell);
            for (auto& kp : cell) {
                kp.x += j * wCell;
                kp.y += i * hCell;
                cand.push_back(kp);
            }
        }
    }
}

// IC_Angle, ORBextractor.cpp:21-48
static float ic_angle(const Image& img, float ptx, float pty, const std::vector<int>& umax) {
    int m01 = 0, m10 = 0;
    const uint8_t* c = img.data + (size_t)cv_round(pty) * img.stride + cv_round(ptx);
    const ptrdiff_t step = (ptrdiff_t)img.stride;
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
    for (int v = 1; v <= kHalfPatch; ++v) {
        int vsum = 0;
        const int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            const int plus = c[u + v * step], minus = c[u - v * step];
            vsum += plus - minus;
            m10 += u * (plus + minus);
        }
        m01 += v * vsum;
    }
    return fast_atan2((float)m01, (float)m10);
}

// computeOrbDescriptor, ORBextractor.cpp:51-97.  cos/sin are the float overloads (using namespace std, :12,56).
static void orb_descriptor(const KeyPoint& kp, const Image& blurred, uint8_t* desc) {
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    const float angle = kp.angle * factorPI;
    const float a = std::cos(angle), b = std::sin(angle);
    const uint8_t* c = blurred.data + (size_t)cv_round(kp.y) * blurred.stride + cv_round(kp.x);
    const int step = (int)blurred.stride;
    const int8_t* pat = kBriefPattern;
    auto sample = [&](int idx) -> int {
        const float px = (float)pat[2 * idx], py = (float)pat[2 * idx + 1];
        return c[cv_round(px * b + py * a) * step + cv_round(px * a - py * b)];
    };
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int k = 0; k < 8; ++k) val |= (sample(2 * k) < sample(2 * k + 1)) << k;
        desc[i] = (uint8_t)val;
    }
}

// ORBextractor::operator(), :990-1055, with ComputePyramid (:1061-1085) and ComputeKeyPointsOctTree (:717-807)
void orb_extract(const OrbParams& p, const Image& img, OrbResult& out, bool keep) {
    out = OrbResult();
    if (img.w == 0 || img.h == 0 || !img.data) return;
    const OrbTables T = make_tables(p);
    const int L = p.nlevels;
    // pyramid: level 0 is the image, level l = resize(level l-1).  The 19-px reflect border the reference adds is
    // never read by this pipeline and is not materialised here.
    out.pyramid.resize(L);
    for (int l = 0; l < L; ++l) {
        const float s = T.inv_scale[l];
        const int w = cv_round((float)img.w * s), h = cv_round((float)img.h * s);
        out.pyramid[l].create(w, h);
        if (l == 0) {
            for (int y = 0; y < h; ++y) std::memcpy(&out.pyramid[0].buf[(size_t)y * w], img.data + (size_t)y * img.stride, w);
        } else {
            resize_linear_u8(out.pyramid[l - 1].view(), out.pyramid[l].buf.data(), w, h, w);
        }
    }
    out.candidates.resize(L);
    out.level_kps.resize(L);
    out.octree_stats.resize(L);
    for (int l = 0; l < L; ++l) {
        const Image lvl = out.pyramid[l].view();
        const int minBX = kEdge - 3, minBY = minBX, maxBX = lvl.w - kEdge + 3, maxBY = lvl.h - kEdge + 3;
        if (maxBX - minBX < 30 || maxBY - minBY < 30) continue;  // nCols/nRows == 0: the reference divides by zero
        detect_cells(lvl, p.iniThFAST, p.minThFAST, out.candidates[l]);
        std::vector<KeyPoint>& kps = out.level_kps[l];
        kps = distribute_octree(out.candidates[l], minBX, maxBX, minBY, maxBY, T.features_per_level[l], &out.octree_stats[l]);
        const int scaledPatch = (int)(kPatchSize * T.scale[l]);
        for (auto& kp : kps) {
            kp.x += minBX;
            kp.y += minBY;
            kp.octave = l;
            kp.size = (float)scaledPatch;
        }
    }
    for (int l = 0; l < L; ++l)
        for (auto& kp : out.level_kps[l]) kp.angle = ic_angle(out.pyramid[l].view(), kp.x, kp.y, T.umax);

    size_t total = 0;
    for (int l = 0; l < L; ++l) total += out.level_kps[l].size();
    out.descriptors.assign(total * 32, 0);
    out.keypoints.reserve(total);
    size_t off = 0;
    OwnedImage blurred;
    for (int l = 0; l < L; ++l) {
        std::vector<KeyPoint>& kps = out.level_kps[l];
        if (kps.empty()) continue;
        blurred.create(out.pyramid[l].w, out.pyramid[l].h);
        gaussian_blur7_sigma2(out.pyramid[l].view(), blurred.buf.data(), blurred.w);
        for (size_t i = 0; i < kps.size(); ++i) orb_descriptor(kps[i], blurred.view(), &out.descriptors[(off + i) * 32]);
        off += kps.size();
        for (const auto& kp : kps) {
            KeyPoint o = kp;
            if (l != 0) { o.x *= T.scale[l]; o.y *= T.scale[l]; }
            out.keypoints.push_back(o);
        }
    }
    if (!keep) { out.pyramid.clear(); out.candidates.clear(); }
}

// The tail of ORBextractor::operator() (:1006-1052) on externally produced intermediates: pyramid levels, their 7x7 blurs and the
// cell loop's candidates (coordinates relative to minBorder).  Used by the CPU baseline variant whose image primitives are the
// REAL OpenCV ones (cv2.resize / cv2.FAST per cell / cv2.GaussianBlur): octree, IC_Angle and the steered BRIEF run here.
void orb_finish(const OrbParams& p, const std::vector<Image>& levels, const std::vector<Image>& blurred,
                const std::vector<std::vector<KeyPoint>>& candidates, std::vector<KeyPoint>& keypoints, std::vector<uint8_t>& descriptors) {
    const OrbTables T = make_tables(p);
    const int L = p.nlevels;
    std::vector<std::vector<KeyPoint>> level_kps(L);
    for (int l = 0; l < L; ++l) {
        const Image& lvl = levels[l];
        const int minBX = kEdge - 3, minBY = minBX, maxBX = lvl.w - kEdge + 3, maxBY = lvl.h - kEdge + 3;
        if (maxBX - minBX < 30 || maxBY - minBY < 30) continue;
        std::vector<KeyPoint>& kps = level_kps[l];
        kps = distribute_octree(candidates[l], minBX, maxBX, minBY, maxBY, T.features_per_level[l], nullptr);
        const int scaledPatch = (int)(kPatchSize * T.scale[l]);
        for (auto& kp : kps) {
            kp.x += minBX;
            kp.y += minBY;
            kp.octave = l;
            kp.size = (float)scaledPatch;
            kp.angle = ic_angle(lvl, kp.x, kp.y, T.umax);
        }
    }
    size_t total = 0;
    for (int l = 0; l < L; ++l) total += level_kps[l].size();
    descriptors.assign(total * 32, 0);
    keypoints.clear();
    keypoints.reserve(total);
    size_t off = 0;
    for (int l = 0; l < L; ++l) {
        const std::vector<KeyPoint>& kps = level_kps[l];
        for (size_t i = 0; i < kps.size(); ++i) orb_descriptor(kps[i], blurred[l], &descriptors[(off + i) * 32]);
        off += kps.size();
        for (const auto& kp : kps) {
            KeyPoint o = kp;
            if (l != 0) { o.x *= T.scale[l]; o.y *= T.scale[l]; }
            keypoints.push_back(o);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Tracking::ORB_Robust_Find_MuliImage_MatchedFeatures, tracking.cpp:534-581 (the part after the two match() calls)
// ---------------------------------------------------------------------------------------------------------------
int orb_match_filter(const int32_t* idx1, const int32_t* dist1, const int32_t* idx2, const int32_t* dist2, int n_last_left,
                     int n_last_right, int n_cur_left, const KeyPoint* last_left, const KeyPoint* last_right,
                     double feature_match_error, int32_t* accepted) {
    const int des_index = std::min(std::min(n_last_left, n_last_right), n_cur_left);  // quirk Q3
    double min_dist = 10000, max_dist = 0;
    for (int i = 0; i < des_index; ++i) {
        double d = std::max((float)dist1[i], (float)dist2[i]);
        if (d < min_dist) min_dist = d;
        if (d > max_dist) max_dist = d;
    }
    int n = 0;
    for (int i = 0; i < des_index; ++i) {
        const double lim = std::max(2 * min_dist, 30.0);
        if ((float)dist1[i] <= lim && (float)dist2[i] <= lim &&
            std::abs(last_left[i].y - last_right[idx1[i]].y) < feature_match_error) {
            accepted[3 * n] = i;
            accepted[3 * n + 1] = idx1[i];
            accepted[3 * n + 2] = idx2[i];
            ++n;
        }
    }
    return n;
}

// Tracking::deleteBadmatchFeatures, tracking.cpp:623-660 (erase-in-place restated as a keep mask)
int lk_chain_filter(int n, const float* p0, const float* p1, const float* p2, const float* p3, const float* p0r,
                    const uint8_t* s0, const uint8_t* s1, const uint8_t* s2, const uint8_t* s3, double fme, uint8_t* keep) {
    int kept = 0;
    for (int i = 0; i < n; ++i) {
        const bool outside = p3[2 * i] < 0 || p3[2 * i + 1] < 0 || p2[2 * i] < 0 || p2[2 * i + 1] < 0 || p1[2 * i] < 0 ||
                             p1[2 * i + 1] < 0 || p0[2 * i] < 0 || p0[2 * i + 1] < 0 || p0r[2 * i] < 0 || p0r[2 * i + 1] < 0;
        const bool bad = !s0[i] || !s1[i] || !s2[i] || !s3[i];
        const bool epi = std::abs(p0[2 * i + 1] - p1[2 * i + 1]) > fme || std::abs(p2[2 * i + 1] - p3[2 * i + 1]) > fme;
        keep[i] = !(outside || bad || epi);
        kept += keep[i];
    }
    return kept;
}

}  // namespace orc
