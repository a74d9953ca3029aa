// oracle/oracle_prims.cpp -- CPU ORACLE (test infrastructure only, see oracle.h).
// Restatements of the OpenCV primitives the reference calls on the hot path.  OpenCV is an un-vendored
// dependency of the reference (CMakeLists.txt:17, "OpenCV 3"); the behaviour restated here is the published
// algorithm of each primitive and is pinned bit-for-bit against cv2 4.13.0 by tests/test_oracle_vs_cv2.py.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "oracle.h"

namespace orc {

int cv_round(float v) { return (int)lrintf(v); }
int cv_round(double v) { return (int)lrint(v); }
static inline int cv_floor(float v) { return (int)floorf(v); }
static inline int cv_floor(double v) { return (int)floor(v); }
static inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}
static inline short sat_short(int v) { return (short)std::min(32767, std::max(-32768, v)); }
static inline uint8_t sat_u8(int v) { return (uint8_t)std::min(255, std::max(0, v)); }

// ---------------------------------------------------------------------------------------------------------------
// cv::resize(src, dst, Size(dw,dh), 0, 0, INTER_LINEAR) for CV_8UC1 -- call site ORBextractor.cpp:1074.
// 11-bit fixed-point coefficients, int32 horizontal pass, the (>>4, *b, >>16, +2, >>2) vertical pass.
// ---------------------------------------------------------------------------------------------------------------
void resize_linear_u8(const Image& src, uint8_t* dst, int dw, int dh, size_t dstride) {
    const int sw = src.w, sh = src.h;
    const double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
    // OpenCV switches INTER_LINEAR to INTER_AREA for exact 2x decimation
    if (sw == dw * 2 && sh == dh * 2) {
        for (int y = 0; y < dh; ++y) {
            const uint8_t* r0 = src.data + (size_t)(2 * y) * src.stride;
            const uint8_t* r1 = r0 + src.stride;
            for (int x = 0; x < dw; ++x) dst[y * dstride + x] = (uint8_t)((r0[2 * x] + r0[2 * x + 1] + r1[2 * x] + r1[2 * x + 1] + 2) >> 2);
        }
        return;
    }
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> alpha(2 * (size_t)dw), beta(2 * (size_t)dh);
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cv_floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        alpha[2 * dx] = sat_short(cv_round((1.f - fx) * 2048.f));
        alpha[2 * dx + 1] = sat_short(cv_round(fx * 2048.f));
    }
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cv_floor(fy);
        fy -= sy;
        yofs[dy] = sy;
        beta[2 * dy] = sat_short(cv_round((1.f - fy) * 2048.f));
        beta[2 * dy + 1] = sat_short(cv_round(fy * 2048.f));
    }
    std::vector<int> row0(dw), row1(dw);
    auto hpass = [&](int sy, std::vector<int>& row) {
        sy = std::min(std::max(sy, 0), sh - 1);
        const uint8_t* S = src.data + (size_t)sy * src.stride;
        for (int dx = 0; dx < dw; ++dx) {
            int sx = xofs[dx];
            int sx1 = std::min(sx + 1, sw - 1);
            row[dx] = S[sx] * alpha[2 * dx] + S[sx1] * alpha[2 * dx + 1];
        }
    };
    for (int dy = 0; dy < dh; ++dy) {
        hpass(yofs[dy], row0);
        hpass(yofs[dy] + 1, row1);
        const int b0 = beta[2 * dy], b1 = beta[2 * dy + 1];
        for (int dx = 0; dx < dw; ++dx) {
            int v = (((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2;
            dst[dy * dstride + dx] = sat_u8(v);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// cv::FAST(img, kps, threshold, nonmaxSuppression) == FAST-9/16 -- call sites ORBextractor.cpp:763,768 and
// tracking.cpp:101.
// ---------------------------------------------------------------------------------------------------------------
static const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// Largest t for which the pixel is a FAST-9 corner at threshold t (cornerScore<16>), or -1 if it is a corner at
// no t >= 0.  Independent of the threshold passed to cv::FAST for every pixel that is a corner at that threshold.
int fast9_score(const uint8_t* p, size_t stride) {
    int d[25];
    const int v = p[0];
    for (int k = 0; k < 25; ++k) d[k] = v - p[(ptrdiff_t)kRingDy[k & 15] * (ptrdiff_t)stride + kRingDx[k & 15]];
    int best = 0;  // max over the 16 arcs of min_k |d| with consistent sign
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
        for (int i = 1; i < 9; ++i) {
            mn = std::min(mn, d[k + i]);
            mx = std::max(mx, d[k + i]);
        }
        best = std::max(best, std::max(mn, -mx));
    }
    return best - 1;
}

void fast9_detect(const Image& roi, int threshold, bool nms, std::vector<KeyPoint>& out) {
    out.clear();
    const int w = roi.w, h = roi.h;
    threshold = std::min(std::max(threshold, 0), 255);
    if (w < 7 || h < 7) return;
    // score map: 0 for non-corners at this threshold and for the 3-px frame
    std::vector<uint8_t> sc((size_t)w * h, 0);
    std::vector<uint8_t> is_corner((size_t)w * h, 0);
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            const uint8_t* c = roi.data + (size_t)y * roi.stride + x;
            // necessary condition (any 9-arc holds one pixel of every opposite pair): cheap reject
            const int v = c[0], st = (int)roi.stride;
            if (std::abs(v - c[3 * st]) <= threshold && std::abs(v - c[-3 * st]) <= threshold) continue;
            if (std::abs(v - c[3]) <= threshold && std::abs(v - c[-3]) <= threshold) continue;
            // segment test at this threshold on the 16-bit ring masks (9 contiguous set bits of the doubled mask); the exact
            // score is only computed for pixels that pass, like cv::FAST does
            uint32_t mb = 0, md = 0;
            for (int k = 0; k < 16; ++k) {
                const int dd = v - c[(ptrdiff_t)kRingDy[k] * st + kRingDx[k]];
                mb |= (uint32_t)(dd < -threshold) << k;
                md |= (uint32_t)(dd > threshold) << k;
            }
            auto run9 = [](uint32_t m) {
                uint32_t x = m | (m << 16);
                x &= x >> 1;
                x &= x >> 2;
                x &= x >> 4;
                x &= x >> 1;
                return x != 0;
            };
            if (!run9(mb) && !run9(md)) continue;
            int s = fast9_score(c, roi.stride);
            if (s >= threshold) {
                is_corner[(size_t)y * w + x] = 1;
                sc[(size_t)y * w + x] = (uint8_t)s;
            }
        }
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            if (!is_corner[(size_t)y * w + x]) continue;
            const int s = sc[(size_t)y * w + x];
            if (nms) {
                bool keep = true;
                for (int dy = -1; dy <= 1 && keep; ++dy)
                    for (int dx = -1; dx <= 1; ++dx) {
                        if (!dx && !dy) continue;
                        if (s <= sc[(size_t)(y + dy) * w + (x + dx)]) { keep = false; break; }
                    }
                if (!keep) continue;
            }
            out.push_back(KeyPoint{(float)x, (float)y, 7.f, -1.f, (float)s, 0, -1});
        }
}

// ---------------------------------------------------------------------------------------------------------------
// cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) for CV_8UC1 (cv2 4.x fixed-point path) -- call
// site ORBextractor.cpp:1035.  Exact 2-D integer convolution with k=[18,34,48,56,48,34,18]/256 per axis.
// ---------------------------------------------------------------------------------------------------------------
void gaussian_blur7_sigma2(const Image& src, uint8_t* dst, size_t dstride) {
    // separable form of the exact 2-D integer convolution: the horizontal sums (<= 255 * 256) are kept unrounded in 16 bits,
    // the vertical pass rounds once.  Interior loops carry no border logic so that the compiler vectorises them.
    static const int k[7] = {18, 34, 48, 56, 48, 34, 18};
    const int w = src.w, h = src.h;
    std::vector<uint16_t> hrow((size_t)w * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t* S = src.data + (size_t)y * src.stride;
        uint16_t* H = hrow.data() + (size_t)y * w;
        auto edge = [&](int x) {
            uint32_t acc = 0;
            for (int j = 0; j < 7; ++j) acc += k[j] * S[reflect101(x + j - 3, w)];
            H[x] = (uint16_t)acc;
        };
        const int x0 = std::min(3, w), x1 = std::max(x0, w - 3);
        for (int x = 0; x < x0; ++x) edge(x);
        for (int x = x0; x < x1; ++x)
            H[x] = (uint16_t)(18 * (S[x - 3] + S[x + 3]) + 34 * (S[x - 2] + S[x + 2]) + 48 * (S[x - 1] + S[x + 1]) + 56 * S[x]);
        for (int x = x1; x < w; ++x) edge(x);
    }
    for (int y = 0; y < h; ++y) {
        const uint16_t* r[7];
        for (int i = 0; i < 7; ++i) r[i] = hrow.data() + (size_t)reflect101(y + i - 3, h) * w;
        uint8_t* D = dst + (size_t)y * dstride;
        for (int x = 0; x < w; ++x) {
            const uint32_t acc = 18u * ((uint32_t)r[0][x] + r[6][x]) + 34u * ((uint32_t)r[1][x] + r[5][x]) + 48u * ((uint32_t)r[2][x] + r[4][x]) +
                                 56u * (uint32_t)r[3][x];
            D[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// cv::fastAtan2(y, x) in degrees -- call site ORBextractor.cpp:47.  Plain f32 arithmetic, no FMA contraction
// (this file is compiled with -ffp-contract=off).
// ---------------------------------------------------------------------------------------------------------------
float fast_atan2(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale, p5 = 0.1555786518463281f * scale,
                p7 = -0.04432655554792128f * scale;
    float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ---------------------------------------------------------------------------------------------------------------
// DescriptorMatcher("BruteForce-Hamming")->match(query, train) -- call sites tracking.cpp:543-544.
// One result per query: lowest train index attaining the minimum popcount distance.
// ---------------------------------------------------------------------------------------------------------------
void match_hamming(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* train_idx, int32_t* dist) {
    for (int i = 0; i < nq; ++i) {
        uint64_t a[4];
        std::memcpy(a, q + (size_t)i * 32, 32);
        int best = 1 << 30, bi = -1;
        for (int j = 0; j < nt; ++j) {
            uint64_t b[4];
            std::memcpy(b, t + (size_t)j * 32, 32);
            int d = __builtin_popcountll(a[0] ^ b[0]) + __builtin_popcountll(a[1] ^ b[1]) + __builtin_popcountll(a[2] ^ b[2]) +
                    __builtin_popcountll(a[3] ^ b[3]);
            if (d < best) { best = d; bi = j; }
        }
        train_idx[i] = bi;
        dist[i] = bi < 0 ? -1 : best;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// cv::pyrDown (5x5 [1 4 6 4 1]^2 / 256, REFLECT_101, round half up) -- used by buildOpticalFlowPyramid inside
// cv::calcOpticalFlowPyrLK (tracking.cpp:593-618).
// ---------------------------------------------------------------------------------------------------------------
void pyr_down(const Image& src, uint8_t* dst, size_t dstride) {
    static const int k[5] = {1, 4, 6, 4, 1};
    const int w = src.w, h = src.h, dw = (w + 1) / 2, dh = (h + 1) / 2;
    std::vector<int> hrow((size_t)dw * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t* S = src.data + (size_t)y * src.stride;
        for (int x = 0; x < dw; ++x) {
            int acc = 0;
            for (int j = 0; j < 5; ++j) acc += k[j] * S[reflect101(2 * x + j - 2, w)];
            hrow[(size_t)y * dw + x] = acc;
        }
    }
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x) {
            int acc = 0;
            for (int i = 0; i < 5; ++i) acc += k[i] * hrow[(size_t)reflect101(2 * y + i - 2, h) * dw + x];
            dst[y * dstride + x] = (uint8_t)((acc + 128) >> 8);
        }
}

// ---------------------------------------------------------------------------------------------------------------
// cv::calcOpticalFlowPyrLK -- call sites tracking.cpp:593,600,607,613 with winSize 21x21, maxLevel 3,
// TermCriteria(COUNT+EPS, 30, 0.01), flags 0, minEigThreshold 1e-3.
//
// Memory model kept from OpenCV: every pyramid level lives in a buffer padded by `win` pixels on each side
// (REFLECT_101 for the image, constant 0 for the Scharr derivative), rows contiguous.  The tracker reads a
// (win+1)x(win+1) patch whose last column may fall one element past the padded row; that element is the first
// element of the next padded row (or, past the last row, 0 here -- OpenCV reads uninitialised memory there).
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct LkLevel {
    int w, h, pad, pw, ph;     // pw = w + 2*pad
    std::vector<uint8_t> img;  // padded image
    std::vector<short> deriv;  // padded, interleaved (Ix, Iy)
    inline int I(int x, int y) const {  // x,y in level coordinates, may be in [-pad, w+pad]
        size_t idx = (size_t)(y + pad) * pw + (size_t)(x + pad);
        return idx < img.size() ? img[idx] : 0;
    }
    inline int D(int x, int y, int c) const {
        size_t idx = ((size_t)(y + pad) * pw + (size_t)(x + pad)) * 2 + c;
        return idx < deriv.size() ? deriv[idx] : 0;
    }
};

void fill_level_image(LkLevel& L, const Image& src) {
    L.w = src.w; L.h = src.h;
    L.pw = L.w + 2 * L.pad; L.ph = L.h + 2 * L.pad;
    L.img.resize((size_t)L.pw * L.ph);
    for (int y = -L.pad; y < L.h + L.pad; ++y) {
        const uint8_t* S = src.data + (size_t)reflect101(y, L.h) * src.stride;
        uint8_t* Drow = &L.img[(size_t)(y + L.pad) * L.pw];
        for (int x = -L.pad; x < L.w + L.pad; ++x) Drow[x + L.pad] = S[reflect101(x, L.w)];
    }
}

// calcScharrDeriv: vertical then horizontal pass, source reflected (REFLECT_101) at the level's own edges,
// result padded with constant 0.
void fill_level_deriv(LkLevel& L) {
    L.deriv.assign((size_t)L.pw * L.ph * 2, 0);
    const int w = L.w, h = L.h;
    std::vector<int> t0(w + 2), t1(w + 2);
    for (int y = 0; y < h; ++y) {
        const int ym = y > 0 ? y - 1 : (h > 1 ? 1 : 0), yp = y < h - 1 ? y + 1 : (h > 1 ? h - 2 : 0);
        for (int x = 0; x < w; ++x) {
            int a = L.I(x, ym), b = L.I(x, y), c = L.I(x, yp);
            t0[x + 1] = (a + c) * 3 + b * 10;
            t1[x + 1] = c - a;
        }
        const int x0 = w > 1 ? 1 : 0, x1 = w > 1 ? w - 2 : 0;
        t0[0] = t0[x0 + 1]; t0[w + 1] = t0[x1 + 1];
        t1[0] = t1[x0 + 1]; t1[w + 1] = t1[x1 + 1];
        short* drow = &L.deriv[((size_t)(y + L.pad) * L.pw + L.pad) * 2];
        for (int x = 0; x < w; ++x) {
            drow[2 * x] = (short)(t0[x + 2] - t0[x]);
            drow[2 * x + 1] = (short)((t1[x + 2] + t1[x]) * 3 + t1[x + 1] * 10);
        }
    }
}

inline int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

int build_lk_pyramid(const Image& img, int win, int max_level, std::vector<LkLevel>& pyr, bool with_deriv) {
    pyr.clear();
    OwnedImage cur;
    Image view = img;
    int level = 0;
    for (; level <= max_level; ++level) {
        pyr.emplace_back();
        LkLevel& L = pyr.back();
        L.pad = win;
        if (level > 0) {
            OwnedImage nxt;
            nxt.create((view.w + 1) / 2, (view.h + 1) / 2);
            pyr_down(view, nxt.buf.data(), nxt.w);
            cur = std::move(nxt);
            view = cur.view();
        }
        fill_level_image(L, view);
        if (with_deriv) fill_level_deriv(L);
        const int nw = (view.w + 1) / 2, nh = (view.h + 1) / 2;
        if (nw <= win || nh <= win) return level;
    }
    return max_level;
}
}  // namespace

void calc_optical_flow_pyr_lk(const Image& prev, const Image& next, const float* prev_xy, int n, float* next_xy,
                              uint8_t* status, float* err, int win, int max_level, int max_iter, double eps,
                              double min_eig) {
    std::vector<LkLevel> P, Q;
    int lp = build_lk_pyramid(prev, win, max_level, P, true);
    int lq = build_lk_pyramid(next, win, max_level, Q, false);
    max_level = std::min(lp, lq);
    max_iter = std::min(std::max(max_iter, 0), 100);
    eps = std::min(std::max(eps, 0.), 10.);
    const double eps2 = eps * eps;  // TermCriteria epsilon is squared once, compared in double (ddot)
    const float half = (win - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    const float min_eig_f = (float)min_eig;
    for (int i = 0; i < n; ++i) { status[i] = 1; if (err) err[i] = 0; }
    std::vector<short> Iw((size_t)win * win), dIw((size_t)win * win * 2);

    for (int level = max_level; level >= 0; --level) {
        const LkLevel& I = P[level];
        const LkLevel& J = Q[level];
        for (int pt = 0; pt < n; ++pt) {
            float px = prev_xy[2 * pt] * (float)(1. / (1 << level)), py = prev_xy[2 * pt + 1] * (float)(1. / (1 << level));
            float nx, ny;
            if (level == max_level) { nx = px; ny = py; }
            else { nx = next_xy[2 * pt] * 2.f; ny = next_xy[2 * pt + 1] * 2.f; }
            next_xy[2 * pt] = nx; next_xy[2 * pt + 1] = ny;

            px -= half; py -= half;
            int ipx = cv_floor(px), ipy = cv_floor(py);
            if (ipx < -win || ipx >= I.w || ipy < -win || ipy >= I.h) {
                if (level == 0) { status[pt] = 0; if (err) err[pt] = 0; }
                continue;
            }
            float a = px - ipx, b = py - ipy;
            int iw00 = cv_round((1.f - a) * (1.f - b) * (1 << 14)), iw01 = cv_round(a * (1.f - b) * (1 << 14)),
                iw10 = cv_round((1.f - a) * b * (1 << 14)), iw11 = (1 << 14) - iw00 - iw01 - iw10;
            // exact integer accumulation, one conversion (OpenCV accumulates in f32 SIMD lanes)
            int64_t sA11 = 0, sA12 = 0, sA22 = 0;
            for (int y = 0; y < win; ++y)
                for (int x = 0; x < win; ++x) {
                    const int X = ipx + x, Y = ipy + y;
                    int ival = descale(I.I(X, Y) * iw00 + I.I(X + 1, Y) * iw01 + I.I(X, Y + 1) * iw10 + I.I(X + 1, Y + 1) * iw11, 14 - 5);
                    int ixv = descale(I.D(X, Y, 0) * iw00 + I.D(X + 1, Y, 0) * iw01 + I.D(X, Y + 1, 0) * iw10 + I.D(X + 1, Y + 1, 0) * iw11, 14);
                    int iyv = descale(I.D(X, Y, 1) * iw00 + I.D(X + 1, Y, 1) * iw01 + I.D(X, Y + 1, 1) * iw10 + I.D(X + 1, Y + 1, 1) * iw11, 14);
                    Iw[(size_t)y * win + x] = (short)ival;
                    dIw[((size_t)y * win + x) * 2] = (short)ixv;
                    dIw[((size_t)y * win + x) * 2 + 1] = (short)iyv;
                    sA11 += (int64_t)ixv * ixv; sA12 += (int64_t)ixv * iyv; sA22 += (int64_t)iyv * iyv;
                }
            float A11 = (float)sA11 * FLT_SCALE, A12 = (float)sA12 * FLT_SCALE, A22 = (float)sA22 * FLT_SCALE;
            float D = A11 * A22 - A12 * A12;
            float minEig = (A22 + A11 - std::sqrt((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (2 * win * win);
            if (minEig < min_eig_f || D < FLT_EPSILON) {
                if (level == 0) status[pt] = 0;
                continue;
            }
            D = 1.f / D;
            nx -= half; ny -= half;
            float pdx = 0, pdy = 0;
            for (int j = 0; j < max_iter; ++j) {
                int inx = cv_floor(nx), iny = cv_floor(ny);
                if (inx < -win || inx >= J.w || iny < -win || iny >= J.h) {
                    if (level == 0) status[pt] = 0;
                    break;
                }
                a = nx - inx; b = ny - iny;
                iw00 = cv_round((1.f - a) * (1.f - b) * (1 << 14)); iw01 = cv_round(a * (1.f - b) * (1 << 14));
                iw10 = cv_round((1.f - a) * b * (1 << 14)); iw11 = (1 << 14) - iw00 - iw01 - iw10;
                int64_t sb1 = 0, sb2 = 0;
                for (int y = 0; y < win; ++y)
                    for (int x = 0; x < win; ++x) {
                        const int X = inx + x, Y = iny + y;
                        int diff = descale(J.I(X, Y) * iw00 + J.I(X + 1, Y) * iw01 + J.I(X, Y + 1) * iw10 + J.I(X + 1, Y + 1) * iw11, 14 - 5) -
                                   Iw[(size_t)y * win + x];
                        sb1 += (int64_t)diff * dIw[((size_t)y * win + x) * 2];
                        sb2 += (int64_t)diff * dIw[((size_t)y * win + x) * 2 + 1];
                    }
                float b1 = (float)sb1 * FLT_SCALE, b2 = (float)sb2 * FLT_SCALE;
                float dx = (float)((A12 * b2 - A22 * b1) * D), dy = (float)((A12 * b1 - A11 * b2) * D);
                nx += dx; ny += dy;
                next_xy[2 * pt] = nx + half; next_xy[2 * pt + 1] = ny + half;
                if ((double)dx * dx + (double)dy * dy <= eps2) break;
                if (j > 0 && (double)std::fabs(dx + pdx) < 0.01 && (double)std::fabs(dy + pdy) < 0.01) {
                    next_xy[2 * pt] -= dx * 0.5f; next_xy[2 * pt + 1] -= dy * 0.5f;
                    break;
                }
                pdx = dx; pdy = dy;
            }
            if (status[pt] && err && level == 0) {
                float fx = next_xy[2 * pt] - half, fy = next_xy[2 * pt + 1] - half;
                int inx = cv_floor(fx), iny = cv_floor(fy);
                if (inx < -win || inx >= J.w || iny < -win || iny >= J.h) { status[pt] = 0; continue; }
                float aa = fx - inx, bb = fy - iny;
                iw00 = cv_round((1.f - aa) * (1.f - bb) * (1 << 14)); iw01 = cv_round(aa * (1.f - bb) * (1 << 14));
                iw10 = cv_round((1.f - aa) * bb * (1 << 14)); iw11 = (1 << 14) - iw00 - iw01 - iw10;
                float errval = 0.f;
                for (int y = 0; y < win; ++y)
                    for (int x = 0; x < win; ++x) {
                        const int X = inx + x, Y = iny + y;
                        int diff = descale(J.I(X, Y) * iw00 + J.I(X + 1, Y) * iw01 + J.I(X, Y + 1) * iw10 + J.I(X + 1, Y + 1) * iw11, 14 - 5) -
                                   Iw[(size_t)y * win + x];
                        errval += std::abs((float)diff);
                    }
                err[pt] = errval * 1.f / (32 * win * win);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// cv::triangulatePoints (calib3d): per point, SVD of the 4x4 matrix with rows x*P[2]-P[0], y*P[2]-P[1] for both views, take
// the right singular vector of the smallest singular value (double), store as float; convertPointsFromHomogeneous then
// scales X, Y, Z by 1.f / W in float.  One-sided (Hestenes) Jacobi like OpenCV's JacobiSVD.
// ---------------------------------------------------------------------------------------------------------------
void triangulate_points(const double* P1, const double* P2, const float* pts1, const float* pts2, int n, float* out4, float* out3) {
    for (int i = 0; i < n; ++i) {
        double a[4][4], v[4][4];  // columns
        const double x1 = pts1[2 * i], y1 = pts1[2 * i + 1], x2 = pts2[2 * i], y2 = pts2[2 * i + 1];
        for (int k = 0; k < 4; ++k) {
            a[k][0] = x1 * P1[8 + k] - P1[k];
            a[k][1] = y1 * P1[8 + k] - P1[4 + k];
            a[k][2] = x2 * P2[8 + k] - P2[k];
            a[k][3] = y2 * P2[8 + k] - P2[4 + k];
            for (int r = 0; r < 4; ++r) v[k][r] = k == r ? 1.0 : 0.0;
        }
        const double eps = DBL_EPSILON * 2;
        for (int sweep = 0; sweep < 30; ++sweep) {
            bool changed = false;
            for (int p = 0; p < 3; ++p)
                for (int q = p + 1; q < 4; ++q) {
                    double alpha = 0, beta = 0, gamma = 0;
                    for (int r = 0; r < 4; ++r) {
                        alpha += a[p][r] * a[p][r];
                        beta += a[q][r] * a[q][r];
                        gamma += a[p][r] * a[q][r];
                    }
                    if (std::fabs(gamma) <= eps * std::sqrt(alpha * beta)) continue;
                    changed = true;
                    const double zeta = (beta - alpha) / (2.0 * gamma);
                    const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                    const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
                    for (int r = 0; r < 4; ++r) {
                        const double ap = a[p][r], aq = a[q][r];
                        a[p][r] = c * ap - s * aq;
                        a[q][r] = s * ap + c * aq;
                        const double vp = v[p][r], vq = v[q][r];
                        v[p][r] = c * vp - s * vq;
                        v[q][r] = s * vp + c * vq;
                    }
                }
            if (!changed) break;
        }
        int jmin = 0;
        double wmin = 1e300;
        for (int k = 0; k < 4; ++k) {
            double w = 0;
            for (int r = 0; r < 4; ++r) w += a[k][r] * a[k][r];
            if (w < wmin) { wmin = w; jmin = k; }
        }
        float h[4];
        for (int r = 0; r < 4; ++r) h[r] = (float)v[jmin][r];
        if (out4) for (int r = 0; r < 4; ++r) out4[4 * i + r] = h[r];
        const float scale = h[3] != 0.f ? 1.f / h[3] : 1.f;
        out3[3 * i] = h[0] * scale;
        out3[3 * i + 1] = h[1] * scale;
        out3[3 * i + 2] = h[2] * scale;
    }
}

}  // namespace orc
