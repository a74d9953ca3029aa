// k_pnp.cu -- K13: the pose step, cv::solvePnPRansac as Tracking::OpenCV_EstimatePose_PnP calls it (tracking.cpp:464-501;
// SURVEY.md section 8f row 2): 3-D points of frame t-1 (K12) + their pixels in left_t -> rvec, tvec, rotation, inlier set.
//
// One CTA of 128 threads per problem (one matched list of one frame).  The RANSAC of OpenCV is sequential -- a better model
// lowers the iteration limit (RANSACUpdateNumIters) -- so hypotheses are evaluated in WAVES of 64 and the sequential rule is
// replayed over each wave's inlier counts afterwards; the result is what the sequential loop returns, because hypothesis k
// depends only on the generator state, never on earlier models:
//   A  thread 0 draws the wave's 5-point subsets: cv::RNG((uint64)-1) multiply-with-carry stream (the raw 32-bit stream is the
//      same for every problem, so it comes from a table made on the host; past its end the generator is stepped), uniform(0, n)
//      by modulo, re-draw on a repeated index;
//   B  thread h solves hypothesis h with EPnP in double (control points by PCA, barycentric coordinates, M^T M, OpenCV's
//      one-sided Jacobi SVD on 12 x 12, three beta approximations x 5 Gauss-Newton steps, absolute orientation, the smallest
//      mean reprojection error wins) -- one thread per hypothesis, its 12 x 12 matrix in shared memory ([element][thread]:
//      conflict-free), the rest in registers / local memory: the Jacobi sweeps are a sequential dependence chain, and
//      2 x 64 x 148 of them in flight is the parallelism;
//   C  consensus: points staged through shared memory in chunks, warp w counts the inliers of hypotheses w, w + 4, ...:
//      projection in double, rounded to float, squared distance in float, inlier iff <= (float)(thr * thr); four points per
//      lane in flight, warp reduction of the count;
//   D  thread 0 replays `count > max(best, 4) -> keep, lower the limit` over the wave in hypothesis order; the loop ends when
//      the hypothesis index reaches the limit;
//   E  inlier mask of the best model, then the Levenberg-Marquardt refinement of [rvec, tvec] over the inliers (CvLevMarq
//      schedule) with block-wide reductions of J^T J / J^T e.
// Steps A-D use + - * / sqrt only and the library is compiled with --fmad=false: every hypothesis is bit-identical to a scalar
// CPU evaluation in the same operation order, so the discrete outputs (best hypothesis, iteration count, inlier mask) can be
// tested for equality.  FP64 is deliberate: cv2 computes all of this in double and a 0.5-px threshold on ~1000-px coordinates
// leaves no room for float.  Bound: the latency of one thread's EPnP (~0.5 ms per wave whatever its size: dependent FP64 and
// shared-memory instructions; issue slots 11 % used) -- no HBM or tensor bound, a problem reads 20 bytes per point.
// The ORB list and the LK list of every frame of a batch are problems of ONE launch (blockIdx.y = list kind).
#include <cfloat>

#include "svo_common.cuh"

namespace {

constexpr int PNP_THREADS = 64;    // hypotheses per wave (one thread each in step B)
constexpr int PNP_BLOCK = 128;     // threads per CTA: steps C and E use all of them.  74 KB + 30 KB of shared memory and 128 x 255
                                   // registers: two CTAs per SM, so the ORB and the LK list of every frame of a batch (2 x 128 CTAs) are resident at once
constexpr int PNP_CHUNK = 1024;  // points staged per consensus chunk (20 KB)
constexpr int PNP_DYN_SMEM = 144 * PNP_THREADS * 8;  // the 12 x 12 matrices of a wave: 74 KB

struct PnpParams {
    double fx, fy, cx, cy;
    int max_iters;
    float thr2;
    double confidence;
};

struct PnpProblemSet {  // problem b of a launch -> slot = (first + b) % ring (standalone: first 0, ring = n)
    const float* p3;    // [slot][cap][3]
    const float* p2;    // [slot * nlists + list][cap][2]
    const int* counts;  // [slot]
    int cap, nlists, list, first, ring;
    double* pose;       // [slot][16]: rvec, tvec, R (row-major), spare
    int* info;          // [slot][4]: ok, inliers, best hypothesis, hypotheses evaluated
    uint8_t* mask;      // [slot][cap]
};

// cv::SVD on n x n (core/src/lapack.cpp JacobiSVDImpl_<double>): rows of `at` are rotated pairwise until orthogonal, sorted by
// decreasing norm and normalised -> at = U^T, w = singular values, vt = V^T (optional).  sqrt(p^2 + beta^2) stands for hypot.
template <int ST, int n>
__device__ void jacobi_svd(double* at, double* w, double* vt) {
    double W[12];
    for (int i = 0; i < n; ++i) {
        double sd = 0;
        for (int k = 0; k < n; ++k) sd += at[(i * n + k) * ST] * at[(i * n + k) * ST];
        W[i] = sd;
        if (vt)
            for (int k = 0; k < n; ++k) vt[i * n + k] = i == k ? 1.0 : 0.0;
    }
    const double eps = DBL_EPSILON * 10;
    const int max_iter = n > 30 ? n : 30;
#pragma unroll 1
    for (int iter = 0; iter < max_iter; ++iter) {
        bool changed = false;
#pragma unroll 1
        for (int i = 0; i < n - 1; ++i)
#pragma unroll 1
            for (int j = i + 1; j < n; ++j) {
                double* ai = at + i * n * ST;
                double* aj = at + j * n * ST;
                double a = W[i], p = 0, b = W[j];
#pragma unroll
                for (int k = 0; k < n; ++k) p += ai[k * ST] * aj[k * ST];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                const double beta = a - b, gamma = sqrt(p * p + beta * beta);
                double c, s;
                if (beta < 0) {
                    const double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
#pragma unroll
                for (int k = 0; k < n; ++k) {
                    const double t0 = c * ai[k * ST] + s * aj[k * ST], t1 = -s * ai[k * ST] + c * aj[k * ST];
                    ai[k * ST] = t0;
                    aj[k * ST] = t1;
                    a += t0 * t0;
                    b += t1 * t1;
                }
                W[i] = a;
                W[j] = b;
                changed = true;
                if (vt) {
                    double* vi = vt + i * n;
                    double* vj = vt + j * n;
                    for (int k = 0; k < n; ++k) {
                        const double t0 = c * vi[k] + s * vj[k], t1 = -s * vi[k] + c * vj[k];
                        vi[k] = t0;
                        vj[k] = t1;
                    }
                }
            }
        if (!changed) break;
    }
    for (int i = 0; i < n; ++i) {
        double sd = 0;
        for (int k = 0; k < n; ++k) sd += at[(i * n + k) * ST] * at[(i * n + k) * ST];
        W[i] = sqrt(sd);
    }
    for (int i = 0; i < n - 1; ++i) {
        int j = i;
        for (int k = i + 1; k < n; ++k)
            if (W[j] < W[k]) j = k;
        if (i != j) {
            const double tw = W[i];
            W[i] = W[j];
            W[j] = tw;
            for (int k = 0; k < n; ++k) {
                const double t = at[(i * n + k) * ST];
                at[(i * n + k) * ST] = at[(j * n + k) * ST];
                at[(j * n + k) * ST] = t;
            }
            if (vt)
                for (int k = 0; k < n; ++k) {
                    const double t = vt[i * n + k];
                    vt[i * n + k] = vt[j * n + k];
                    vt[j * n + k] = t;
                }
        }
    }
    for (int i = 0; i < n; ++i) {
        w[i] = W[i];
        const double s = W[i] > DBL_MIN ? 1 / W[i] : 0.;
        for (int k = 0; k < n; ++k) at[(i * n + k) * ST] *= s;
    }
}

// min |A x - b|, m x n (m = 6, n <= 5): Householder QR without pivoting (epnp.cpp qr_solve / cvSolve(DECOMP_SVD))
__device__ void lstsq_qr(double* A, double* b, const int m, const int n, double* x) {
    for (int k = 0; k < n; ++k) {
        double norm2 = 0.0;
        for (int i = k; i < m; ++i) norm2 += A[i * n + k] * A[i * n + k];
        const double norm = sqrt(norm2);
        if (norm == 0.0) continue;
        const double alpha = A[k * n + k] > 0.0 ? -norm : norm;
        const double vk = A[k * n + k] - alpha;
        double vnorm2 = vk * vk;
        for (int i = k + 1; i < m; ++i) vnorm2 += A[i * n + k] * A[i * n + k];
        if (vnorm2 == 0.0) continue;
        for (int j = k + 1; j < n; ++j) {
            double dot = vk * A[k * n + j];
            for (int i = k + 1; i < m; ++i) dot += A[i * n + k] * A[i * n + j];
            const double f = 2.0 * dot / vnorm2;
            A[k * n + j] -= f * vk;
            for (int i = k + 1; i < m; ++i) A[i * n + j] -= f * A[i * n + k];
        }
        {
            double dot = vk * b[k];
            for (int i = k + 1; i < m; ++i) dot += A[i * n + k] * b[i];
            const double f = 2.0 * dot / vnorm2;
            b[k] -= f * vk;
            for (int i = k + 1; i < m; ++i) b[i] -= f * A[i * n + k];
        }
        A[k * n + k] = alpha;
    }
    for (int k = n - 1; k >= 0; --k) {
        double s = b[k];
        for (int j = k + 1; j < n; ++j) s -= A[k * n + j] * x[j];
        x[k] = A[k * n + k] != 0.0 ? s / A[k * n + k] : 0.0;
    }
}

__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ double dist2(const double* a, const double* b) {
    return (a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]);
}

struct EpnpState {  // the five correspondences of one hypothesis and what the stages hand on (thread-local)
    double pws[15], us[10], cws[4][3], alphas[20], ccs[4][3], pcs[15];
};

__device__ void gauss_newton(const double* l, const double* rho, double* betas) {
    for (int it = 0; it < 5; ++it) {
        double A[24], b[6], x[4];
        for (int i = 0; i < 6; ++i) {
            const double* L = l + i * 10;
            A[i * 4 + 0] = 2 * L[0] * betas[0] + L[1] * betas[1] + L[3] * betas[2] + L[6] * betas[3];
            A[i * 4 + 1] = L[1] * betas[0] + 2 * L[2] * betas[1] + L[4] * betas[2] + L[7] * betas[3];
            A[i * 4 + 2] = L[3] * betas[0] + L[4] * betas[1] + 2 * L[5] * betas[2] + L[8] * betas[3];
            A[i * 4 + 3] = L[6] * betas[0] + L[7] * betas[1] + L[8] * betas[2] + 2 * L[9] * betas[3];
            b[i] = rho[i] - (L[0] * betas[0] * betas[0] + L[1] * betas[0] * betas[1] + L[2] * betas[1] * betas[1] + L[3] * betas[0] * betas[2] +
                             L[4] * betas[1] * betas[2] + L[5] * betas[2] * betas[2] + L[6] * betas[0] * betas[3] + L[7] * betas[1] * betas[3] +
                             L[8] * betas[2] * betas[3] + L[9] * betas[3] * betas[3]);
        }
        lstsq_qr(A, b, 6, 4, x);
        for (int i = 0; i < 4; ++i) betas[i] += x[i];
    }
}

// epnp.cpp compute_R_and_t for the 5 points: camera-frame control points from the betas, sign, absolute orientation, error
__device__ double compute_R_and_t(EpnpState& e, const PnpParams& P, const double* ut4, const double* betas, double* R, double* t) {
    constexpr int n = 5;
    for (int i = 0; i < 4; ++i) e.ccs[i][0] = e.ccs[i][1] = e.ccs[i][2] = 0.0;
    for (int i = 0; i < 4; ++i) {
        const double* v = ut4 + 12 * i;
        for (int j = 0; j < 4; ++j)
            for (int k = 0; k < 3; ++k) e.ccs[j][k] += betas[i] * v[3 * j + k];
    }
    for (int i = 0; i < n; ++i) {
        const double* a = e.alphas + 4 * i;
        for (int j = 0; j < 3; ++j) e.pcs[3 * i + j] = a[0] * e.ccs[0][j] + a[1] * e.ccs[1][j] + a[2] * e.ccs[2][j] + a[3] * e.ccs[3][j];
    }
    if (e.pcs[2] < 0.0) {
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 3; ++j) e.ccs[i][j] = -e.ccs[i][j];
        for (int i = 0; i < 3 * n; ++i) e.pcs[i] = -e.pcs[i];
    }
    double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < 3; ++j) {
            pc0[j] += e.pcs[3 * i + j];
            pw0[j] += e.pws[3 * i + j];
        }
    for (int j = 0; j < 3; ++j) {
        pc0[j] /= n;
        pw0[j] /= n;
    }
    double abtT[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, d[3], vt[9];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < 3; ++j) {
            abtT[j] += (e.pcs[3 * i + j] - pc0[j]) * (e.pws[3 * i] - pw0[0]);
            abtT[3 + j] += (e.pcs[3 * i + j] - pc0[j]) * (e.pws[3 * i + 1] - pw0[1]);
            abtT[6 + j] += (e.pcs[3 * i + j] - pc0[j]) * (e.pws[3 * i + 2] - pw0[2]);
        }
    jacobi_svd<1, 3>(abtT, d, vt);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[3 * i + j] = abtT[i] * vt[j] + abtT[3 + i] * vt[3 + j] + abtT[6 + i] * vt[6 + j];
    const double det = R[0] * R[4] * R[8] + R[1] * R[5] * R[6] + R[2] * R[3] * R[7] - R[2] * R[4] * R[6] - R[1] * R[3] * R[8] - R[0] * R[5] * R[7];
    if (det < 0) {
        R[6] = -R[6];
        R[7] = -R[7];
        R[8] = -R[8];
    }
    for (int i = 0; i < 3; ++i) t[i] = pc0[i] - dot3(R + 3 * i, pw0);
    double sum2 = 0.0;
    for (int i = 0; i < n; ++i) {
        const double* pw = e.pws + 3 * i;
        const double Xc = dot3(R, pw) + t[0], Yc = dot3(R + 3, pw) + t[1], inv_Zc = 1.0 / (dot3(R + 6, pw) + t[2]);
        const double ue = P.cx + P.fx * Xc * inv_Zc, ve = P.cy + P.fy * Yc * inv_Zc;
        const double u = e.us[2 * i], v = e.us[2 * i + 1];
        sum2 += sqrt((u - ue) * (u - ue) + (v - ve) * (v - ve));
    }
    return sum2 / n;
}

__device__ __forceinline__ bool finite_d(double v) { return fabs(v) <= DBL_MAX; }  // false for NaN and infinities

// EPnP on 5 correspondences (s3: 15 floats, s2: 10 floats) -> model[12] = R (row-major), t.  false: degenerate (non-finite).
__device__ bool epnp5(const PnpParams& P, const float* s3, const float* s2, double* model, double* mtm /* shared: element e of this thread at mtm[e * PNP_THREADS] */) {
    constexpr int n = 5;
    EpnpState e;
    const double ifx = 1.0 / P.fx, ify = 1.0 / P.fy;
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < 3; ++j) e.pws[3 * i + j] = s3[3 * i + j];
        const float xn = (float)(((double)s2[2 * i] - P.cx) * ifx), yn = (float)(((double)s2[2 * i + 1] - P.cy) * ify);
        e.us[2 * i] = (double)xn * P.fx + P.cx;
        e.us[2 * i + 1] = (double)yn * P.fy + P.cy;
    }
    {  // choose_control_points
        e.cws[0][0] = e.cws[0][1] = e.cws[0][2] = 0.0;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < 3; ++j) e.cws[0][j] += e.pws[3 * i + j];
        for (int j = 0; j < 3; ++j) e.cws[0][j] /= n;
        double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < n; ++i) {
            double d[3];
            for (int j = 0; j < 3; ++j) d[j] = e.pws[3 * i + j] - e.cws[0][j];
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) c[3 * a + b] += d[a] * d[b];
        }
        double dc[3];
        jacobi_svd<1, 3>(c, dc, nullptr);
        for (int i = 1; i < 4; ++i) {
            const double k = sqrt(dc[i - 1] / n);
            for (int j = 0; j < 3; ++j) e.cws[i][j] = e.cws[0][j] + k * c[3 * (i - 1) + j];
        }
    }
    {  // compute_barycentric_coordinates
        double cc[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 1; j < 4; ++j) cc[3 * i + j - 1] = e.cws[j][i] - e.cws[0][i];
        const double det = cc[0] * (cc[4] * cc[8] - cc[5] * cc[7]) - cc[1] * (cc[3] * cc[8] - cc[5] * cc[6]) + cc[2] * (cc[3] * cc[7] - cc[4] * cc[6]);
        if (det == 0.0 || !(fabs(det) > 0.0)) return false;
        const double id = 1.0 / det;
        double ci[9];
        ci[0] = (cc[4] * cc[8] - cc[5] * cc[7]) * id;
        ci[1] = (cc[2] * cc[7] - cc[1] * cc[8]) * id;
        ci[2] = (cc[1] * cc[5] - cc[2] * cc[4]) * id;
        ci[3] = (cc[5] * cc[6] - cc[3] * cc[8]) * id;
        ci[4] = (cc[0] * cc[8] - cc[2] * cc[6]) * id;
        ci[5] = (cc[2] * cc[3] - cc[0] * cc[5]) * id;
        ci[6] = (cc[3] * cc[7] - cc[4] * cc[6]) * id;
        ci[7] = (cc[1] * cc[6] - cc[0] * cc[7]) * id;
        ci[8] = (cc[0] * cc[4] - cc[1] * cc[3]) * id;
        for (int i = 0; i < n; ++i) {
            const double* pi = e.pws + 3 * i;
            double* a = e.alphas + 4 * i;
            for (int j = 0; j < 3; ++j)
                a[1 + j] = ci[3 * j] * (pi[0] - e.cws[0][0]) + ci[3 * j + 1] * (pi[1] - e.cws[0][1]) + ci[3 * j + 2] * (pi[2] - e.cws[0][2]);
            a[0] = 1.0 - a[1] - a[2] - a[3];
        }
    }
    constexpr int ST = PNP_THREADS;
    for (int i = 0; i < 144; ++i) mtm[i * ST] = 0.0;
    for (int i = 0; i < n; ++i) {
        const double* a = e.alphas + 4 * i;
        double m1[12], m2[12];
        for (int j = 0; j < 4; ++j) {
            m1[3 * j] = a[j] * P.fx;
            m1[3 * j + 1] = 0.0;
            m1[3 * j + 2] = a[j] * (P.cx - e.us[2 * i]);
            m2[3 * j] = 0.0;
            m2[3 * j + 1] = a[j] * P.fy;
            m2[3 * j + 2] = a[j] * (P.cy - e.us[2 * i + 1]);
        }
        for (int r = 0; r < 12; ++r)
            for (int c = r; c < 12; ++c) mtm[(r * 12 + c) * ST] += m1[r] * m1[c] + m2[r] * m2[c];
    }
    for (int r = 0; r < 12; ++r)
        for (int c = 0; c < r; ++c) mtm[(r * 12 + c) * ST] = mtm[(c * 12 + r) * ST];
    double d12[12];
    jacobi_svd<ST, 12>(mtm, d12, nullptr);
    // the four rows epnp.cpp uses (v[i] = row 11 - i of U^T) move to thread-local storage: ut4[i * 12 + k]
    double ut4[48];
    for (int i = 0; i < 4; ++i)
        for (int k = 0; k < 12; ++k) ut4[i * 12 + k] = mtm[((11 - i) * 12 + k) * ST];

    double l[60], rho[6];
    {
        double dv[4][6][3];
        for (int i = 0; i < 4; ++i) {
            const double* v = ut4 + 12 * i;
            int a = 0, b = 1;
            for (int j = 0; j < 6; ++j) {
                dv[i][j][0] = v[3 * a] - v[3 * b];
                dv[i][j][1] = v[3 * a + 1] - v[3 * b + 1];
                dv[i][j][2] = v[3 * a + 2] - v[3 * b + 2];
                if (++b > 3) {
                    ++a;
                    b = a + 1;
                }
            }
        }
        for (int i = 0; i < 6; ++i) {
            double* row = l + 10 * i;
            row[0] = dot3(dv[0][i], dv[0][i]);
            row[1] = 2.0 * dot3(dv[0][i], dv[1][i]);
            row[2] = dot3(dv[1][i], dv[1][i]);
            row[3] = 2.0 * dot3(dv[0][i], dv[2][i]);
            row[4] = 2.0 * dot3(dv[1][i], dv[2][i]);
            row[5] = dot3(dv[2][i], dv[2][i]);
            row[6] = 2.0 * dot3(dv[0][i], dv[3][i]);
            row[7] = 2.0 * dot3(dv[1][i], dv[3][i]);
            row[8] = 2.0 * dot3(dv[2][i], dv[3][i]);
            row[9] = dot3(dv[3][i], dv[3][i]);
        }
        rho[0] = dist2(e.cws[0], e.cws[1]);
        rho[1] = dist2(e.cws[0], e.cws[2]);
        rho[2] = dist2(e.cws[0], e.cws[3]);
        rho[3] = dist2(e.cws[1], e.cws[2]);
        rho[4] = dist2(e.cws[1], e.cws[3]);
        rho[5] = dist2(e.cws[2], e.cws[3]);
    }
    double betas[3][4];
    {  // find_betas_approx_1: [B11 B12 B13 B14]
        double A[24], b[6], x[4];
        for (int i = 0; i < 6; ++i) {
            A[4 * i] = l[10 * i];
            A[4 * i + 1] = l[10 * i + 1];
            A[4 * i + 2] = l[10 * i + 3];
            A[4 * i + 3] = l[10 * i + 6];
            b[i] = rho[i];
        }
        lstsq_qr(A, b, 6, 4, x);
        double* be = betas[0];
        if (x[0] < 0) {
            be[0] = sqrt(-x[0]);
            be[1] = -x[1] / be[0];
            be[2] = -x[2] / be[0];
            be[3] = -x[3] / be[0];
        } else {
            be[0] = sqrt(x[0]);
            be[1] = x[1] / be[0];
            be[2] = x[2] / be[0];
            be[3] = x[3] / be[0];
        }
    }
    {  // find_betas_approx_2: [B11 B12 B22]
        double A[18], b[6], x[3];
        for (int i = 0; i < 6; ++i) {
            A[3 * i] = l[10 * i];
            A[3 * i + 1] = l[10 * i + 1];
            A[3 * i + 2] = l[10 * i + 2];
            b[i] = rho[i];
        }
        lstsq_qr(A, b, 6, 3, x);
        double* be = betas[1];
        if (x[0] < 0) {
            be[0] = sqrt(-x[0]);
            be[1] = x[2] < 0 ? sqrt(-x[2]) : 0.0;
        } else {
            be[0] = sqrt(x[0]);
            be[1] = x[2] > 0 ? sqrt(x[2]) : 0.0;
        }
        if (x[1] < 0) be[0] = -be[0];
        be[2] = 0.0;
        be[3] = 0.0;
    }
    {  // find_betas_approx_3: [B11 B12 B22 B13 B23]
        double A[30], b[6], x[5];
        for (int i = 0; i < 6; ++i) {
            for (int j = 0; j < 5; ++j) A[5 * i + j] = l[10 * i + j];
            b[i] = rho[i];
        }
        lstsq_qr(A, b, 6, 5, x);
        double* be = betas[2];
        if (x[0] < 0) {
            be[0] = sqrt(-x[0]);
            be[1] = x[2] < 0 ? sqrt(-x[2]) : 0.0;
        } else {
            be[0] = sqrt(x[0]);
            be[1] = x[2] > 0 ? sqrt(x[2]) : 0.0;
        }
        if (x[1] < 0) be[0] = -be[0];
        be[2] = x[3] / be[0];
        be[3] = 0.0;
    }
    double best_err = 0.0;
    for (int k = 0; k < 3; ++k) {
        double R[9], t[3];
        gauss_newton(l, rho, betas[k]);
        const double err = compute_R_and_t(e, P, ut4, betas[k], R, t);
        // N = 1; if (err2 < err1) N = 2; if (err3 < err_N) N = 3  (a NaN never wins a `<`, and a NaN first error is never beaten)
        if (k == 0 || err < best_err) {
            best_err = err;
            for (int i = 0; i < 9; ++i) model[i] = R[i];
            for (int i = 0; i < 3; ++i) model[9 + i] = t[i];
        }
    }
    if (!finite_d(best_err)) return false;
    for (int i = 0; i < 12; ++i)
        if (!finite_d(model[i])) return false;
    return true;
}

// cv::Rodrigues vector -> matrix with dR/dr (3 x 9)
__device__ void rodrigues_vec_to_mat(const double* r_in, double* R, double* jac) {
    const double theta = sqrt(r_in[0] * r_in[0] + r_in[1] * r_in[1] + r_in[2] * r_in[2]);
    if (theta < DBL_EPSILON) {
        for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
        if (jac) {
            for (int i = 0; i < 27; ++i) jac[i] = 0.0;
            jac[5] = jac[15] = jac[19] = -1;
            jac[7] = jac[11] = jac[21] = 1;
        }
        return;
    }
    const double c = cos(theta), s = sin(theta), c1 = 1.0 - c, itheta = 1.0 / theta;
    const double r[3] = {r_in[0] * itheta, r_in[1] * itheta, r_in[2] * itheta};
    const double rrt[9] = {r[0] * r[0], r[0] * r[1], r[0] * r[2], r[0] * r[1], r[1] * r[1], r[1] * r[2], r[0] * r[2], r[1] * r[2], r[2] * r[2]};
    const double rx[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0};
    for (int k = 0; k < 9; ++k) R[k] = c * ((k % 4 == 0) ? 1.0 : 0.0) + c1 * rrt[k] + s * rx[k];
    if (jac) {
        const double drrt[27] = {r[0] + r[0], r[1], r[2], r[1], 0, 0, r[2], 0, 0, 0, r[0], 0, r[0], r[1] + r[1], r[2], 0, r[2], 0,
                                 0, 0, r[0], 0, 0, r[1], r[0], r[1], r[2] + r[2]};
        const double drx[27] = {0, 0, 0, 0, 0, -1, 0, 1, 0, 0, 0, 1, 0, 0, 0, -1, 0, 0, 0, -1, 0, 1, 0, 0, 0, 0, 0};
        for (int i = 0; i < 3; ++i) {
            const double ri = r[i];
            const double a0 = -s * ri, a1 = (s - 2 * c1 * itheta) * ri, a2 = c1 * itheta, a3 = (c - s * itheta) * ri, a4 = s * itheta;
            for (int k = 0; k < 9; ++k)
                jac[i * 9 + k] = a0 * ((k % 4 == 0) ? 1.0 : 0.0) + a1 * rrt[k] + a2 * drrt[i * 9 + k] + a3 * rx[k] + a4 * drx[i * 9 + k];
        }
    }
}

// cv::Rodrigues matrix -> vector (input is a rotation: the SVD clean-up OpenCV does first is the identity here)
__device__ void rodrigues_mat_to_vec(const double* R, double* r) {
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = c > 1. ? 1. : c < -1. ? -1. : c;
    double theta = acos(c);
    if (s < 1e-5) {
        if (c > 0) {
            r[0] = r[1] = r[2] = 0;
            return;
        }
        double t;
        t = (R[0] + 1) * 0.5;
        rx = sqrt(t > 0. ? t : 0.);
        t = (R[4] + 1) * 0.5;
        ry = sqrt(t > 0. ? t : 0.) * (R[1] < 0 ? -1. : 1.);
        t = (R[8] + 1) * 0.5;
        rz = sqrt(t > 0. ? t : 0.) * (R[2] < 0 ? -1. : 1.);
        if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
        theta /= sqrt(rx * rx + ry * ry + rz * rz);
        r[0] = rx * theta;
        r[1] = ry * theta;
        r[2] = rz * theta;
        return;
    }
    const double vth = 1 / (2 * s) * theta;
    r[0] = rx * vth;
    r[1] = ry * vth;
    r[2] = rz * vth;
}

// RANSACUpdateNumIters (ptsetreg.cpp); pow(1 - ep, 5) by repeated multiplication
__device__ int ransac_update_num_iters(double p, double ep, int max_iters) {
    p = p < 0. ? 0. : p > 1. ? 1. : p;
    ep = ep < 0. ? 0. : ep > 1. ? 1. : ep;
    double num = 1. - p > DBL_MIN ? 1. - p : DBL_MIN;
    const double w = 1. - ep;
    double wn = w;
    for (int i = 1; i < 5; ++i) wn *= w;
    double denom = 1. - wn;
    if (denom < DBL_MIN) return 0;
    num = log(num);
    denom = log(denom);
    return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)rint(num / denom);
}

__device__ __forceinline__ bool is_inlier(const PnpParams& P, const double* m, float X3, float Y3, float Z3, float u2, float v2) {
    const double X = X3, Y = Y3, Z = Z3;
    const double x = m[0] * X + m[1] * Y + m[2] * Z + m[9];
    const double y = m[3] * X + m[4] * Y + m[5] * Z + m[10];
    double z = m[6] * X + m[7] * Y + m[8] * Z + m[11];
    z = z ? 1. / z : 1;
    const float u = (float)(x * z * P.fx + P.cx), v = (float)(y * z * P.fy + P.cy);
    const float dx = u2 - u, dy = v2 - v;
    const float e = dx * dx + dy * dy;
    return e <= P.thr2;
}

// block-wide sum of NV doubles per thread; result valid in thread 0 (and broadcast through red[0..NV))
template <int NV>
__device__ void block_sum(double* v, double* red /* [PNP_BLOCK / 32][NV] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) red[warp * NV + i] = x;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        const int i = threadIdx.x;
        double x = red[i];
#pragma unroll
        for (int wv = 1; wv < PNP_BLOCK / 32; ++wv) x += red[wv * NV + i];
        red[i] = x;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(PNP_BLOCK, 2) k_pnp_ransac(const PnpProblemSet S0, const PnpProblemSet S1, PnpParams P, const uint32_t* __restrict__ rng_table, int rng_table_n,
                                                            uint64_t rng_table_end_state) {
    extern __shared__ double s_mtm[];  // [144][PNP_THREADS]: M^T M / U^T of every hypothesis of the wave
    __shared__ float s_p3[PNP_CHUNK * 3];
    __shared__ float s_p2[PNP_CHUNK * 2];
    __shared__ double s_model[PNP_THREADS][12];
    __shared__ int s_subset[PNP_THREADS][5];
    __shared__ int s_count[PNP_THREADS];
    __shared__ unsigned char s_valid[PNP_THREADS];
    __shared__ double s_best[12];
    __shared__ double s_red[(PNP_BLOCK / 32) * 28];
    __shared__ double s_param[6];
    __shared__ double s_lm[48];
    __shared__ int s_ctl[8];  // 0 niters, 1 max_good, 2 best_iter, 3 done, 4 iter (hypotheses replayed so far), 5 LM go-on flag

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const PnpProblemSet& S = blockIdx.y ? S1 : S0;  // list kind (front-end: 0 ORB associations, 1 LK tracks)
    int slot = S.first + blockIdx.x;
    if (slot >= S.ring) slot -= S.ring;
    const int n = S.counts[slot];
    const float* __restrict__ p3 = S.p3 + (size_t)slot * S.cap * 3;
    const float* __restrict__ p2 = S.p2 + ((size_t)slot * S.nlists + S.list) * S.cap * 2;
    double* pose = S.pose + (size_t)slot * 16;
    int* info = S.info + (size_t)slot * 4;
    uint8_t* mask = S.mask + (size_t)slot * S.cap;

    if (n < 5 || n > S.cap) {  // cv2 needs >= 4 points and switches to P3P at exactly 4; the reference rejects such frames earlier
        if (tid < 16) pose[tid] = (tid == 6 || tid == 10 || tid == 14) ? 1.0 : 0.0;
        if (tid < 4) info[tid] = 0;
        for (int i = tid; i < (n > 0 && n <= S.cap ? n : 0); i += PNP_BLOCK) mask[i] = 0;
        return;
    }
    // generator position: index into the shared raw table, then the live state past its end
    int rng_pos = 0;
    uint64_t rng_state = rng_table_end_state;
    if (tid == 0) {
        s_ctl[0] = P.max_iters > 1 ? P.max_iters : 1;
        s_ctl[1] = 0;
        s_ctl[2] = -1;
        s_ctl[3] = 0;
        s_ctl[4] = 0;
    }
    __syncthreads();

    for (int base = 0;; base += PNP_THREADS) {
        // ---- A: subsets of this wave
        if (tid == 0) {
            const int niters = s_ctl[0];
            for (int h = 0; h < PNP_THREADS && base + h < niters; ++h) {
                if (n > 5) {
                    for (int i = 0; i < 5; ++i) {
                        int v;
                        bool dup;
                        do {
                            uint32_t raw;
                            if (rng_pos < rng_table_n) {
                                raw = rng_table[rng_pos++];
                            } else {
                                rng_state = (uint64_t)(uint32_t)rng_state * 4164903690u + (uint32_t)(rng_state >> 32);
                                raw = (uint32_t)rng_state;
                            }
                            v = (int)(raw % (uint32_t)n);
                            dup = false;
                            for (int j = 0; j < i; ++j) dup |= s_subset[h][j] == v;
                        } while (dup);
                        s_subset[h][i] = v;
                    }
                } else {
                    for (int i = 0; i < 5; ++i) s_subset[h][i] = i;
                }
            }
        }
        __syncthreads();
        const int niters_wave = s_ctl[0];
        const int wave_n = min(PNP_THREADS, niters_wave - base);  // >= 1
        // ---- B: one hypothesis per thread
        if (tid < PNP_THREADS) s_count[tid] = 0;
        if (tid < wave_n) {
            float s3[15], s2[10];
            for (int i = 0; i < 5; ++i) {
                const int idx = s_subset[tid][i];
                for (int j = 0; j < 3; ++j) s3[3 * i + j] = p3[3 * idx + j];
                for (int j = 0; j < 2; ++j) s2[2 * i + j] = p2[2 * idx + j];
            }
            double model[12];
            const bool ok = epnp5(P, s3, s2, model, s_mtm + tid);
            s_valid[tid] = ok ? 1 : 0;
            for (int i = 0; i < 12; ++i) s_model[tid][i] = model[i];
        } else if (tid < PNP_THREADS) {
            s_valid[tid] = 0;
        }
        __syncthreads();
        // ---- C: consensus, points staged in chunks
        for (int c0 = 0; c0 < n; c0 += PNP_CHUNK) {
            const int cn = min(PNP_CHUNK, n - c0);
            for (int i = tid; i < cn * 3; i += PNP_BLOCK) s_p3[i] = p3[(size_t)c0 * 3 + i];
            for (int i = tid; i < cn * 2; i += PNP_BLOCK) s_p2[i] = p2[(size_t)c0 * 2 + i];
            __syncthreads();
            for (int h = warp; h < wave_n; h += PNP_BLOCK / 32) {
                if (!s_valid[h]) continue;
                double m[12];
#pragma unroll
                for (int i = 0; i < 12; ++i) m[i] = s_model[h][i];
                int cnt = 0;
                // four points per lane in flight: one point is a ~45-instruction FP64 dependence chain (with a division), and a CTA
                // has only four warps to hide it
#pragma unroll 4
                for (int i = lane; i < cn; i += 32) cnt += is_inlier(P, m, s_p3[3 * i], s_p3[3 * i + 1], s_p3[3 * i + 2], s_p2[2 * i], s_p2[2 * i + 1]) ? 1 : 0;
                cnt = __reduce_add_sync(0xffffffffu, cnt);
                if (lane == 0) s_count[h] += cnt;
            }
            __syncthreads();
        }
        // ---- D: the sequential acceptance rule over the wave
        if (tid == 0) {
            int niters = s_ctl[0], max_good = s_ctl[1], iter = base;
            for (int h = 0; h < wave_n; ++h, ++iter) {
                if (iter >= niters) break;
                if (!s_valid[h]) continue;
                const int good = s_count[h];
                if (good > (max_good > 4 ? max_good : 4)) {
                    max_good = good;
                    s_ctl[2] = iter;
                    for (int i = 0; i < 12; ++i) s_best[i] = s_model[h][i];
                    niters = ransac_update_num_iters(P.confidence, (double)(n - good) / n, niters);
                }
            }
            s_ctl[0] = niters;
            s_ctl[1] = max_good;
            s_ctl[4] = iter;
            s_ctl[3] = iter >= niters ? 1 : 0;
        }
        __syncthreads();
        if (s_ctl[3]) break;
    }
    const int max_good = s_ctl[1];
    if (max_good <= 0) {
        if (tid < 16) pose[tid] = (tid == 6 || tid == 10 || tid == 14) ? 1.0 : 0.0;
        if (tid == 0) {
            info[0] = 0;
            info[1] = 0;
            info[2] = 0;
            info[3] = s_ctl[4];
        }
        for (int i = tid; i < n; i += PNP_BLOCK) mask[i] = 0;
        return;
    }
    // ---- E: inlier mask of the best model
    {
        double m[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) m[i] = s_best[i];
        for (int i = tid; i < n; i += PNP_BLOCK)
            mask[i] = is_inlier(P, m, p3[3 * i], p3[3 * i + 1], p3[3 * i + 2], p2[2 * i], p2[2 * i + 1]) ? 1 : 0;
    }
    __syncthreads();  // mask (global) written by this block is read below by other threads of the block
    // ---- LM refinement over the inliers (cvFindExtrinsicCameraParams2 with a guess: CvLevMarq, <= 20 iterations, FLT_EPSILON)
    if (tid == 0) {
        double r[3];
        rodrigues_mat_to_vec(s_best, r);
        s_param[0] = r[0];
        s_param[1] = r[1];
        s_param[2] = r[2];
        s_param[3] = s_best[9];
        s_param[4] = s_best[10];
        s_param[5] = s_best[11];
    }
    __syncthreads();
    // LM state of thread 0 (the linearisation at the last accepted point): kept in shared memory
    double* const JtJ = s_lm;
    double* const JtE = s_lm + 36;
    double* const prev = s_lm + 42;
    int lambda_lg10 = -3, iters = 0;
    double prev_err = 0.0;
    bool need_jac = true;
    for (;;) {
        double param[6], R[9], jac[27];
        for (int i = 0; i < 6; ++i) param[i] = s_param[i];
        rodrigues_vec_to_mat(param, R, need_jac ? jac : nullptr);
        double acc[28];
        for (int i = 0; i < 28; ++i) acc[i] = 0.0;
        for (int i = tid; i < n; i += PNP_BLOCK) {
            if (!mask[i]) continue;
            const double X = p3[3 * i], Y = p3[3 * i + 1], Z = p3[3 * i + 2];
            const double x = R[0] * X + R[1] * Y + R[2] * Z + param[3];
            const double y = R[3] * X + R[4] * Y + R[5] * Z + param[4];
            double z = R[6] * X + R[7] * Y + R[8] * Z + param[5];
            z = z ? 1. / z : 1;
            const double xn = x * z, yn = y * z;
            const double ex = xn * P.fx + P.cx - (double)p2[2 * i], ey = yn * P.fy + P.cy - (double)p2[2 * i + 1];
            acc[27] += ex * ex + ey * ey;
            if (need_jac) {
                double ju[6], jv[6];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double* d = jac + 9 * k;
                    const double dx = d[0] * X + d[1] * Y + d[2] * Z, dy = d[3] * X + d[4] * Y + d[5] * Z, dz = d[6] * X + d[7] * Y + d[8] * Z;
                    ju[k] = P.fx * (dx - xn * dz) * z;
                    jv[k] = P.fy * (dy - yn * dz) * z;
                }
                ju[3] = P.fx * z;
                ju[4] = 0;
                ju[5] = -P.fx * xn * z;
                jv[3] = 0;
                jv[4] = P.fy * z;
                jv[5] = -P.fy * yn * z;
                int q = 0;
#pragma unroll
                for (int a = 0; a < 6; ++a) {
#pragma unroll
                    for (int b = a; b < 6; ++b) acc[q++] += ju[a] * ju[b] + jv[a] * jv[b];
                }
#pragma unroll
                for (int a = 0; a < 6; ++a) acc[21 + a] += ju[a] * ex + jv[a] * ey;
            }
        }
        block_sum<28>(acc, s_red);
        if (tid == 0) {
            const double err_norm = sqrt(s_red[27]);
#ifdef SVO_PNP_DEBUG
            if (blockIdx.x == 0)
                printf("LM %s iters %d lambda %d err %.17g prev_err %.17g param %.12g %.12g %.12g %.12g %.12g %.12g\n", need_jac ? "J" : "e", iters, lambda_lg10,
                       err_norm, prev_err, s_param[0], s_param[1], s_param[2], s_param[3], s_param[4], s_param[5]);
#endif
            bool go_on = true, solve = false;
            if (need_jac) {  // CALC_J: remember the linearisation, take a step
                int q = 0;
                for (int a = 0; a < 6; ++a)
                    for (int b = a; b < 6; ++b) {
                        JtJ[6 * a + b] = s_red[q];
                        JtJ[6 * b + a] = s_red[q];
                        ++q;
                    }
                for (int a = 0; a < 6; ++a) JtE[a] = s_red[21 + a];
                for (int i = 0; i < 6; ++i) prev[i] = s_param[i];
                prev_err = err_norm;
                solve = true;
            } else {  // CHECK_ERR
                if (err_norm > prev_err && ++lambda_lg10 <= 16) {
                    solve = true;
                } else {
                    lambda_lg10 = lambda_lg10 - 1 > -16 ? lambda_lg10 - 1 : -16;
                    double dn = 0, pn = 0;
                    for (int i = 0; i < 6; ++i) {
                        dn += (s_param[i] - prev[i]) * (s_param[i] - prev[i]);
                        pn += prev[i] * prev[i];
                    }
                    const double rel = sqrt(dn) / (sqrt(pn) + DBL_EPSILON);
                    if (++iters >= 20 || rel < FLT_EPSILON) go_on = false;
#ifdef SVO_PNP_DEBUG
                    if (blockIdx.x == 0) printf("   accept: rel %.6g dn %.6g pn %.6g iters %d go_on %d eps %.6g\n", rel, dn, pn, iters, (int)go_on, (double)FLT_EPSILON);
#endif
                    s_ctl[6] = 1;  // next evaluation needs the Jacobian
                }
            }
            if (solve) {  // (JtJ with its diagonal scaled by 1 + lambda) x = JtE, Gaussian elimination with partial pivoting
                const double lambda = exp(lambda_lg10 * log(10.));
                double a[6][7], x[6];
                for (int i = 0; i < 6; ++i) {
                    for (int j = 0; j < 6; ++j) a[i][j] = JtJ[6 * i + j];
                    a[i][i] *= 1.0 + lambda;
                    a[i][6] = JtE[i];
                }
                for (int k = 0; k < 6; ++k) {
                    int piv = k;
                    for (int i = k + 1; i < 6; ++i)
                        if (fabs(a[i][k]) > fabs(a[piv][k])) piv = i;
                    if (piv != k)
                        for (int j = 0; j < 7; ++j) {
                            const double t = a[k][j];
                            a[k][j] = a[piv][j];
                            a[piv][j] = t;
                        }
                    if (a[k][k] == 0.0) continue;
                    for (int i = k + 1; i < 6; ++i) {
                        const double f = a[i][k] / a[k][k];
                        for (int j = k; j < 7; ++j) a[i][j] -= f * a[k][j];
                    }
                }
                for (int k = 5; k >= 0; --k) {
                    double s = a[k][6];
                    for (int j = k + 1; j < 6; ++j) s -= a[k][j] * x[j];
                    x[k] = a[k][k] != 0.0 ? s / a[k][k] : 0.0;
                }
                for (int i = 0; i < 6; ++i) s_param[i] = prev[i] - x[i];
                s_ctl[6] = 0;  // next evaluation: error only
            }
            s_ctl[5] = go_on ? 1 : 0;
        }
        __syncthreads();
        if (!s_ctl[5]) break;
        need_jac = s_ctl[6] != 0;
        __syncthreads();
    }
    if (tid == 0) {
        double R[9];
        rodrigues_vec_to_mat(s_param, R, nullptr);
        for (int i = 0; i < 6; ++i) pose[i] = s_param[i];
        for (int i = 0; i < 9; ++i) pose[6 + i] = R[i];
        pose[15] = (double)iters + 0.01 * lambda_lg10;  // spare slot: LM iterations + lambda exponent / 100 (diagnostics)
        info[0] = 1;
        info[1] = max_good;
        info[2] = s_ctl[2];
        info[3] = s_ctl[4];
    }
}

int pnp_ensure_rng(svo_ctx* c) {
    if (c->d_pnp_rng) return SVO_OK;
    SVO_CUDA_TRY(c, cudaFuncSetAttribute(k_pnp_ransac, cudaFuncAttributeMaxDynamicSharedMemorySize, PNP_DYN_SMEM));
    // 500 hypotheses x 5 draws = 2500 raw numbers when no index repeats; 4096 covers the re-draws of all but tiny lists
    const int n = 4096;
    std::vector<uint32_t> t(n);
    uint64_t state = (uint64_t)-1;
    for (int i = 0; i < n; ++i) {
        state = (uint64_t)(uint32_t)state * 4164903690u + (uint32_t)(state >> 32);
        t[i] = (uint32_t)state;
    }
    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_pnp_rng, sizeof(uint32_t) * n));
    SVO_CUDA_TRY(c, cudaMemcpy(c->d_pnp_rng, t.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice));
    c->pnp_rng_n = n;
    c->pnp_rng_end = state;
    return SVO_OK;
}

PnpParams make_params(const double* K, int iters, float reproj, double conf) {
    PnpParams P;
    P.fx = K[0];
    P.fy = K[4];
    P.cx = K[2];
    P.cy = K[5];
    P.max_iters = iters;
    P.thr2 = (float)((double)reproj * (double)reproj);
    P.confidence = conf;
    return P;
}

}  // namespace

int svo_pnp_free_frontend(svo_ctx* c) {
    void** ptrs[] = {(void**)&c->d_pnp_pose, (void**)&c->d_pnp_info, (void**)&c->d_pnp_mask_orb, (void**)&c->d_pnp_mask_lk};
    for (void** p : ptrs) {
        if (*p) svo_dev_free(*p);
        *p = nullptr;
    }
    c->fe_pnp = false;
    return SVO_OK;
}

// Tracking::iterationsCount_ / reprojectionError_ / confidence_ (parameter.cpp:57-59); the intrinsic matrix is the left 3 x 3 of
// projMatr1_ (tracking.cpp:476-477)
extern "C" int svo_frontend_set_pnp(svo_ctx* c, int enable, int iterations, float reproj_error, double confidence) {
    if (!c) return SVO_ERR_INVALID;
    if (!c->fe_configured) return c->fail(SVO_ERR_INVALID, "svo_frontend_configure() first");
    if (c->fe_nq > 0) return c->fail(SVO_ERR_INVALID, "svo_frontend_set_pnp: download (or discard) the pending runs first: their pose buffers would be freed");
    cudaSetDevice(c->device);
    SVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (!enable) return svo_pnp_free_frontend(c);
    if (!c->have_proj) return c->fail(SVO_ERR_INVALID, "svo_frontend_set_pnp: svo_frontend_set_projection() first (the 3-D points come from it)");
    if (iterations < 1 || !(reproj_error > 0.f) || !(confidence > 0.0 && confidence < 1.0))
        return c->fail(SVO_ERR_INVALID, "svo_frontend_set_pnp: iterations >= 1, reproj_error > 0, 0 < confidence < 1");
    int rc;
    if ((rc = pnp_ensure_rng(c)) != SVO_OK) return rc;
    svo_pnp_free_frontend(c);
    const size_t ring = (size_t)c->fe_ring;
    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_pnp_pose, sizeof(double) * 16 * 2 * ring));
    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_pnp_info, sizeof(int) * 4 * 2 * ring));
    SVO_CUDA_TRY(c, cudaMemset(c->d_pnp_pose, 0, sizeof(double) * 16 * 2 * ring));
    SVO_CUDA_TRY(c, cudaMemset(c->d_pnp_info, 0, sizeof(int) * 4 * 2 * ring));
    if (c->fe_mode & SVO_MODE_ORB) SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_pnp_mask_orb, (size_t)c->kp_cap * ring));
    if (c->fe_mode & SVO_MODE_LK) SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_pnp_mask_lk, (size_t)c->pt_cap * ring));
    c->pnp_iters = iterations;
    c->pnp_reproj = reproj_error;
    c->pnp_conf = confidence;
    c->fe_pnp = true;
    return SVO_OK;
}

int launch_pnp_frontend(svo_ctx* c, int first, int ring, int n_frames) {
    const double K[9] = {c->proj1[0], c->proj1[1], c->proj1[2], c->proj1[4], c->proj1[5], c->proj1[6], c->proj1[8], c->proj1[9], c->proj1[10]};
    const PnpParams P = make_params(K, c->pnp_iters, c->pnp_reproj, c->pnp_conf);
    StageTimer timer(c, 12);
    // 3-D points of the accepted ORB matches (K12) and their pixels in left_t: list 0 (t2_left) of d_orb_pts;
    // surviving LK tracks: list 3 (t2_left) of d_lkf_pts.  Both kinds in ONE launch (blockIdx.y), all CTAs resident at once.
    const PnpProblemSet So{c->d_orb_xyz, c->d_orb_pts, c->d_norb, c->kp_cap, 3, 0, first, ring, c->d_pnp_pose, c->d_pnp_info, c->d_pnp_mask_orb};
    const PnpProblemSet Sl{c->d_lkf_xyz, c->d_lkf_pts, c->d_nlkf, c->pt_cap, 5, 3, first, ring, c->d_pnp_pose + (size_t)16 * ring,
                           c->d_pnp_info + (size_t)4 * ring, c->d_pnp_mask_lk};
    const bool orb = c->fe_mode & SVO_MODE_ORB, lk = c->fe_mode & SVO_MODE_LK;
    k_pnp_ransac<<<dim3(n_frames, orb && lk ? 2 : 1), PNP_BLOCK, PNP_DYN_SMEM, c->stream>>>(orb ? So : Sl, Sl, P, c->d_pnp_rng, c->pnp_rng_n, c->pnp_rng_end);
    ++c->launches;
    SVO_CUDA_TRY(c, cudaGetLastError());
    return SVO_OK;
}

extern "C" int svo_pnp_ransac(svo_ctx* c, const double* K, const float* pts3d, const float* pts2d, const int32_t* counts, int n_problems, int cap,
                              int iterations, float reproj_error, double confidence, double* rvec, double* tvec, double* rotation,
                              uint8_t* inlier_mask, int32_t* n_inliers, int32_t* info, int32_t* ok) {
    if (!c) return SVO_ERR_INVALID;
    if (!K || n_problems < 0 || cap < 0 || (n_problems > 0 && (!pts3d || !pts2d || !counts || !rvec || !tvec || !ok)) || iterations < 1 ||
        !(reproj_error > 0.f) || !(confidence > 0.0 && confidence < 1.0))
        return c->fail(SVO_ERR_INVALID, "svo_pnp_ransac: bad arguments");
    if (n_problems == 0) return SVO_OK;
    for (int i = 0; i < n_problems; ++i)
        if (counts[i] < 0 || counts[i] > cap) return c->fail(SVO_ERR_INVALID, "svo_pnp_ransac: counts[i] must be within 0..cap");
    cudaSetDevice(c->device);
    int rc;
    if ((rc = pnp_ensure_rng(c)) != SVO_OK) return rc;
    const size_t np = (size_t)n_problems, capz = (size_t)(cap > 0 ? cap : 1);
    const size_t b3 = svo_align(np * capz * 12, 256), b2 = svo_align(np * capz * 8, 256), bc = svo_align(np * 4, 256), bp = svo_align(np * 128, 256),
                 bi = svo_align(np * 16, 256), bm = svo_align(np * capz, 256);
    if ((rc = svo_ensure_scratch(c, b3 + b2 + bc + bp + bi + bm)) != SVO_OK) return rc;
    uint8_t* base = (uint8_t*)c->d_scratch;
    float* d3 = (float*)base;
    float* d2 = (float*)(base + b3);
    int* dc = (int*)(base + b3 + b2);
    double* dp = (double*)(base + b3 + b2 + bc);
    int* di = (int*)(base + b3 + b2 + bc + bp);
    uint8_t* dm = base + b3 + b2 + bc + bp + bi;
    SVO_CUDA_TRY(c, cudaMemcpyAsync(d3, pts3d, np * capz * 12, cudaMemcpyHostToDevice, c->stream));
    SVO_CUDA_TRY(c, cudaMemcpyAsync(d2, pts2d, np * capz * 8, cudaMemcpyHostToDevice, c->stream));
    SVO_CUDA_TRY(c, cudaMemcpyAsync(dc, counts, np * 4, cudaMemcpyHostToDevice, c->stream));
    PnpProblemSet S{d3, d2, dc, (int)capz, 1, 0, 0, n_problems, dp, di, dm};
    k_pnp_ransac<<<n_problems, PNP_BLOCK, PNP_DYN_SMEM, c->stream>>>(S, S, make_params(K, iterations, reproj_error, confidence), c->d_pnp_rng, c->pnp_rng_n, c->pnp_rng_end);
    ++c->launches;
    SVO_CUDA_TRY(c, cudaGetLastError());
    std::vector<double> hp(np * 16);
    std::vector<int> hi(np * 4);
    SVO_CUDA_TRY(c, cudaMemcpyAsync(hp.data(), dp, np * 128, cudaMemcpyDeviceToHost, c->stream));
    SVO_CUDA_TRY(c, cudaMemcpyAsync(hi.data(), di, np * 16, cudaMemcpyDeviceToHost, c->stream));
    if (inlier_mask && cap > 0) SVO_CUDA_TRY(c, cudaMemcpyAsync(inlier_mask, dm, np * capz, cudaMemcpyDeviceToHost, c->stream));  // (cap == 0: the caller's array is empty)
    SVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < np; ++i) {
        for (int k = 0; k < 3; ++k) {
            rvec[3 * i + k] = hp[16 * i + k];
            tvec[3 * i + k] = hp[16 * i + 3 + k];
        }
        if (rotation)
            for (int k = 0; k < 9; ++k) rotation[9 * i + k] = hp[16 * i + 6 + k];
        ok[i] = hi[4 * i];
        if (n_inliers) n_inliers[i] = hi[4 * i + 1];
        if (info) {
            info[2 * i] = hi[4 * i + 2];
            info[2 * i + 1] = hi[4 * i + 3];
        }
        if (inlier_mask && cap > 0 && !hi[4 * i])
            for (int k = 0; k < counts[i]; ++k) inlier_mask[i * capz + k] = 0;
    }
    return SVO_OK;
}
