// k_triangulate.cu -- K12: cv::triangulatePoints + cv::convertPointsFromHomogeneous (tracking.cpp:190-192, 292-294), the
// second half of SURVEY.md section 8f's first "next" row: the matched point lists become 3-D points on the device.
//
// OpenCV's algorithm (calib3d, triangulate.cpp): per point the 4x4 matrix A (double) with rows
//     x1 * P1[2] - P1[0],   y1 * P1[2] - P1[1],   x2 * P2[2] - P2[0],   y2 * P2[2] - P2[1]
// is decomposed by SVD and the right singular vector of the smallest singular value is the homogeneous point, stored as
// float; convertPointsFromHomogeneous then multiplies X, Y, Z by 1.f / W in float.  Here: one thread per point, one-sided
// (Hestenes) Jacobi SVD in double like OpenCV's JacobiSVD.  The singular vector is unique up to sign (which cancels in the
// division) and to ~1e-13 relative, so the float results agree with OpenCV to the last ulp or two; tests use 2e-5 relative.
#include <cfloat>

#include "svo_common.cuh"

namespace {

struct Proj {
    double p1[12], p2[12];
};

__device__ __forceinline__ void triangulate_one(const Proj& P, float x1, float y1, float x2, float y2, float* out4, float* out3) {
    // columns of A: a[k][r], k = column (0..3), r = row (0..3)
    double a[4][4], v[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        a[k][0] = (double)x1 * P.p1[8 + k] - P.p1[k];
        a[k][1] = (double)y1 * P.p1[8 + k] - P.p1[4 + k];
        a[k][2] = (double)x2 * P.p2[8 + k] - P.p2[k];
        a[k][3] = (double)y2 * P.p2[8 + k] - P.p2[4 + k];
#pragma unroll
        for (int r = 0; r < 4; ++r) v[k][r] = k == r ? 1.0 : 0.0;
    }
    const double eps = DBL_EPSILON * 2;
    for (int sweep = 0; sweep < 30; ++sweep) {
        bool changed = false;
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int q = p + 1; q < 4; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    alpha += a[p][r] * a[p][r];
                    beta += a[q][r] * a[q][r];
                    gamma += a[p][r] * a[q][r];
                }
                if (fabs(gamma) <= eps * sqrt(alpha * beta)) continue;
                changed = true;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
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
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        double w = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) w += a[k][r] * a[k][r];
        if (w < wmin) {
            wmin = w;
            jmin = k;
        }
    }
    float h[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        double val = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) val = k == jmin ? v[k][r] : val;  // static indexing only: keeps v in registers
        h[r] = (float)val;
    }
    if (out4) {
#pragma unroll
        for (int r = 0; r < 4; ++r) out4[r] = h[r];
    }
    const float scale = h[3] != 0.f ? __fdiv_rn(1.f, h[3]) : 1.f;  // convertPointsFromHomogeneous, float path
    out3[0] = __fmul_rn(h[0], scale);
    out3[1] = __fmul_rn(h[1], scale);
    out3[2] = __fmul_rn(h[2], scale);
}

__global__ void __launch_bounds__(128) k_triangulate(Proj P, const float* __restrict__ pts1, const float* __restrict__ pts2, int n,
                                                     float* __restrict__ out4, float* __restrict__ out3) {
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    triangulate_one(P, pts1[2 * i], pts1[2 * i + 1], pts2[2 * i], pts2[2 * i + 1], out4 ? out4 + 4 * i : nullptr, out3 + 3 * i);
}

// frontend: per unit frame, triangulate (t1_left, t1_right) of the filtered ORB list (lists 1, 2 of d_orb_pts) and of the
// filtered LK list (lists 0, 1 of d_lkf_pts)
__global__ void __launch_bounds__(128) k_triangulate_frontend(Proj P, const float* __restrict__ pts, const int* __restrict__ counts, int cap, int nlists,
                                                              int la, int lb, int first, int ring, float* __restrict__ out3) {
    const int f = blockIdx.y;
    const int cur_u = first + f, cur = cur_u >= ring ? cur_u - ring : cur_u;
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= counts[cur]) return;
    const float* a = pts + ((size_t)cur * nlists + la) * cap * 2;
    const float* b = pts + ((size_t)cur * nlists + lb) * cap * 2;
    triangulate_one(P, a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1], nullptr, out3 + ((size_t)cur * cap + i) * 3);
}

}  // namespace

int launch_triangulate_frontend(svo_ctx* c, int first, int ring, int n_frames) {
    Proj P;
    for (int i = 0; i < 12; ++i) {
        P.p1[i] = c->proj1[i];
        P.p2[i] = c->proj2[i];
    }
    if (c->fe_mode & SVO_MODE_ORB) {
        dim3 grid((c->kp_cap + 127) / 128, n_frames);
        k_triangulate_frontend<<<grid, 128, 0, c->stream>>>(P, c->d_orb_pts, c->d_norb, c->kp_cap, 3, 1, 2, first, ring, c->d_orb_xyz);
        ++c->launches;
    }
    if (c->fe_mode & SVO_MODE_LK) {
        dim3 grid((c->pt_cap + 127) / 128, n_frames);
        k_triangulate_frontend<<<grid, 128, 0, c->stream>>>(P, c->d_lkf_pts, c->d_nlkf, c->pt_cap, 5, 0, 1, first, ring, c->d_lkf_xyz);
        ++c->launches;
    }
    SVO_CUDA_TRY(c, cudaGetLastError());
    return SVO_OK;
}

extern "C" int svo_triangulate_points(svo_ctx* c, const double* proj1, const double* proj2, const float* pts1, const float* pts2, int n,
                                      float* points4d, float* points3d) {
    if (!c) return SVO_ERR_INVALID;
    if (!proj1 || !proj2 || n < 0 || (n > 0 && (!pts1 || !pts2 || !points3d))) return c->fail(SVO_ERR_INVALID, "svo_triangulate_points: bad arguments");
    if (n == 0) return SVO_OK;
    cudaSetDevice(c->device);
    const size_t pb = svo_align((size_t)n * 8, 256), ob4 = svo_align((size_t)n * 16, 256), ob3 = svo_align((size_t)n * 12, 256);
    int rc = svo_ensure_scratch(c, 2 * pb + ob4 + ob3);
    if (rc != SVO_OK) return rc;
    float* d1 = (float*)c->d_scratch;
    float* d2 = (float*)((uint8_t*)d1 + pb);
    float* d4 = (float*)((uint8_t*)d2 + pb);
    float* d3 = (float*)((uint8_t*)d4 + ob4);
    SVO_CUDA_TRY(c, cudaMemcpyAsync(d1, pts1, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    SVO_CUDA_TRY(c, cudaMemcpyAsync(d2, pts2, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    Proj P;
    for (int i = 0; i < 12; ++i) {
        P.p1[i] = proj1[i];
        P.p2[i] = proj2[i];
    }
    k_triangulate<<<(n + 127) / 128, 128, 0, c->stream>>>(P, d1, d2, n, d4, d3);
    ++c->launches;
    SVO_CUDA_TRY(c, cudaGetLastError());
    if (points4d) SVO_CUDA_TRY(c, cudaMemcpyAsync(points4d, d4, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream));
    SVO_CUDA_TRY(c, cudaMemcpyAsync(points3d, d3, (size_t)n * 12, cudaMemcpyDeviceToHost, c->stream));
    SVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SVO_OK;
}
