// k_pyramid.cu -- K1: the 8-level image pyramid of ORBextractor::ComputePyramid (ORBextractor.cpp:1061-1085).
//
// Level l is cv::resize(level l-1, INTER_LINEAR) in OpenCV's 11-bit fixed point:
//   T[y][x] = S[y][sx]*a0 + S[y][sx+1]*a1                               (int32)
//   D       = (((b0*(T0>>4))>>16) + ((b1*(T1>>4))>>16) + 2) >> 2        (u8)
// with the per-column / per-row offset and coefficient tables built once on the host (svo_layout), exactly as
// cv::resize builds them per call.  The 19-px reflect border the reference adds around every level (:1076-1082) is
// never read on this path and is not materialised (it is produced on the host only when mvImagePyramid is asked for).
//
// Roofline: HBM-bound streaming.  Algorithmic bytes per image = read L(l-1) + write L(l) summed over l = 1..7.
#include "svo_common.cuh"

namespace {

// Each thread produces a 4 x 4 block of destination pixels; a CTA covers 256 x 16.  Per source row the horizontal pass of a
// pixel is ONE IDP.2A: the two source bytes S[sx], S[sx+1] are cut out of an aligned word pair with a funnel shift and
// multiplied with the packed (a0, a1) u16 coefficients.  Consecutive destination rows usually share a source row
// (sy(dy+1) == sy(dy)+1 for scale 1.2 in 5 of 6 rows); its horizontal results are reused from registers.
constexpr int RY = 4;

__device__ __forceinline__ void hpass4(const uint8_t* __restrict__ row, const unsigned (&woff)[4], const unsigned (&sh)[4], const uint32_t (&al)[4],
                                       int (&T)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t wa = __ldg(reinterpret_cast<const uint32_t*>(row + woff[i]));
        const uint32_t wb = __ldg(reinterpret_cast<const uint32_t*>(row + woff[i] + 4));
        T[i] = (int)__dp2a_lo(al[i], __funnelshift_r(wa, wb, sh[i]), 0u);  // S[sx]*a0 + S[sx+1]*a1
    }
}

__global__ void __launch_bounds__(256) k_resize_level(LevelGeom S, LevelGeom D, SlotMap map) {
    const int slot = map.slot(blockIdx.z);
    const int dx0 = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4;
    const int dy0 = (blockIdx.y * 4 + (threadIdx.x >> 6)) * RY;
    if (dx0 >= D.w || dy0 >= D.h) return;
    const uint8_t* __restrict__ src = S.base + (size_t)slot * S.plane;
    uint8_t* __restrict__ dst = D.base + (size_t)slot * D.plane + (size_t)dy0 * D.pitch + dx0;

    const int4 xo = *reinterpret_cast<const int4*>(D.xofs + dx0);
    const int4 alv = *reinterpret_cast<const int4*>(D.alpha + 2 * dx0);  // 4 x (a0 | a1 << 16)
    const uint32_t al[4] = {(uint32_t)alv.x, (uint32_t)alv.y, (uint32_t)alv.z, (uint32_t)alv.w};
    const int sx[4] = {xo.x, xo.y, xo.z, xo.w};
    unsigned woff[4], sh[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        woff[i] = (unsigned)sx[i] & ~3u;  // a1 == 0 whenever sx is the last column, so the byte after the row never contributes
        sh[i] = ((unsigned)sx[i] & 3u) * 8u;
    }
    int cached_row = -1;
    int Tc[4] = {0, 0, 0, 0};  // horizontal pass of source row `cached_row`
#pragma unroll
    for (int r = 0; r < RY; ++r) {
        const int dy = dy0 + r;
        if (dy >= D.h) break;
        const int2 yo = *reinterpret_cast<const int2*>(D.yofs + 2 * dy);
        const int bt = *reinterpret_cast<const int*>(D.beta + 2 * dy);
        const int b0 = (short)(bt & 0xFFFF), b1 = (short)(bt >> 16);
        int T0[4], T1[4];
        if (yo.x == cached_row) {
#pragma unroll
            for (int i = 0; i < 4; ++i) T0[i] = Tc[i];
        } else {
            hpass4(src + (size_t)yo.x * S.pitch, woff, sh, al, T0);
        }
        if (yo.y == yo.x) {
#pragma unroll
            for (int i = 0; i < 4; ++i) T1[i] = T0[i];
        } else {
            hpass4(src + (size_t)yo.y * S.pitch, woff, sh, al, T1);
        }
        uint32_t out = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int v = (((b0 * (T0[i] >> 4)) >> 16) + ((b1 * (T1[i] >> 4)) >> 16) + 2) >> 2;
            v = min(255, max(0, v));
            out |= (uint32_t)v << (8 * i);
            Tc[i] = T1[i];
        }
        cached_row = yo.y;
        // pitch is a multiple of 64, so the 4-byte store stays inside the row even when dx0+3 >= w
        *reinterpret_cast<uint32_t*>(dst + (size_t)r * D.pitch) = out;
    }
}

// exact 2x decimation: cv::resize silently uses INTER_AREA (2x2 mean, round half up)
__global__ void __launch_bounds__(256) k_resize_half(LevelGeom S, LevelGeom D, SlotMap map) {
    const int slot = map.slot(blockIdx.z);
    const int dx = blockIdx.x * 64 + (threadIdx.x & 63);
    const int dy = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (dx >= D.w || dy >= D.h) return;
    const uint8_t* src = S.base + (size_t)slot * S.plane + (size_t)(2 * dy) * S.pitch + 2 * dx;
    D.base[(size_t)slot * D.plane + (size_t)dy * D.pitch + dx] = (uint8_t)((src[0] + src[1] + src[S.pitch] + src[S.pitch + 1] + 2) >> 2);
}

// frontend upload: rows of `stride` bytes (device memory: the H2D staging buffer or the caller's own tensor) -> the pitched
// level-0 planes of the frame ring.  One CTA per (row, frame, side); each thread writes 4 aligned bytes.
__global__ void __launch_bounds__(256) k_repitch(const uint8_t* __restrict__ left, const uint8_t* __restrict__ right, size_t stride,
                                                 LevelGeom L0, int first, int ring) {
    const int y = blockIdx.x, f = blockIdx.y, side = blockIdx.z;
    const uint8_t* __restrict__ src = (side ? right : left) + ((size_t)f * L0.h + y) * stride;
    uint8_t* dst = L0.base + (size_t)(side * ring + (first + f) % ring) * L0.plane + (size_t)y * L0.pitch;
    for (int x = threadIdx.x * 4; x < L0.w; x += 256 * 4) {
        uint32_t v = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (x + i < L0.w) v |= (uint32_t)__ldg(src + x + i) << (8 * i);
        *reinterpret_cast<uint32_t*>(dst + x) = v;
    }
}

}  // namespace

int launch_repitch(svo_ctx* c, const uint8_t* left, const uint8_t* right, size_t stride, int first, int ring, int n_frames) {
    const LevelGeom& L0 = c->geom.lvl[0];
    dim3 grid(L0.h, n_frames, 2);
    k_repitch<<<grid, 256, 0, c->stream>>>(left, right, stride, L0, first, ring);
    ++c->launches;
    SVO_CUDA_TRY(c, cudaGetLastError());
    return SVO_OK;
}

int launch_pyramid(svo_ctx* c, const SlotMap& map, int n_img) {
    StageTimer t(c, 0);
    const PyrGeom& g = c->geom;
    for (int l = 1; l < g.nlevels; ++l) {
        const LevelGeom& S = g.lvl[l - 1];
        const LevelGeom& D = g.lvl[l];
        if (S.w == 2 * D.w && S.h == 2 * D.h) {
            dim3 grid((D.w + 63) / 64, (D.h + 3) / 4, n_img);
            k_resize_half<<<grid, 256, 0, c->stream>>>(S, D, map);
        } else {
            dim3 grid((D.w + 255) / 256, (D.h + 4 * RY - 1) / (4 * RY), n_img);
            k_resize_level<<<grid, 256, 0, c->stream>>>(S, D, map);
        }
        ++c->launches;
    }
    SVO_CUDA_TRY(c, cudaGetLastError());
    return SVO_OK;
}
