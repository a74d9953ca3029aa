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

// Each thread produces 4 horizontally adjacent pixels of one row; a CTA covers 256 x 4 destination pixels.
__global__ void __launch_bounds__(256) k_resize_level(LevelGeom S, LevelGeom D, SlotMap map) {
    const int slot = map.slot(blockIdx.z);
    const int dx0 = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4;
    const int dy = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (dx0 >= D.w || dy >= D.h) return;
    const uint8_t* __restrict__ src = S.base + (size_t)slot * S.plane;
    uint8_t* __restrict__ dst = D.base + (size_t)slot * D.plane;

    const int4 xo = *reinterpret_cast<const int4*>(D.xofs + dx0);
    const int4 al = *reinterpret_cast<const int4*>(D.alpha + 2 * dx0);  // 8 shorts
    const int2 yo = *reinterpret_cast<const int2*>(D.yofs + 2 * dy);
    const int bt = *reinterpret_cast<const int*>(D.beta + 2 * dy);
    const int b0 = (short)(bt & 0xFFFF), b1 = (short)(bt >> 16);
    const uint8_t* r0 = src + (size_t)yo.x * S.pitch;
    const uint8_t* r1 = src + (size_t)yo.y * S.pitch;
    const int sx[4] = {xo.x, xo.y, xo.z, xo.w};
    const int aw[4] = {al.x, al.y, al.z, al.w};
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int a0 = (short)(aw[i] & 0xFFFF), a1 = (short)(aw[i] >> 16);
        // a1 == 0 whenever sx is the last column, so the +1 read never contributes past the row
        const int t0 = __ldg(r0 + sx[i]) * a0 + __ldg(r0 + sx[i] + 1) * a1;
        const int t1 = __ldg(r1 + sx[i]) * a0 + __ldg(r1 + sx[i] + 1) * a1;
        int v = (((b0 * (t0 >> 4)) >> 16) + ((b1 * (t1 >> 4)) >> 16) + 2) >> 2;
        v = min(255, max(0, v));
        out |= (uint32_t)v << (8 * i);
    }
    // pitch is a multiple of 64, so the 4-byte store stays inside the row even when dx0+3 >= w
    *reinterpret_cast<uint32_t*>(dst + (size_t)dy * D.pitch + dx0) = out;
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
            dim3 grid((D.w + 255) / 256, (D.h + 3) / 4, n_img);
            k_resize_level<<<grid, 256, 0, c->stream>>>(S, D, map);
        }
        ++c->launches;
    }
    SVO_CUDA_TRY(c, cudaGetLastError());
    return SVO_OK;
}
