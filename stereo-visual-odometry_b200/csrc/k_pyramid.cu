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
#include <cuda.h>

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
    const int fr = first + f, slot = side * ring + (fr >= ring ? fr - ring : fr);
    uint8_t* dst = L0.base + (size_t)slot * L0.plane + (size_t)y * L0.pitch;
    // 16 destination bytes per thread: the source row starts at an arbitrary byte, so five aligned words are funnel-shifted
    // into four (the pitched row has room for the last partial vector; the row's last, partial source word is read byte-wise)
    const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(src) & 3u), sh = mis * 8u;
    const uint32_t* __restrict__ sw = reinterpret_cast<const uint32_t*>(src - mis);
    const int nwords = (L0.w + (int)mis + 3) >> 2;  // aligned source words that hold bytes of this row
    const int tailb = ((L0.w + (int)mis - 1) & 3) + 1;  // valid bytes of the last of them
    for (int x = threadIdx.x * 16; x < L0.w; x += 256 * 16) {
        const int w0 = x >> 2;
        uint32_t a[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const int wi = w0 + i;
            if (wi < nwords - 1 || (wi == nwords - 1 && tailb == 4)) {
                a[i] = __ldg(sw + wi);
            } else {
                a[i] = 0u;
                if (wi == nwords - 1)  // last, partial word of the row: byte loads, never past the caller's buffer
                    for (int b = 0; b < tailb; ++b) a[i] |= (uint32_t)__ldg(reinterpret_cast<const uint8_t*>(sw + wi) + b) << (8 * b);
            }
        }
        uint4 v;
        v.x = __funnelshift_r(a[0], a[1], sh);
        v.y = __funnelshift_r(a[1], a[2], sh);
        v.z = __funnelshift_r(a[2], a[3], sh);
        v.w = __funnelshift_r(a[3], a[4], sh);
        *reinterpret_cast<uint4*>(dst + x) = v;  // pitch is a multiple of 64: the vector stays inside the destination row
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K1 fused: all pyramid levels of one tile in ONE pass.  The level-0 region of the tile is staged in shared
// memory by TMA (one cp.async.bulk.tensor.2d per CTA, completion on an mbarrier); level l is then computed from level l-1
// entirely in shared memory (ping-pong between two buffers) and only the pixels the tile owns are written to HBM, so level
// 0 is read once and levels 1..7 are written once.  Arithmetic is identical to k_resize_level (IDP.2A horizontal pass).
// Thread layout per level: thread = (4-pixel group g, row lane), groups aligned to 4 in the level's x coordinate so that
// shared-memory and HBM stores are aligned words; a thread keeps its column coefficients in registers and walks the rows.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256) k_pyramid_fused(const __grid_constant__ CUtensorMap tmap, PyrGeom g, SlotMap map,
                                                       const PyrTile* __restrict__ tiles, int bufB_off) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long mbar;
    const int slot = map.slot(blockIdx.y);
    const PyrTile& T = tiles[blockIdx.x];
    uint8_t* bufA = smem;             // level-0 box (SVO_PYR_BOX_H x SVO_PYR_BOX_W), later the even levels
    uint8_t* bufB = smem + bufB_off;  // odd levels
    const PyrRect r0 = T.lv[0];
    if (r0.x1n <= r0.x0) return;  // empty tile (CTA-uniform)

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(SVO_PYR_BOX_W * SVO_PYR_BOX_H) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(bufA)),
                     "l"(&tmap), "r"((int)r0.x0 & ~15), "r"(slot * g.lvl[0].h + (int)r0.y0), "r"(smem_u32(&mbar))
                     : "memory");
    }
    {  // every thread waits for the TMA bytes (phase 0)
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&mbar)) : "memory");
        }
    }

    const uint8_t* src = bufA;
    int s_orgx = r0.x0 & ~15, s_orgy = r0.y0, s_pitch = SVO_PYR_BOX_W;  // the box starts at a 16-byte aligned column
    for (int l = 1; l < g.nlevels; ++l) {
        const LevelGeom& D = g.lvl[l];
        const PyrRect r = T.lv[l];
        const int nw = r.x1n - r.x0, nh = r.y1n - r.y0;
        if (nw <= 0 || nh <= 0) break;  // deeper levels of this tile are empty too (CTA-uniform)
        uint8_t* dbuf = (l & 1) ? bufB : bufA;
        const int d_orgx = r.x0 & ~3;
        const int ng = (r.x1n - d_orgx + 3) >> 2;             // 4-pixel groups
        const int d_pitch = ng * 4 + 8;
        // small-integer divisions through the float reciprocal (exact here: operands < 2^9, (a + 0.5) / b is never within
        // 2^-8 of an integer): the generic 32-bit division costs ~25 instructions, three times per thread and level
        const float inv_ng = __frcp_rn((float)ng);
        const int nrl = (int)(256.5f * inv_ng);               // row lanes = 256 / ng
        const int rl = (int)(((float)threadIdx.x + 0.5f) * inv_ng), gi = (int)threadIdx.x - rl * ng;
        if (rl < nrl) {
            const int gx = d_orgx + 4 * gi;                   // level x of the group's first pixel
            // column coefficients of the 4 pixels (tables are padded to a multiple of 4 entries)
            const int4 xo = *reinterpret_cast<const int4*>(D.xofs + gx);
            const int4 alv = *reinterpret_cast<const int4*>(D.alpha + 2 * gx);
            const uint32_t al[4] = {(uint32_t)alv.x, (uint32_t)alv.y, (uint32_t)alv.z, (uint32_t)alv.w};
            const int sxs[4] = {xo.x, xo.y, xo.z, xo.w};
            unsigned woff[4], sh[4];
            bool valid[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                valid[i] = gx + i >= r.x0 && gx + i < r.x1n;
                const int rel = valid[i] ? sxs[i] - s_orgx : 0;
                woff[i] = (unsigned)rel & ~3u;
                sh[i] = ((unsigned)rel & 3u) * 8u;
            }
            const bool own_full = gx >= r.x0 && gx + 4 <= r.x1o;
            uint8_t* gdst = D.base + (size_t)slot * D.plane + gx;
            // each row lane owns a contiguous run of rows, so the horizontal pass of the lower source row is reused as the
            // upper row of the next destination row (5 of 6 rows at scale 1.2)
            const int chunk = (int)(((float)(nh + nrl - 1) + 0.5f) * __frcp_rn((float)nrl));  // ceil(nh / nrl)
            const int ya = r.y0 + rl * chunk, yb = min(ya + chunk, (int)r.y1n);
            int cached_row = -1;
            int Tc[4] = {0, 0, 0, 0};
            for (int y = ya; y < yb; ++y) {
                const int2 yo = *reinterpret_cast<const int2*>(D.yofs + 2 * y);
                const int bt = *reinterpret_cast<const int*>(D.beta + 2 * y);
                const int b0 = (short)(bt & 0xFFFF), b1 = (short)(bt >> 16);
                int T0[4], T1[4];
                if (yo.x == cached_row) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) T0[i] = Tc[i];
                } else {
                    const uint8_t* row0 = src + (yo.x - s_orgy) * s_pitch;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t a0 = *reinterpret_cast<const uint32_t*>(row0 + woff[i]), a1 = *reinterpret_cast<const uint32_t*>(row0 + woff[i] + 4);
                        T0[i] = (int)(__dp2a_lo(al[i], __funnelshift_r(a0, a1, sh[i]), 0u) >> 4);  // rows are kept as T >> 4
                    }
                }
                {
                    const uint8_t* row1 = src + (yo.y - s_orgy) * s_pitch;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t c0 = *reinterpret_cast<const uint32_t*>(row1 + woff[i]), c1 = *reinterpret_cast<const uint32_t*>(row1 + woff[i] + 4);
                        T1[i] = (int)(__dp2a_lo(al[i], __funnelshift_r(c0, c1, sh[i]), 0u) >> 4);
                    }
                }
                uint32_t out = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    // pixels outside the needed range read column 0 of the source region: harmless garbage, never used
                    const int v = (((b0 * T0[i]) >> 16) + ((b1 * T1[i]) >> 16) + 2) >> 2;  // valid pixels: always in [0, 255]
                    out |= (uint32_t)(v & 0xFF) << (8 * i);
                    Tc[i] = T1[i];
                }
                cached_row = yo.y;
                *reinterpret_cast<uint32_t*>(dbuf + (y - r.y0) * d_pitch + 4 * gi) = out;
                if (y < r.y1o) {
                    uint8_t* p = gdst + (size_t)y * D.pitch;
                    if (own_full) {
                        *reinterpret_cast<uint32_t*>(p) = out;
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (gx + i >= r.x0 && gx + i < r.x1o) p[i] = (uint8_t)(out >> (8 * i));
                    }
                }
            }
        }
        __syncthreads();
        src = dbuf;
        s_orgx = d_orgx;
        s_orgy = r.y0;
        s_pitch = d_pitch;
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
    if (c->fused_pyr) {
        const int bufA = SVO_PYR_BOX_W * SVO_PYR_BOX_H;
        const size_t smem = (size_t)bufA + c->pyr_bufB_bytes;
        SVO_CUDA_TRY(c, cudaFuncSetAttribute(k_pyramid_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid(c->n_ptiles, n_img);
        k_pyramid_fused<<<grid, 256, smem, c->stream>>>(*reinterpret_cast<const CUtensorMap*>(c->tmap_l0), g, map, c->d_ptiles, bufA);
        ++c->launches;
        SVO_CUDA_TRY(c, cudaGetLastError());
        return SVO_OK;
    }
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
