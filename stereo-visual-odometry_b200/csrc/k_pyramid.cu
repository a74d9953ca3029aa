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
    __shared__ __align__(16) PyrTile sT;  // the tile's rectangles and thread layouts: read once, not once per level from L2
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
    static_assert(sizeof(PyrTile) % 4 == 0 && sizeof(PyrTile) / 4 <= 256, "PyrTile is copied by one word per thread");
    if (threadIdx.x < sizeof(PyrTile) / 4) reinterpret_cast<uint32_t*>(&sT)[threadIdx.x] = __ldg(reinterpret_cast<const uint32_t*>(&T) + threadIdx.x);
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
        const PyrRect r = sT.lv[l];
        if (r.ng <= 0) break;  // deeper levels of this tile are empty too (CTA-uniform)
        uint8_t* dbuf = (l & 1) ? bufB : bufA;
        const int d_orgx = (r.x0 & ~3) - 4;  // level x of byte 0 of a buffer row (one group of front margin)
        const int rl = ((int)threadIdx.x * r.magic) >> 16, gi = (int)threadIdx.x - rl * r.ng;  // row lane, 4-pixel group
        if (rl < r.nrl) {
            const int gx = d_orgx + 4 + 4 * gi;  // level x of the group's first pixel
            const int4 m0 = __ldg(D.xmeta + (gx >> 1)), m1 = __ldg(D.xmeta + (gx >> 1) + 1);
            const int rel0 = m0.x - s_orgx;  // >= 0 (host-checked): source byte of pixel 0's left tap in the buffer row
            const unsigned woff = (unsigned)rel0 & ~3u, sh0 = ((unsigned)rel0 & 3u) * 8u;
            const uint32_t sel1 = (uint32_t)m0.y, sel2 = sel1 >> 8, sel3 = sel1 >> 16;
            const uint32_t al0 = (uint32_t)m0.z, al1 = (uint32_t)m0.w, al2 = (uint32_t)m1.x, al3 = (uint32_t)m1.y;
            const bool store_g = gx >= r.x0 && gx < r.x1o;  // a group is written by the tile that owns its first pixel
            const uint8_t* rowbase = src + woff - s_orgy * s_pitch;  // + sy * s_pitch: the 3 words of source row sy
            // horizontal pass of one source row for the 4 pixels: 3 aligned words hold all 8 taps; pixel i's byte pair is cut out
            // with one PRMT (selector from the host table) and multiplied with its (a0, a1) pair by one IDP.2A; kept as T >> 4
            auto hrow = [&](int sy, uint32_t (&Tq)[4]) {
                const uint32_t* w = reinterpret_cast<const uint32_t*>(rowbase + sy * s_pitch);
                const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
                const uint32_t u0 = __funnelshift_r(w0, w1, sh0), u1 = __funnelshift_r(w1, w2, sh0);
                Tq[0] = __dp2a_lo(al0, u0, 0u) >> 4;
                Tq[1] = __dp2a_lo(al1, __byte_perm(u0, u1, sel1), 0u) >> 4;
                Tq[2] = __dp2a_lo(al2, __byte_perm(u0, u1, sel2), 0u) >> 4;
                Tq[3] = __dp2a_lo(al3, __byte_perm(u0, u1, sel3), 0u) >> 4;
            };
            // each row lane owns a contiguous run of rows, so the horizontal pass of the lower source row is reused as the
            // upper row of the next destination row (5 of 6 rows at scale 1.2)
            const int ya = r.y0 + rl * r.chunk, yb = min(ya + r.chunk, (int)r.y1n);
            const int ystore = store_g ? min(yb, (int)r.y1o) : ya;  // rows [ya, ystore) of this group go to HBM
            uint8_t* sp = dbuf + 4 + 4 * gi + (ya - r.y0) * r.pitch;
            uint8_t* gp = D.base + (size_t)slot * D.plane + gx + (size_t)ya * D.pitch;
            const int4* ymp = D.ymeta + ya;
            int4 ym = ya < yb ? __ldg(ymp) : make_int4(0, 0, 0, 0);  // {top tap row, bottom tap row, beta0, beta1}; the next row's is requested one iteration ahead (the table has a spare entry)
            int cached_row = -1;
            uint32_t Tc[4] = {0, 0, 0, 0};
#pragma unroll 2
            for (int y = ya; y < yb; ++y) {
                const int4 cur = ym;
                ym = __ldg(++ymp);
                uint32_t T0[4], T1[4];
                if (cur.x == cached_row) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) T0[i] = Tc[i];
                } else {
                    hrow(cur.x, T0);
                }
                hrow(cur.y, T1);
                // v = (((b0 * T0) >> 16) + ((b1 * T1) >> 16) + 2) >> 2 for two pixels at a time: the upper halves of the two
                // products are gathered by PRMT, both sums and the rounding constant are one 3-input add (no carry between
                // the halves: each sum is <= 1023), and the bytes of (sum >> 2) are picked by a last PRMT
                const uint32_t b0 = (uint32_t)cur.z, b1 = (uint32_t)cur.w;
                const uint32_t s01 = __byte_perm(b0 * T0[0], b0 * T0[1], 0x7632) + __byte_perm(b1 * T1[0], b1 * T1[1], 0x7632) + 0x00020002u;
                const uint32_t s23 = __byte_perm(b0 * T0[2], b0 * T0[3], 0x7632) + __byte_perm(b1 * T1[2], b1 * T1[3], 0x7632) + 0x00020002u;
                const uint32_t out = __byte_perm(s01 >> 2, s23 >> 2, 0x6420);
#pragma unroll
                for (int i = 0; i < 4; ++i) Tc[i] = T1[i];
                cached_row = cur.y;
                *reinterpret_cast<uint32_t*>(sp) = out;
                sp += r.pitch;
                if (y < ystore) *reinterpret_cast<uint32_t*>(gp) = out;
                gp += D.pitch;
            }
        }
        __syncthreads();
        src = dbuf;
        s_orgx = d_orgx;
        s_orgy = r.y0;
        s_pitch = r.pitch;
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
