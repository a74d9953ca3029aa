// k_fast.cu -- K2 / K2b / K9: FAST-9 corner detection.
//
//  K2  k_fast_cells   the per-cell FAST loop of ORBextractor::ComputeKeyPointsOctTree (ORBextractor.cpp:742-781):
//                     cv::FAST(cell window, iniThFAST, nms) and, if the cell is empty, cv::FAST(.., minThFAST, nms).
//                     Single-pass form: with S(p) = the FAST-9 corner score (largest threshold at
//                     which p is still a corner, cornerScore<16>) and S'(p) = S(p) if S(p) >= minTh else 0,
//                         M(p)      = S'(p) > 0  and  S'(p) > S'(q) for the 8 neighbours q INSIDE the cell's inner region
//                         cell out  = { p : M(p), S'(p) >= iniTh }   if non-empty,   else { p : M(p) }.
//                     One CTA = one tile = one cell row x a run of <= 32 cells.  The tile window is staged in shared
//                     memory with 16-byte loads; a cheap necessary test (opposite ring pairs) rejects most pixels and
//                     survivors are compacted into per-warp queues so that the full arc score runs on dense warps.
//                     The score uses packed s16x2 min/max (VIMNMX3): low half = v - ring, high half = ring - v, so one
//                     instruction stream scores the dark and the bright arcs at once.
//  K2b k_compact      concatenates the per-tile lists of a level in tile order == the reference's cell-row-major order.
//  K9  k_fast_full    cv::FAST(image, 20, true) on the whole frame (Tracking::Detect_OpenCVFASTFeatures,
//                     tracking.cpp:94-113): same score stage on full-width row strips, NMS across strip borders, output
//                     in row-major order.
//
// Roofline: every level pixel is read from HBM once (window halos come from L2); the kernels are ALU-bound (about
// 40 integer instructions per pixel), see DESIGN.md.
#include "svo_common.cuh"

namespace {

constexpr int FT = 256;  // threads per CTA
constexpr int FW = FT / 32;

struct FastSmem {
    int PW, PH;  // pixel window pitch (multiple of 16) / rows
    int SW, SH;  // score map pitch / rows (1-px zero frame)
    int MC, CH, CK;  // max cells per tile, max rows per cell, 32-px chunks per cell row
    __host__ __device__ size_t pix_off() const { return 0; }
    __host__ __device__ size_t sc_off() const { return (size_t)PW * PH; }
    __host__ __device__ size_t mask_off() const { return sc_off() + (((size_t)SW * SH + 15) & ~(size_t)15); }
    __host__ __device__ size_t cell_off() const { return mask_off() + (size_t)MC * CH * CK * 2 * sizeof(uint32_t); }
    __host__ __device__ size_t queue_off() const { return cell_off() + (size_t)(3 * MC + 4) * sizeof(int); }
    __host__ __device__ size_t total() const { return queue_off() + (size_t)FW * 64 * sizeof(uint32_t); }
};

// FAST-9 corner score of the pixel at p (shared memory, row pitch `pw`): max over the 16 arcs of 9 contiguous ring
// pixels of min(v - ring) [dark arc] and of min(ring - v) [bright arc], minus 1 (== cv cornerScore<16>).
__device__ __forceinline__ int fast9_score(const uint8_t* p, int pw) {
    const int v = p[0];
    const uint32_t base = ((uint32_t)((-v) & 0xFFFF) << 16) | (uint32_t)(v + 1);
    uint32_t d[16];
    {
        const int r0 = p[3 * pw], r1 = p[3 * pw + 1], r2 = p[2 * pw + 2], r3 = p[pw + 3], r4 = p[3], r5 = p[-pw + 3],
                  r6 = p[-2 * pw + 2], r7 = p[-3 * pw + 1], r8 = p[-3 * pw], r9 = p[-3 * pw - 1], r10 = p[-2 * pw - 2],
                  r11 = p[-pw - 3], r12 = p[-3], r13 = p[pw - 3], r14 = p[2 * pw - 2], r15 = p[3 * pw - 1];
        const int r[16] = {r0, r1, r2, r3, r4, r5, r6, r7, r8, r9, r10, r11, r12, r13, r14, r15};
#pragma unroll
        for (int k = 0; k < 16; ++k) d[k] = __vadd2(((uint32_t)r[k] * 0x10001u) ^ 0x0000FFFFu, base);  // hi: r-v, lo: v-r
    }
    uint32_t m3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m3[k] = __vimin3_s16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
    uint32_t best = 0x80008000u;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const uint32_t a = __vimin3_s16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
        const uint32_t b = __vimin3_s16x2(m3[k + 1], m3[(k + 4) & 15], m3[(k + 7) & 15]);
        best = __vimax3_s16x2(best, a, b);
    }
    const int lo = (short)(best & 0xFFFF), hi = (short)(best >> 16);
    return max(lo, hi) - 1;
}

// Necessary condition for a corner at threshold t: every 9-arc holds one pixel of each opposite ring pair.
__device__ __forceinline__ bool fast9_pretest(const uint8_t* p, int pw, int t) {
    const int v = p[0];
    const bool a = abs(v - (int)p[3 * pw]) > t || abs(v - (int)p[-3 * pw]) > t;
    const bool b = abs(v - (int)p[3]) > t || abs(v - (int)p[-3]) > t;
    return a && b;
}

__device__ __forceinline__ void load_window(const uint8_t* __restrict__ img, int pitch, int xa, int ya, int rows, int nchunk,
                                            uint8_t* pix, int PW) {
    for (int idx = threadIdx.x; idx < rows * nchunk; idx += FT) {
        const int r = idx / nchunk, ch = idx - r * nchunk;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(img + (size_t)(ya + r) * pitch + xa + 16 * ch));
        *reinterpret_cast<uint4*>(pix + (size_t)r * PW + 16 * ch) = v;
    }
}

// Score stage shared by K2 and K9: scores of the pixels [sx0, sx0+sw) x [sy0, sy0+sh) (window coordinates of the pixel
// array) that reach `thr` are written to sc at (+1,+1); sc must be zero-initialised.
__device__ __forceinline__ void score_region(const uint8_t* pix, int PW, int sx0, int sy0, int sw, int sh, int thr, uint8_t* sc,
                                             int SW, uint32_t* wq) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int qn = 0;
    for (int yy = warp; yy < sh; yy += FW) {
        const uint8_t* row = pix + (size_t)(sy0 + yy) * PW + sx0;
        for (int xb = 0; xb < sw; xb += 32) {
            const int xx = xb + lane;
            const bool pass = xx < sw && fast9_pretest(row + xx, PW, thr);
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, pass);
            if (pass) wq[qn + __popc(m & ((1u << lane) - 1))] = ((uint32_t)yy << 16) | (uint32_t)xx;
            qn += __popc(m);
            __syncwarp();
            if (qn >= 32) {
                qn -= 32;
                const uint32_t e = wq[qn + lane];
                const int ey = (int)(e >> 16), ex = (int)(e & 0xFFFF);
                const int s = fast9_score(pix + (size_t)(sy0 + ey) * PW + sx0 + ex, PW);
                if (s >= thr) sc[(size_t)(ey + 1) * SW + ex + 1] = (uint8_t)s;
                __syncwarp();
            }
        }
    }
    if (lane < qn) {
        const uint32_t e = wq[lane];
        const int ey = (int)(e >> 16), ex = (int)(e & 0xFFFF);
        const int s = fast9_score(pix + (size_t)(sy0 + ey) * PW + sx0 + ex, PW);
        if (s >= thr) sc[(size_t)(ey + 1) * SW + ex + 1] = (uint8_t)s;
    }
}

__global__ void __launch_bounds__(FT) k_fast_cells(PyrGeom g, SlotMap map, const FastTile* __restrict__ tiles, FastSmem L,
                                                   uint32_t* __restrict__ tile_buf, int* __restrict__ tile_cnt, int* __restrict__ flags,
                                                   uint32_t tile_buf_per_slot, int n_tiles) {
    extern __shared__ __align__(16) uint8_t smem[];
    const FastTile t = tiles[blockIdx.x];
    const int slot = map.slot(blockIdx.y);
    const LevelGeom& G = g.lvl[t.level];
    uint8_t* pix = smem + L.pix_off();
    uint8_t* sc = smem + L.sc_off();
    uint32_t* masks = reinterpret_cast<uint32_t*>(smem + L.mask_off());
    int* cellflag = reinterpret_cast<int*>(smem + L.cell_off());
    int* cellcnt = cellflag + L.MC;
    int* cellbase = cellcnt + L.MC;  // MC + 1 entries
    uint32_t* queues = reinterpret_cast<uint32_t*>(smem + L.queue_off());
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int tw = t.x1 - t.x0, th = t.y1 - t.y0;
    const int xa = (t.x0 - 3) & ~15;
    const int nchunk = (t.x1 + 3 - xa + 15) >> 4;
    load_window(G.base + (size_t)slot * G.plane, G.pitch, xa, t.y0 - 3, th + 6, nchunk, pix, L.PW);
    for (int i = threadIdx.x; i < (L.SW * (th + 2) + 15) / 16; i += FT) reinterpret_cast<uint4*>(sc)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();

    score_region(pix, L.PW, t.x0 - xa, 3, tw, th, g.minTh, sc, L.SW, queues + warp * 64);
    __syncthreads();

    // ---- non-maximum suppression clipped to the cell, per-cell threshold selection
    const int wCell = G.wCell;
    for (int cidx = warp; cidx < t.ncells; cidx += FW) {
        const int cx0 = cidx * wCell, cw = min(cx0 + wCell, tw) - cx0;
        uint32_t strongAny = 0;
        int cntAll = 0, cntStr = 0;
        for (int ly = 0; ly < th; ++ly) {
            const uint8_t* s1 = sc + (size_t)(ly + 1) * L.SW + cx0 + 1;
            for (int ck = 0; ck < L.CK; ++ck) {
                const int lx = ck * 32 + lane;
                const int s = lx < cw ? s1[lx] : 0;
                uint32_t mAll = 0, mStr = 0;
                if (__ballot_sync(0xFFFFFFFFu, s > 0)) {
                    bool keep = s > 0;
                    if (keep) {
                        const uint8_t* s0 = s1 - L.SW;
                        const uint8_t* s2 = s1 + L.SW;
                        keep = s > s0[lx] && s > s2[lx];
                        if (lx > 0) keep = keep && s > s0[lx - 1] && s > s1[lx - 1] && s > s2[lx - 1];
                        if (lx < cw - 1) keep = keep && s > s0[lx + 1] && s > s1[lx + 1] && s > s2[lx + 1];
                    }
                    mAll = __ballot_sync(0xFFFFFFFFu, keep);
                    mStr = __ballot_sync(0xFFFFFFFFu, keep && s >= g.iniTh);
                }
                if (lane == 0) {
                    uint32_t* mrow = masks + ((size_t)(cidx * L.CH + ly) * L.CK + ck) * 2;
                    mrow[0] = mAll;
                    mrow[1] = mStr;
                }
                strongAny |= mStr;
                cntAll += __popc(mAll);
                cntStr += __popc(mStr);
            }
        }
        if (lane == 0) {
            cellflag[cidx] = strongAny != 0;
            cellcnt[cidx] = strongAny ? cntStr : cntAll;
        }
    }
    __syncthreads();
    if (warp == 0) {  // exclusive scan over <= 32 cells
        int v = lane < t.ncells ? cellcnt[lane] : 0, incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += n;
        }
        cellbase[lane] = incl - v;
        if (lane == 31) cellbase[32] = incl;
    }
    __syncthreads();
    const int total = cellbase[32];
    uint32_t* out = tile_buf + (size_t)slot * tile_buf_per_slot + t.out_off;
    for (int cidx = warp; cidx < t.ncells; cidx += FW) {
        const int cx0 = cidx * wCell;
        const int which = cellflag[cidx];
        int pos = cellbase[cidx];
        for (int ly = 0; ly < th; ++ly)
            for (int ck = 0; ck < L.CK; ++ck) {
                const uint32_t m = masks[((size_t)(cidx * L.CH + ly) * L.CK + ck) * 2 + which];
                if (!m) continue;
                if ((m >> lane) & 1u) {
                    const int lx = ck * 32 + lane;
                    const int o = pos + __popc(m & ((1u << lane) - 1));
                    const int s = sc[(size_t)(ly + 1) * L.SW + cx0 + lx + 1];
                    if (o < t.out_cap) out[o] = pack_cand(t.x0 + cx0 + lx - SVO_MINB, t.y0 + ly - SVO_MINB, s);
                }
                pos += __popc(m);
            }
    }
    if (threadIdx.x == 0) {
        tile_cnt[(size_t)slot * n_tiles + blockIdx.x] = min(total, t.out_cap);
        if (total > t.out_cap) atomicOr(flags + slot, SVO_FLAG_TILE_OVERFLOW);
    }
}

// K2b: one CTA per (level, image): exclusive scan of the level's tile counts, gather into the candidate array.
__global__ void __launch_bounds__(128) k_compact(PyrGeom g, SlotMap map, const FastTile* __restrict__ tiles,
                                                 const uint32_t* __restrict__ tile_buf, const int* __restrict__ tile_cnt,
                                                 uint32_t* __restrict__ cand, int* __restrict__ ncand, int* __restrict__ flags,
                                                 uint32_t tile_buf_per_slot, uint32_t cand_per_slot, int n_tiles) {
    __shared__ int off[129];
    __shared__ int wsum[4];
    __shared__ int running;
    const int l = blockIdx.x, slot = map.slot(blockIdx.y);
    const LevelGeom& G = g.lvl[l];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    uint32_t* dst = cand + (size_t)slot * cand_per_slot + G.cand_off;
    const uint32_t* src = tile_buf + (size_t)slot * tile_buf_per_slot;
    for (int t0 = 0; t0 < G.tile_count; t0 += 128) {
        const int ti = t0 + threadIdx.x;
        const int v = ti < G.tile_count ? tile_cnt[(size_t)slot * n_tiles + G.tile_begin + ti] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int wbase = 0;
        for (int w = 0; w < warp; ++w) wbase += wsum[w];
        off[threadIdx.x] = wbase + incl - v;
        if (threadIdx.x == 127) off[128] = wbase + incl;
        __syncthreads();
        const int chunk_total = off[128], base = running;
        for (int e = threadIdx.x; e < chunk_total; e += 128) {
            int lo = 0, hi = 127;  // last tile with off <= e
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (off[mid] <= e) lo = mid; else hi = mid - 1;
            }
            const int o = base + e;
            if (o < G.cand_cap) dst[o] = src[tiles[G.tile_begin + t0 + lo].out_off + (e - off[lo])];
        }
        __syncthreads();
        if (threadIdx.x == 0) running = base + chunk_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        ncand[slot * SVO_MAX_LEVELS + l] = min(running, G.cand_cap);
        if (running > G.cand_cap) atomicOr(flags + slot, SVO_FLAG_CAND_OVERFLOW);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K9: full-frame FAST.  Tile = full-width strip of rows; NMS looks across strip borders (scores of one extra row on
// each side are computed), the 3-px image frame scores 0 like in cv::FAST.  Output per strip in row-major order.
// ---------------------------------------------------------------------------------------------------------------
struct FullSmem {
    int PW, PH, SW, SH, NI;  // NI: max (row, 32-px chunk) items per strip
    __host__ __device__ size_t sc_off() const { return (size_t)PW * PH; }
    __host__ __device__ size_t item_off() const { return sc_off() + (((size_t)SW * SH + 15) & ~(size_t)15); }
    __host__ __device__ size_t queue_off() const { return item_off() + (size_t)(2 * NI + 16) * sizeof(int); }
    __host__ __device__ size_t total() const { return queue_off() + (size_t)FW * 64 * sizeof(uint32_t); }
};

__device__ int block_scan_excl_ft(int* a, int n, int* wtmp) {  // in-place exclusive scan, all FT threads call
    const int per = (n + FT - 1) / FT;
    const int b = threadIdx.x * per, e = min(b + per, n);
    int sum = 0;
    for (int i = b; i < e; ++i) sum += a[i];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) wtmp[warp] = incl;
    __syncthreads();
    int wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < FW; ++w) {
        const int v = wtmp[w];
        if (w < warp) wbase += v;
        total += v;
    }
    int run = wbase + incl - sum;
    for (int i = b; i < e; ++i) {
        const int v = a[i];
        a[i] = run;
        run += v;
    }
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(FT) k_fast_full(LevelGeom G, SlotMap map, const FastTile* __restrict__ tiles, FullSmem L, int thr,
                                                  uint32_t* __restrict__ tile_buf, int* __restrict__ tile_cnt, int* __restrict__ flags,
                                                  uint32_t buf_per_slot, int n_tiles) {
    extern __shared__ __align__(16) uint8_t smem[];
    const FastTile t = tiles[blockIdx.x];
    const int slot = map.slot(blockIdx.y);
    uint8_t* pix = smem;
    uint8_t* sc = smem + L.sc_off();
    int* cnt = reinterpret_cast<int*>(smem + L.item_off());
    uint32_t* masks = reinterpret_cast<uint32_t*>(cnt + L.NI);
    int* wtmp = reinterpret_cast<int*>(masks + L.NI);
    uint32_t* queues = reinterpret_cast<uint32_t*>(smem + L.queue_off());
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tw = t.x1 - t.x0, th = t.y1 - t.y0;
    const int ys0 = max(t.y0 - 1, 3), ys1 = min(t.y1 + 1, G.h - 3), nsr = ys1 - ys0;
    const int xa = (t.x0 - 3) & ~15;
    const int nchunk = (t.x1 + 3 - xa + 15) >> 4;
    load_window(G.base + (size_t)slot * G.plane, G.pitch, xa, ys0 - 3, nsr + 6, nchunk, pix, L.PW);
    for (int i = threadIdx.x; i < (L.SW * (th + 2) + 15) / 16; i += FT) reinterpret_cast<uint4*>(sc)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    score_region(pix, L.PW, t.x0 - xa, 3, tw, nsr, thr, sc + (ys0 - t.y0) * L.SW, L.SW, queues + warp * 64);
    __syncthreads();
    const int nck = (tw + 31) >> 5, ni = th * nck;
    for (int item = warp; item < ni; item += FW) {
        const int r = item / nck, ck = item - r * nck;
        const int lx = ck * 32 + lane;
        const uint8_t* s1 = sc + (size_t)(r + 1) * L.SW + 1;
        const int s = lx < tw ? s1[lx] : 0;
        uint32_t m = 0;
        if (__ballot_sync(0xFFFFFFFFu, s > 0)) {
            bool keep = s > 0;
            if (keep) {
                const uint8_t* s0 = s1 - L.SW;
                const uint8_t* s2 = s1 + L.SW;
                keep = s > s0[lx - 1] && s > s0[lx] && s > s0[lx + 1] && s > s1[lx - 1] && s > s1[lx + 1] && s > s2[lx - 1] && s > s2[lx] &&
                       s > s2[lx + 1];
            }
            m = __ballot_sync(0xFFFFFFFFu, keep);
        }
        if (lane == 0) {
            masks[item] = m;
            cnt[item] = __popc(m);
        }
    }
    __syncthreads();
    const int total = block_scan_excl_ft(cnt, ni, wtmp);
    uint32_t* out = tile_buf + (size_t)slot * buf_per_slot + t.out_off;
    for (int item = warp; item < ni; item += FW) {
        const uint32_t m = masks[item];
        if (!((m >> lane) & 1u)) continue;
        const int r = item / nck, ck = item - r * nck;
        const int lx = ck * 32 + lane;
        const int o = cnt[item] + __popc(m & ((1u << lane) - 1));
        if (o < t.out_cap) out[o] = pack_cand(t.x0 + lx, t.y0 + r, sc[(size_t)(r + 1) * L.SW + lx + 1]);
    }
    if (threadIdx.x == 0) {
        tile_cnt[(size_t)slot * n_tiles + blockIdx.x] = min(total, t.out_cap);
        if (total > t.out_cap) atomicOr(flags + slot, SVO_FLAG_FAST_OVERFLOW);
    }
}

// one CTA per image: concatenate the strips' lists (row-major order) into cv::KeyPoint records
__global__ void __launch_bounds__(128) k_compact_full(SlotMap map, const FastTile* __restrict__ tiles, int n_tiles,
                                                      const uint32_t* __restrict__ tile_buf, const int* __restrict__ tile_cnt,
                                                      svo_keypoint* __restrict__ kps, int* __restrict__ nkps, int* __restrict__ flags,
                                                      uint32_t buf_per_slot, int pt_cap) {
    __shared__ int off[129];
    __shared__ int wsum[4];
    __shared__ int running;
    const int slot = map.slot(blockIdx.x);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    svo_keypoint* dst = kps + (size_t)slot * pt_cap;
    const uint32_t* src = tile_buf + (size_t)slot * buf_per_slot;
    for (int t0 = 0; t0 < n_tiles; t0 += 128) {
        const int ti = t0 + threadIdx.x;
        const int v = ti < n_tiles ? tile_cnt[(size_t)slot * n_tiles + ti] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int wbase = 0;
        for (int w = 0; w < warp; ++w) wbase += wsum[w];
        off[threadIdx.x] = wbase + incl - v;
        if (threadIdx.x == 127) off[128] = wbase + incl;
        __syncthreads();
        const int chunk_total = off[128], base = running;
        for (int e = threadIdx.x; e < chunk_total; e += 128) {
            int lo = 0, hi = 127;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (off[mid] <= e) lo = mid; else hi = mid - 1;
            }
            const int o = base + e;
            if (o < pt_cap) {
                const uint32_t c = src[tiles[t0 + lo].out_off + (e - off[lo])];
                dst[o] = svo_keypoint{(float)cand_x(c), (float)cand_y(c), 7.f, -1.f, (float)cand_r(c), 0, -1};
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) running = base + chunk_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        nkps[slot] = min(running, pt_cap);
        if (running > pt_cap) atomicOr(flags + slot, SVO_FLAG_FAST_OVERFLOW);
    }
}

}  // namespace

int launch_fast_full(svo_ctx* c, const SlotMap& map, int n_img, int threshold) {
    StageTimer t(c, 6);
    if (c->n_ftiles == 0) return SVO_OK;
    FullSmem L{};
    for (const FastTile& ft : c->ftiles) {
        const int xa = (ft.x0 - 3) & ~15;
        L.PW = std::max(L.PW, (int)svo_align(ft.x1 + 3 - xa, 16));
        L.PH = std::max(L.PH, ft.y1 - ft.y0 + 8);
        L.SW = std::max(L.SW, (int)svo_align(ft.x1 - ft.x0 + 2, 16));
        L.SH = std::max(L.SH, ft.y1 - ft.y0 + 2);
        L.NI = std::max(L.NI, (ft.y1 - ft.y0) * ((ft.x1 - ft.x0 + 31) / 32));
    }
    const size_t smem = L.total();
    if (smem > 200 * 1024) return c->fail(SVO_ERR_INVALID, "full-frame FAST strip does not fit in shared memory");
    SVO_CUDA_TRY(c, cudaFuncSetAttribute(k_fast_full, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    threshold = std::min(std::max(threshold, 0), 255);
    dim3 grid(c->n_ftiles, n_img);
    k_fast_full<<<grid, FT, smem, c->stream>>>(c->geom.lvl[0], map, c->d_ftiles, L, threshold, c->d_ftile_buf, c->d_ftile_cnt, c->d_flags,
                                               c->ftile_buf_per_slot, c->n_ftiles);
    ++c->launches;
    k_compact_full<<<n_img, 128, 0, c->stream>>>(map, c->d_ftiles, c->n_ftiles, c->d_ftile_buf, c->d_ftile_cnt, c->d_fast_kps, c->d_nfast,
                                                 c->d_flags, c->ftile_buf_per_slot, c->pt_cap);
    ++c->launches;
    SVO_CUDA_TRY(c, cudaGetLastError());
    return SVO_OK;
}

int launch_fast_cells(svo_ctx* c, const SlotMap& map, int n_img) {
    const PyrGeom& g = c->geom;
    {
        StageTimer t(c, 1);
        if (c->n_tiles > 0) {
            FastSmem L{};
            for (const FastTile& ft : c->tiles) {
                const int xa = (ft.x0 - 3) & ~15;
                L.PW = std::max(L.PW, (int)svo_align(ft.x1 + 3 - xa, 16));
                L.PH = std::max(L.PH, ft.y1 - ft.y0 + 6);
                L.SW = std::max(L.SW, (int)svo_align(ft.x1 - ft.x0 + 2, 16));
                L.SH = std::max(L.SH, ft.y1 - ft.y0 + 2);
                L.MC = std::max(L.MC, (int)ft.ncells);
                L.CH = std::max(L.CH, ft.y1 - ft.y0);
                L.CK = std::max(L.CK, (g.lvl[ft.level].wCell + 31) / 32);
            }
            L.MC = 32;  // the cell scan is one warp wide
            const size_t smem = L.total();
            if (smem > 200 * 1024) return c->fail(SVO_ERR_INVALID, "FAST tile does not fit in shared memory");
            SVO_CUDA_TRY(c, cudaFuncSetAttribute(k_fast_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            dim3 grid(c->n_tiles, n_img);
            k_fast_cells<<<grid, FT, smem, c->stream>>>(g, map, c->d_tiles, L, c->d_tile_buf, c->d_tile_cnt, c->d_flags,
                                                        c->tile_buf_per_slot, c->n_tiles);
            ++c->launches;
        }
    }
    {
        StageTimer t(c, 2);
        dim3 grid(g.nlevels, n_img);
        k_compact<<<grid, 128, 0, c->stream>>>(g, map, c->d_tiles, c->d_tile_buf, c->d_tile_cnt, c->d_cand, c->d_ncand, c->d_flags,
                                               c->tile_buf_per_slot, c->cand_per_slot, c->n_tiles);
        ++c->launches;
    }
    SVO_CUDA_TRY(c, cudaGetLastError());
    return SVO_OK;
}
