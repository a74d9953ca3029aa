// k_fast.cu -- K2 / K2b / K9: FAST-9 corner detection.
//
//  K2  k_fast_cells   the per-cell FAST loop of ORBextractor::ComputeKeyPointsOctTree (ORBextractor.cpp:742-781):
//                     cv::FAST(cell window, iniThFAST, nms) and, if the cell is empty, cv::FAST(.., minThFAST, nms).
//                     Single-pass form: with S(p) = the FAST-9 corner score (largest threshold at
//                     which p is still a corner, cornerScore<16>) and S'(p) = S(p) if S(p) >= minTh else 0,
//                         M(p)      = S'(p) > 0  and  S'(p) > S'(q) for the 8 neighbours q INSIDE the cell's inner region
//                         cell out  = { p : M(p), S'(p) >= iniTh }   if non-empty,   else { p : M(p) }.
//                     One CTA = one tile = one cell row x a run of <= 32 cells, staged in shared memory with 16-byte loads.
//  K2b k_compact      concatenates the per-tile lists of a level in tile order == the reference's cell-row-major order.
//  K9  k_fast_full    cv::FAST(image, 20, true) on the whole frame (Tracking::Detect_OpenCVFASTFeatures,
//                     tracking.cpp:94-113): the same score stage on full-width row strips, NMS across strip borders,
//                     output in row-major order.
//
// Score stage (shared by K2 and K9), three steps per CTA:
//   1. byte-SIMD pre-test, 4 pixels per thread: |centre - ring| for the 4 compass ring pixels with VABSDIFF4 on 32-bit
//      words of the staged window, thresholded per byte with one shift/and/add.  A pixel can be a corner at threshold t
//      only if one pixel of each opposite ring pair differs by more than t (necessary, deliberately conservative by the
//      halving); ~91 % of pixels stop here.
//   2. survivors are compacted into per-warp queues (ballot + popc) so that the exact score runs on dense warps.
//   3. exact score with packed s16x2 arithmetic: ring * 0xFFFF0001 gives (-ring, +ring) in the two halves, one VIADD.16x2
//      adds (+v, -v), and VIMNMX3.S16x2 min/max trees score the dark and the bright arcs in one instruction stream.
//      Non-zero scores also set a bit in a per-row bitmap, so non-maximum suppression, the per-cell threshold selection
//      and the ordered emission only ever touch the ~4 % of pixels that are corners (one thread per bitmap word).
// Ordered output without atomics on the order: per-cell (K2) / per-word (K9) keep masks, a scan, then positional writes.
//
// Roofline: every level pixel is read from HBM once (window halos come from L2); the kernels are integer-ALU bound.
#include "svo_common.cuh"

namespace {

constexpr int FT = 256;  // threads per CTA
constexpr int FW = FT / 32;
constexpr int QCAP = 160;  // per-warp queue of pre-test survivors: up to 31 left over + 128 new per pre-test iteration

// FAST-9 corner score of the pixel at p (shared memory, row pitch `pw`): max over the 16 arcs of 9 contiguous ring
// pixels of min(v - ring) [dark arc] and of min(ring - v) [bright arc], minus 1 (== cv cornerScore<16>).
__device__ __forceinline__ int fast9_score(const uint8_t* p, int pw) {
    const int v = p[0];
    const uint32_t base = ((uint32_t)v << 16) | (uint32_t)((-v) & 0xFFFF);  // hi: +v, lo: -v
    uint32_t d[16];
    {
        const int r0 = p[3 * pw], r1 = p[3 * pw + 1], r2 = p[2 * pw + 2], r3 = p[pw + 3], r4 = p[3], r5 = p[-pw + 3],
                  r6 = p[-2 * pw + 2], r7 = p[-3 * pw + 1], r8 = p[-3 * pw], r9 = p[-3 * pw - 1], r10 = p[-2 * pw - 2],
                  r11 = p[-pw - 3], r12 = p[-3], r13 = p[pw - 3], r14 = p[2 * pw - 2], r15 = p[3 * pw - 1];
        const int r[16] = {r0, r1, r2, r3, r4, r5, r6, r7, r8, r9, r10, r11, r12, r13, r14, r15};
        // r * 0xFFFF0001 = r - (r << 16): hi half = -r, lo half = +r (no borrow); + base -> hi: v - r, lo: r - v
#pragma unroll
        for (int k = 0; k < 16; ++k) d[k] = __vadd2((uint32_t)r[k] * 0xFFFF0001u, base);
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

__device__ __forceinline__ void load_window(const uint8_t* __restrict__ img, int pitch, int xa, int ya, int rows, int nchunk,
                                            uint8_t* pix, int PW) {
    for (int idx = threadIdx.x; idx < rows * nchunk; idx += FT) {
        const int r = idx / nchunk, ch = idx - r * nchunk;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(img + (size_t)(ya + r) * pitch + xa + 16 * ch));
        *reinterpret_cast<uint4*>(pix + (size_t)r * PW + 16 * ch) = v;
    }
}

__device__ __forceinline__ void score_pixel(const uint8_t* pix, int PW, int sx0, int sy0, uint32_t e, int thr, uint8_t* sc, int SW,
                                            uint32_t* nz, int NWB) {
    const int ey = (int)(e >> 16), ex = (int)(e & 0xFFFF);
    const int s = fast9_score(pix + (size_t)(sy0 + ey) * PW + sx0 + ex, PW);
    if (s >= thr && s > 0) {
        sc[(size_t)(ey + 1) * SW + ex + 1] = (uint8_t)s;
        atomicOr(nz + (ey + 1) * NWB + (ex >> 5), 1u << (ex & 31));
    }
}

// Score stage: scores >= thr of the pixels [sx0, sx0+sw) x [sy0, sy0+sh) (window coordinates; sx0 >= 3, PW % 4 == 0) are
// written to sc at (+1,+1) and flagged in the row bitmap nz at row +1; sc and nz must be zero-initialised.  Each warp owns
// rows warp, warp+8, ...; pre-test survivors are appended to the warp's queue (slot = prefix over lanes from 3 ballots of
// the per-lane survivor count) and scored 32 at a time, so the exact score always runs on (almost) full warps.
__device__ __forceinline__ void score_region(const uint8_t* pix, int PW, int sx0, int sy0, int sw, int sh, int thr, uint8_t* sc,
                                             int SW, uint32_t* nz, int NWB, uint32_t* wq, uint32_t* tail_q) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t8 = (thr + 1) >> 1;
    const uint32_t K = (uint32_t)(0x80 - t8) * 0x01010101u;
    const int w0 = sx0 >> 2, w1 = (sx0 + sw - 1) >> 2;  // first / last 4-pixel word of the region
    const uint32_t m_first = 0x80808080u << (8 * (sx0 & 3));                  // bytes >= sx0 of word w0
    const uint32_t m_last = 0x80808080u >> (8 * (3 - ((sx0 + sw - 1) & 3)));  // bytes <= last column of word w1
    const uint32_t lt = (1u << lane) - 1;
    int qn = 0;
    for (int yy = warp; yy < sh; yy += FW) {
        const uint32_t* row = reinterpret_cast<const uint32_t*>(pix + (size_t)(sy0 + yy) * PW);
        const uint32_t* rup = reinterpret_cast<const uint32_t*>(pix + (size_t)(sy0 + yy - 3) * PW);
        const uint32_t* rdn = reinterpret_cast<const uint32_t*>(pix + (size_t)(sy0 + yy + 3) * PW);
        for (int wb = w0; wb <= w1; wb += 32) {
            const int wi = wb + lane;
            uint32_t pass = 0;
            if (wi <= w1) {
                const uint32_t c0 = row[wi], cm = row[wi - 1], cp = row[wi + 1];
                const uint32_t dl = __vabsdiffu4(c0, __funnelshift_r(cm, c0, 8));   // pixels x - 3
                const uint32_t dr = __vabsdiffu4(c0, __funnelshift_r(c0, cp, 24));  // pixels x + 3
                const uint32_t du = __vabsdiffu4(c0, rup[wi]), dd = __vabsdiffu4(c0, rdn[wi]);
                const uint32_t ml = ((dl >> 1) & 0x7F7F7F7Fu) + K, mr = ((dr >> 1) & 0x7F7F7F7Fu) + K;
                const uint32_t mu = ((du >> 1) & 0x7F7F7F7Fu) + K, md = ((dd >> 1) & 0x7F7F7F7Fu) + K;
                pass = (ml | mr) & (mu | md) & 0x80808080u;
                if (wi == w0) pass &= m_first;
                if (wi == w1) pass &= m_last;
            }
            const int n = __popc(pass);  // 0..4 survivors in this lane
            const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, n & 1), b1 = __ballot_sync(0xFFFFFFFFu, n & 2), b2 = __ballot_sync(0xFFFFFFFFu, n & 4);
            if (b0 | b1 | b2) {
                if (pass) {
                    int pos = qn + __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt);
                    const uint32_t e0 = ((uint32_t)yy << 16) + (uint32_t)(4 * wi - sx0);
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (pass & (0x80u << (8 * b))) wq[pos++] = e0 + b;
                }
                qn += __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
                __syncwarp();
                while (qn >= 32) {
                    qn -= 32;
                    score_pixel(pix, PW, sx0, sy0, wq[qn + lane], thr, sc, SW, nz, NWB);
                    __syncwarp();
                }
            }
        }
    }
    // leftovers (< 32 per warp) are pooled in the first warp's queue area and scored by full warps after a barrier
    int* tail_n = reinterpret_cast<int*>(tail_q);
    uint32_t* tail = tail_q + 4;
    if (qn) {
        int base = 0;
        if (lane == 0) base = atomicAdd(tail_n, qn);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (lane < qn) tail[base + lane] = wq[lane];
    }
    __syncthreads();
    const int tn = *tail_n;
    for (int i = threadIdx.x; i < tn; i += FT) score_pixel(pix, PW, sx0, sy0, tail[i], thr, sc, SW, nz, NWB);
}

// bits [x0, x0 + n) (n <= 64) of a bitmap row
__device__ __forceinline__ unsigned long long row_bits(const uint32_t* nzrow, int x0, int n) {
    const int w = x0 >> 5, sh = x0 & 31;
    unsigned long long lo = (unsigned long long)nzrow[w] | ((unsigned long long)nzrow[w + 1] << 32);
    unsigned long long m = lo >> sh;
    if (sh) m |= (unsigned long long)nzrow[w + 2] << (64 - sh);
    return n >= 64 ? m : (m & ((1ull << n) - 1ull));
}

struct FastSmem {
    int PW, PH;  // pixel window pitch (multiple of 16) / rows
    int SW, SH;  // score map pitch / rows (1-px zero frame)
    int NWB;     // bitmap words per row (+3 slack words so that row_bits may read past the end)
    int MC, CH;  // max cells per tile, max rows per cell
    // the per-(cell,row) keep masks alias the pixel window, which is dead once the scores exist
    __host__ __device__ size_t mask_bytes() const { return (size_t)MC * CH * 2 * sizeof(unsigned long long); }
    __host__ __device__ size_t pix_bytes() const { return ((size_t)PW * PH > mask_bytes() ? (size_t)PW * PH : mask_bytes() + 15) & ~(size_t)15; }
    __host__ __device__ size_t sc_off() const { return pix_bytes(); }
    __host__ __device__ size_t nz_off() const { return sc_off() + (((size_t)SW * SH + 15) & ~(size_t)15); }
    __host__ __device__ size_t cell_off() const { return nz_off() + (((size_t)NWB * SH * 4 + 15) & ~(size_t)15); }
    __host__ __device__ size_t queue_off() const { return cell_off() + (size_t)(3 * MC + 4) * sizeof(int); }
    __host__ __device__ size_t total() const { return queue_off() + (size_t)(FW * QCAP + 4 + FW * 32) * sizeof(uint32_t); }
};

__global__ void __launch_bounds__(FT) k_fast_cells(PyrGeom g, SlotMap map, const FastTile* __restrict__ tiles, FastSmem L,
                                                   uint32_t* __restrict__ tile_buf, int* __restrict__ tile_cnt, int* __restrict__ flags,
                                                   uint32_t tile_buf_per_slot, int n_tiles) {
    extern __shared__ __align__(16) uint8_t smem[];
    const FastTile t = tiles[blockIdx.x];
    const int slot = map.slot(blockIdx.y);
    const LevelGeom& G = g.lvl[t.level];
    uint8_t* pix = smem;
    uint8_t* sc = smem + L.sc_off();
    uint32_t* nz = reinterpret_cast<uint32_t*>(smem + L.nz_off());
    unsigned long long* masks = reinterpret_cast<unsigned long long*>(smem);  // [cell][row][all, strong]; aliases pix after the score stage
    int* cellflag = reinterpret_cast<int*>(smem + L.cell_off());
    int* cntAll = cellflag + L.MC;
    int* cntStr = cntAll + L.MC;
    uint32_t* queues = reinterpret_cast<uint32_t*>(smem + L.queue_off());
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int tw = t.x1 - t.x0, th = t.y1 - t.y0;
    const int xa = (t.x0 - 3) & ~15;
    const int nchunk = (t.x1 + 3 - xa + 15) >> 4;
    load_window(G.base + (size_t)slot * G.plane, G.pitch, xa, t.y0 - 3, th + 6, nchunk, pix, L.PW);
    {  // zero the score map, the bitmap and the per-cell counters (contiguous regions up to mask_off, then the cell block)
        uint4* z = reinterpret_cast<uint4*>(sc);
        const int nz16 = (int)((L.cell_off() - L.sc_off()) >> 4);
        for (int i = threadIdx.x; i < nz16; i += FT) z[i] = make_uint4(0, 0, 0, 0);
        for (int i = threadIdx.x; i < 3 * L.MC + 4; i += FT) cellflag[i] = 0;
        if (threadIdx.x == 0) queues[FW * QCAP] = 0;
    }
    __syncthreads();

    score_region(pix, L.PW, t.x0 - xa, 3, tw, th, g.minTh, sc, L.SW, nz, L.NWB, queues + warp * QCAP, queues + FW * QCAP);
    __syncthreads();

    // ---- one thread per (cell, row): non-maximum suppression clipped to the cell, per-cell threshold statistics
    const int wCell = G.wCell, nitems = t.ncells * th;
    for (int it = threadIdx.x; it < nitems; it += FT) {
        const int cidx = it / th, ly = it - cidx * th;
        const int cx0 = cidx * wCell, cw = min(cx0 + wCell, tw) - cx0;
        unsigned long long m = row_bits(nz + (ly + 1) * L.NWB, cx0, cw), all = 0, strong = 0;
        const uint8_t* s1 = sc + (size_t)(ly + 1) * L.SW + cx0 + 1;
        while (m) {
            const int lx = __ffsll((long long)m) - 1;
            m &= m - 1;
            const int s = s1[lx];
            const uint8_t* s0 = s1 - L.SW;
            const uint8_t* s2 = s1 + L.SW;
            bool keep = s > s0[lx] && s > s2[lx];
            if (lx > 0) keep = keep && s > s0[lx - 1] && s > s1[lx - 1] && s > s2[lx - 1];
            if (lx < cw - 1) keep = keep && s > s0[lx + 1] && s > s1[lx + 1] && s > s2[lx + 1];
            if (keep) {
                all |= 1ull << lx;
                if (s >= g.iniTh) strong |= 1ull << lx;
            }
        }
        masks[2 * it] = all;
        masks[2 * it + 1] = strong;
        if (all) atomicAdd(&cntAll[cidx], __popcll(all));
        if (strong) {
            atomicAdd(&cntStr[cidx], __popcll(strong));
            cellflag[cidx] = 1;
        }
    }
    __syncthreads();
    // ---- every warp scans the (<= 32) per-cell counts, then owns whole cells: lanes = rows, warp scan of the row counts,
    //      and each lane writes the survivors of its row at their final position (cell-major, row-major inside the cell)
    int total;
    int mybase;
    {
        const int v = lane < t.ncells ? (cellflag[lane] ? cntStr[lane] : cntAll[lane]) : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += n;
        }
        total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        mybase = incl - v;  // lane c holds the base of cell c
    }
    uint32_t* out = tile_buf + (size_t)slot * tile_buf_per_slot + t.out_off;
    for (int cidx = warp; cidx < t.ncells; cidx += FW) {
        const int which = cellflag[cidx], cx0 = cidx * wCell;
        int run = __shfl_sync(0xFFFFFFFFu, mybase, cidx);
        for (int r0 = 0; r0 < th; r0 += 32) {
            const int ly = r0 + lane;
            unsigned long long m = ly < th ? masks[2 * (cidx * th + ly) + which] : 0ull;
            const int c = __popcll(m);
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += n;
            }
            int o = run + incl - c;
            run += __shfl_sync(0xFFFFFFFFu, incl, 31);
            while (m) {
                const int lx = __ffsll((long long)m) - 1;
                m &= m - 1;
                if (o < t.out_cap) out[o] = pack_cand(t.x0 + cx0 + lx - SVO_MINB, t.y0 + ly - SVO_MINB, sc[(size_t)(ly + 1) * L.SW + cx0 + lx + 1]);
                ++o;
            }
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
    int PW, PH, SW, SH, NWB, NI;  // NI: max bitmap words (row x 32-px chunk) of the strip's own rows
    __host__ __device__ size_t sc_off() const { return (size_t)PW * PH; }
    __host__ __device__ size_t nz_off() const { return sc_off() + (((size_t)SW * SH + 15) & ~(size_t)15); }
    __host__ __device__ size_t item_off() const { return nz_off() + (((size_t)NWB * SH * 4 + 15) & ~(size_t)15); }
    __host__ __device__ size_t queue_off() const { return item_off() + (size_t)(2 * NI + 16) * sizeof(int); }
    __host__ __device__ size_t total() const { return queue_off() + (size_t)(FW * QCAP + 4 + FW * 32) * sizeof(uint32_t); }
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
    uint32_t* nz = reinterpret_cast<uint32_t*>(smem + L.nz_off());
    int* cnt = reinterpret_cast<int*>(smem + L.item_off());
    uint32_t* masks = reinterpret_cast<uint32_t*>(cnt + L.NI);
    int* wtmp = reinterpret_cast<int*>(masks + L.NI);
    uint32_t* queues = reinterpret_cast<uint32_t*>(smem + L.queue_off());
    const int warp = threadIdx.x >> 5;
    const int tw = t.x1 - t.x0, th = t.y1 - t.y0;
    const int ys0 = max(t.y0 - 1, 3), ys1 = min(t.y1 + 1, G.h - 3), nsr = ys1 - ys0;
    const int xa = (t.x0 - 3) & ~15;
    const int nchunk = (t.x1 + 3 - xa + 15) >> 4;
    load_window(G.base + (size_t)slot * G.plane, G.pitch, xa, ys0 - 3, nsr + 6, nchunk, pix, L.PW);
    {
        uint4* z = reinterpret_cast<uint4*>(sc);
        const int nz16 = (int)((L.item_off() - L.sc_off()) >> 4);
        for (int i = threadIdx.x; i < nz16; i += FT) z[i] = make_uint4(0, 0, 0, 0);
        if (threadIdx.x == 0) queues[FW * QCAP] = 0;
    }
    __syncthreads();
    // map row of image row y is y - (t.y0 - 1); the score stage writes row (y - ys0) + 1, hence the shift
    const int shift = ys0 - t.y0;
    score_region(pix, L.PW, t.x0 - xa, 3, tw, nsr, thr, sc + shift * L.SW, L.SW, nz + shift * L.NWB, L.NWB, queues + warp * QCAP, queues + FW * QCAP);
    __syncthreads();
    // ---- one thread per bitmap word of the strip's own rows: 3x3 non-maximum suppression
    const int nck = (tw + 31) >> 5, ni = th * nck;
    for (int item = threadIdx.x; item < ni; item += FT) {
        const int r = item / nck, ck = item - r * nck;
        uint32_t m = nz[(r + 1) * L.NWB + ck], keep = 0;
        const uint8_t* s1 = sc + (size_t)(r + 1) * L.SW + ck * 32 + 1;
        while (m) {
            const int b = __ffs((int)m) - 1;
            m &= m - 1;
            const int s = s1[b];
            const uint8_t* s0 = s1 - L.SW;
            const uint8_t* s2 = s1 + L.SW;
            if (s > s0[b - 1] && s > s0[b] && s > s0[b + 1] && s > s1[b - 1] && s > s1[b + 1] && s > s2[b - 1] && s > s2[b] && s > s2[b + 1])
                keep |= 1u << b;
        }
        masks[item] = keep;
        cnt[item] = __popc(keep);
    }
    __syncthreads();
    const int total = block_scan_excl_ft(cnt, ni, wtmp);
    uint32_t* out = tile_buf + (size_t)slot * buf_per_slot + t.out_off;
    for (int item = threadIdx.x; item < ni; item += FT) {
        uint32_t m = masks[item];
        if (!m) continue;
        const int r = item / nck, ck = item - r * nck;
        int o = cnt[item];
        while (m) {
            const int b = __ffs((int)m) - 1;
            m &= m - 1;
            const int lx = ck * 32 + b;
            if (o < t.out_cap) out[o] = pack_cand(t.x0 + lx, t.y0 + r, sc[(size_t)(r + 1) * L.SW + lx + 1]);
            ++o;
        }
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
        L.PW = std::max(L.PW, (int)svo_align(ft.x1 + 3 - xa + 4, 16));
        L.PH = std::max(L.PH, ft.y1 - ft.y0 + 8);
        L.SW = std::max(L.SW, (int)svo_align(ft.x1 - ft.x0 + 2, 16));
        L.SH = std::max(L.SH, ft.y1 - ft.y0 + 2);
        L.NWB = std::max(L.NWB, (ft.x1 - ft.x0 + 31) / 32 + 3);
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
                L.PW = std::max(L.PW, (int)svo_align(ft.x1 + 3 - xa + 4, 16));
                L.PH = std::max(L.PH, ft.y1 - ft.y0 + 6);
                L.SW = std::max(L.SW, (int)svo_align(ft.x1 - ft.x0 + 2, 16));
                L.SH = std::max(L.SH, ft.y1 - ft.y0 + 2);
                L.NWB = std::max(L.NWB, (ft.x1 - ft.x0 + 31) / 32 + 3);
                L.CH = std::max(L.CH, ft.y1 - ft.y0);
                if (g.lvl[ft.level].wCell > 64) return c->fail(SVO_ERR_INVALID, "FAST cells wider than 64 px are not supported");
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
