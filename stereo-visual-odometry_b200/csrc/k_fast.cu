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
#include <cuda.h>

#include <type_traits>

#include "svo_common.cuh"

namespace {

constexpr int FT = 256;  // threads per CTA
constexpr int FW = FT / 32;
constexpr int QCAP = 288;  // per-warp queue of pre-test survivors: up to 31 left over + 256 new per pre-test step

// FAST-9 corner score of the pixel at p (shared memory, row pitch `pw`): max over the 16 arcs of 9 contiguous ring
// pixels of min(v - ring) [dark arc] and of min(ring - v) [bright arc], minus 1 (== cv cornerScore<16>).
__device__ __forceinline__ int fast9_score(const uint8_t* p, int pw) {
    const int v = p[0];
    // biased differences, two per register: hi half = v - ring + 256, lo half = ring - v + 256 (both in [1, 511], so the halves
    // never borrow from each other and ONE 32-bit multiply-add per ring pixel builds both: ring * 0xFFFF0001 = (-ring : +ring))
    const uint32_t base = ((uint32_t)(v + 256) << 16) | (uint32_t)(256 - v);
    uint32_t d[16];
    {
        const int r0 = p[3 * pw], r1 = p[3 * pw + 1], r2 = p[2 * pw + 2], r3 = p[pw + 3], r4 = p[3], r5 = p[-pw + 3],
                  r6 = p[-2 * pw + 2], r7 = p[-3 * pw + 1], r8 = p[-3 * pw], r9 = p[-3 * pw - 1], r10 = p[-2 * pw - 2],
                  r11 = p[-pw - 3], r12 = p[-3], r13 = p[pw - 3], r14 = p[2 * pw - 2], r15 = p[3 * pw - 1];
        const int r[16] = {r0, r1, r2, r3, r4, r5, r6, r7, r8, r9, r10, r11, r12, r13, r14, r15};
#pragma unroll
        for (int k = 0; k < 16; ++k) d[k] = (uint32_t)r[k] * 0xFFFF0001u + base;
    }
    uint32_t m3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m3[k] = __vimin3_s16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
    uint32_t best = 0;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const uint32_t a = __vimin3_s16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
        const uint32_t b = __vimin3_s16x2(m3[k + 1], m3[(k + 4) & 15], m3[(k + 7) & 15]);
        best = __vimax3_s16x2(best, a, b);
    }
    return max((int)(best & 0xFFFF), (int)(best >> 16)) - 257;  // remove the bias, then cornerScore's -1
}

__device__ __forceinline__ void load_window(const uint8_t* __restrict__ img, int pitch, int xa, int ya, int rows, int nchunk,
                                            uint8_t* pix, int PW) {
    for (int idx = threadIdx.x; idx < rows * nchunk; idx += FT) {
        const int r = idx / nchunk, ch = idx - r * nchunk;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(img + (size_t)(ya + r) * pitch + xa + 16 * ch));
        *reinterpret_cast<uint4*>(pix + (size_t)r * PW + 16 * ch) = v;
    }
}

// where the score stage records the pixels with a score >= thr: a row bitmap (K9) or a compact corner list (K2)
struct CornerSink {
    uint32_t* nz;    // bitmap: row (ey + 1) * NWB, bit ex          (nullptr: not used)
    int NWB;
    uint16_t* list;  // corner list entries ey << 10 | ex             (nullptr: not used)
    int* list_n;
    int list_cap;
};

// queue entry of a pre-test survivor: row << 16 | x (32-bit, any strip width) or row << 10 | x (16-bit, tiles up to 1024 x 64)
template <bool E16>
__device__ __forceinline__ void score_pixel(const uint8_t* pix, int PW, int sx0, int sy0, uint32_t e, int thr, uint8_t* sc, int SW,
                                            const CornerSink& sink) {
    const int ey = E16 ? (int)(e >> 10) : (int)(e >> 16), ex = E16 ? (int)(e & 1023) : (int)(e & 0xFFFF);
    const int s = fast9_score(pix + (size_t)(sy0 + ey) * PW + sx0 + ex, PW);
    const bool hit = s >= thr && s > 0;
    if (hit) sc[(size_t)(ey + 1) * SW + ex + 1] = (uint8_t)s;
    if (sink.list) {  // warp-aggregated append (order is irrelevant: the list only feeds the suppression stage)
        const unsigned act = __activemask();
        const unsigned bal = __ballot_sync(act, hit);
        if (bal) {
            const int lane = threadIdx.x & 31, leader = __ffs((int)bal) - 1;
            int base = 0;
            // plain PTX atomic: the aggregation is already done by hand (nvcc would wrap atomicAdd in its own MATCH / REDUX pass)
            if (lane == leader)
                asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(base) : "r"((uint32_t)__cvta_generic_to_shared(sink.list_n)), "r"(__popc(bal)) : "memory");
            base = __shfl_sync(act, base, leader);
            const int o = base + __popc(bal & ((1u << lane) - 1));
            if (hit && o < sink.list_cap) sink.list[o] = (uint16_t)((ey << 10) | ex);
        }
    } else if (hit) {
        atomicOr(sink.nz + (ey + 1) * sink.NWB + (ex >> 5), 1u << (ex & 31));
    }
}

// byte-SIMD necessary condition for 4 pixels (one word c of the centre row): a 9-arc holds one pixel of each opposite ring
// pair, so a corner at threshold t needs |c - left| > t or |c - right| > t, AND the same for up / down.  MSB of each byte.
// POW2 (t + 1 a power of two, e.g. minThFAST = 7): "d >= t + 1" depends only on the high bits of d, so the two differences
// of a pair can be OR-ed before the comparison (exact); otherwise each difference is compared on its own.
template <bool POW2>
__device__ __forceinline__ uint32_t pretest4(uint32_t c, uint32_t cm, uint32_t cp, uint32_t up, uint32_t dn, uint32_t K7) {
    const uint32_t dl = __vabsdiffu4(c, __funnelshift_r(cm, c, 8));   // pixels x - 3
    const uint32_t dr = __vabsdiffu4(c, __funnelshift_r(c, cp, 24));  // pixels x + 3
    const uint32_t du = __vabsdiffu4(c, up), dd = __vabsdiffu4(c, dn);
    if (POW2) {
        const uint32_t h = dl | dr, v = du | dd;
        return (((h & 0x7F7F7F7Fu) + K7) | h) & (((v & 0x7F7F7F7Fu) + K7) | v) & 0x80808080u;
    }
    const uint32_t h = ((dl & 0x7F7F7F7Fu) + K7) | dl | ((dr & 0x7F7F7F7Fu) + K7) | dr;
    const uint32_t v = ((du & 0x7F7F7F7Fu) + K7) | du | ((dd & 0x7F7F7F7Fu) + K7) | dd;
    return h & v & 0x80808080u;
}

// Score stage: scores >= thr of the pixels [sx0, sx0+sw) x [sy0, sy0+sh) (window coordinates; sx0 >= 3, PW % 16 == 0, pix
// 16-byte aligned) are written to sc at (+1,+1) and recorded in `sink`; sc (and the bitmap) must be zero-initialised.  Each
// warp owns rows warp, warp+8, ...; a lane pre-tests 8 pixels (an aligned word pair) per step; survivors are appended to the
// warp's queue (slot = prefix over lanes from 4 ballots of the per-lane survivor count) and scored 32 at a time, so the exact
// score always runs on (almost) full warps.
template <bool POW2, bool E16>
__device__ __forceinline__ void score_region(const uint8_t* pix, int PW, int sx0, int sy0, int sw, int sh, int thr, uint8_t* sc,
                                             int SW, const CornerSink& sink, typename std::conditional<E16, uint16_t, uint32_t>::type* wq,
                                             uint32_t* tail_q) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = min(thr + 1, 128);  // necessary: |difference| >= T
    const uint32_t K7 = (uint32_t)(0x80 - T) * 0x01010101u;
    const int w0 = sx0 >> 2, w1 = (sx0 + sw - 1) >> 2;  // first / last 4-pixel word of the region
    const int w0e = w0 & ~1;
    const uint32_t m_first = 0x80808080u << (8 * (sx0 & 3));                  // bytes >= sx0 of word w0
    const uint32_t m_last = 0x80808080u >> (8 * (3 - ((sx0 + sw - 1) & 3)));  // bytes <= last column of word w1
    volatile int* wcnt = reinterpret_cast<volatile int*>(tail_q) + 4 + warp;  // this warp's queue fill (zeroed by the caller)
    const uint32_t wcnt_addr = (uint32_t)__cvta_generic_to_shared(const_cast<int*>(wcnt));
    // validity masks of a lane's two words (bytes inside [sx0, sx0 + sw)); with a single step per row (tiles of K2) they are the
    // same for every row and are computed once
    auto lane_masks = [&](int wi, uint32_t& k0, uint32_t& k1) {
        k0 = wi < w0 || wi > w1 ? 0u : 0x80808080u;
        k1 = wi + 1 < w0 || wi + 1 > w1 ? 0u : 0x80808080u;
        if (wi == w0) k0 &= m_first;
        if (wi == w1) k0 &= m_last;
        if (wi + 1 == w0) k1 &= m_first;
        if (wi + 1 == w1) k1 &= m_last;
    };
    const bool single = w1 - w0e < 64;
    uint32_t k0 = 0, k1 = 0;
    if (single) lane_masks(w0e + 2 * lane, k0, k1);
    for (int yy = warp; yy < sh; yy += FW) {
        const uint32_t* row = reinterpret_cast<const uint32_t*>(pix + (size_t)(sy0 + yy) * PW);
        const uint32_t* rup = reinterpret_cast<const uint32_t*>(pix + (size_t)(sy0 + yy - 3) * PW);
        const uint32_t* rdn = reinterpret_cast<const uint32_t*>(pix + (size_t)(sy0 + yy + 3) * PW);
        for (int wb = w0e; wb <= w1; wb += 64) {
            const int wi = wb + 2 * lane;  // this lane: words wi, wi + 1
            uint32_t pass0 = 0, pass1 = 0;
            if (!single) lane_masks(wi, k0, k1);
            if (k0 | k1) {
                const uint2 c = *reinterpret_cast<const uint2*>(row + wi);
                const uint2 u = *reinterpret_cast<const uint2*>(rup + wi), d = *reinterpret_cast<const uint2*>(rdn + wi);
                const uint32_t cm = row[wi - 1], cp = row[wi + 2];
                pass0 = pretest4<POW2>(c.x, cm, c.y, u.x, d.x, K7) & k0;
                pass1 = pretest4<POW2>(c.y, c.x, cp, u.y, d.y, K7) & k1;
            }
            // survivors -> the warp's queue.  The kernel is bound by instruction issue, not by the shared-memory pipe, so the
            // slot comes from ONE shared-memory atomic per lane with survivors (order inside the queue is irrelevant) instead
            // of a ballot / popc prefix; plain PTX so that nvcc does not wrap it in its own MATCH / REDUX aggregation.
            const int n = __popc(pass0) + __popc(pass1);  // 0..8 survivors in this lane
            if (n) {
                uint32_t pos;
                asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(pos) : "r"(wcnt_addr), "r"(n) : "memory");
                const uint32_t e0 = ((uint32_t)yy << (E16 ? 10 : 16)) + (uint32_t)(4 * wi - sx0);
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (pass0 & (0x80u << (8 * b))) wq[pos++] = e0 + b;
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (pass1 & (0x80u << (8 * b))) wq[pos++] = e0 + 4 + b;
            }
            __syncwarp();
            int qn = *wcnt;
            if (qn >= 32) {
                do {
                    qn -= 32;
                    score_pixel<E16>(pix, PW, sx0, sy0, wq[qn + lane], thr, sc, SW, sink);
                } while (qn >= 32);
                __syncwarp();
                if (lane == 0) *wcnt = qn;
                __syncwarp();
            }
        }
    }
    const int qn = *wcnt;
    // leftovers (< 32 per warp) are pooled in the first warp's queue area and scored by full warps after a barrier
    int* tail_n = reinterpret_cast<int*>(tail_q);
    uint32_t* tail = tail_q + 4 + FW;
    if (qn) {
        int base = 0;
        if (lane == 0) base = atomicAdd(tail_n, qn);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (lane < qn) tail[base + lane] = wq[lane];
    }
    __syncthreads();
    const int tn = *tail_n;
    for (int i = threadIdx.x; i < tn; i += FT) score_pixel<E16>(pix, PW, sx0, sy0, tail[i], thr, sc, SW, sink);
}

// ---------------------------------------------------------------------------------------------------------------
// K2 (v4).  One CTA = one tile = one cell row x up to 8 cells (<= 250 px wide, <= 64 rows).
//   load     the pixel window (tile + 3 px) arrives by ONE TMA copy: the level plane is described as a 3-D tensor
//            [rows][pitch / 16][16 bytes] with a box of 17 chunks, so the box is 272 bytes wide (a 2-D box is limited to 256)
//            and lands in shared memory with a compile-time pitch -- the 16 ring offsets of the score are immediates.
//   pretest  a lane tests 16 pixels (one LDS.128 per row) and a warp two rows per step; the 16 pass flags are gathered
//            into a bit code with IDP.4A and expanded into queue entries through a 256-entry table (no per-bit branches).
//   score    as before (dense warps from per-warp queues); a hit writes its score into a map whose cells are one column
//            apart (so the suppression needs no cell-border tests) and sets its bit in a [row][cell] 64-bit hit map.
//   select   warp w owns cell w, lane = row: non-maximum suppression of the row's hits, the cell's "any corner >= iniTh"
//            vote, a warp scan of the row counts and the ordered emission all stay in registers of one warp.
// ---------------------------------------------------------------------------------------------------------------
constexpr int PW4 = 272;    // pitch of the pixel window and of the score map
constexpr int MC4 = 8;      // cells per tile (wCell >= 30, tile <= 250 px)
constexpr int QCAP4 = 544;  // per-warp queue: up to 31 left over + 512 new per pre-test step

struct FastTmaps {
    CUtensorMap m[SVO_MAX_LEVELS];
};
struct FastLevel {
    int h;            // level height (rows per slot in the tensor)
    int wCell;
    unsigned magic;   // floor(x / wCell) == umulhi(x, magic) for x * wCell < 2^32
    int box_bytes;    // bytes one TMA box delivers: PW4 * (hCell + 6)
};
struct FastParams {
    FastLevel lv[SVO_MAX_LEVELS];
    int minTh, iniTh;
    int sc_off, hit_off, q_off, ctl_off, tail_off, cl_off, ccap, zero16;  // shared-memory layout (bytes); zero16: uint4 words to clear from sc_off
};

// offsets (3 bits each, ascending) of the set bits of an 8-bit code, count in bits 24..27
struct BitLut {
    uint32_t v[256];
    constexpr BitLut() : v() {
        for (int i = 0; i < 256; ++i) {
            uint32_t w = 0;
            int n = 0;
            for (int b = 0; b < 8; ++b)
                if ((i >> b) & 1) w |= (uint32_t)b << (3 * n++);
            v[i] = w | ((uint32_t)n << 24);
        }
    }
};
__device__ const BitLut kBitLut{};

// FAST-9 score with a compile-time pitch: every ring load is [p + immediate]
template <int PW>
__device__ __forceinline__ int fast9_score_c(const uint8_t* p) {
    const int v = p[0];
    const uint32_t base = ((uint32_t)(v + 256) << 16) | (uint32_t)(256 - v);
    constexpr int off[16] = {3 * PW, 3 * PW + 1, 2 * PW + 2, PW + 3, 3, -PW + 3, -2 * PW + 2, -3 * PW + 1,
                             -3 * PW, -3 * PW - 1, -2 * PW - 2, -PW - 3, -3, PW - 3, 2 * PW - 2, 3 * PW - 1};
    uint32_t d[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) d[k] = (uint32_t)p[off[k]] * 0xFFFF0001u + base;
    uint32_t m3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m3[k] = __vimin3_s16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
    uint32_t best = 0;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const uint32_t a = __vimin3_s16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
        const uint32_t b = __vimin3_s16x2(m3[k + 1], m3[(k + 4) & 15], m3[(k + 7) & 15]);
        best = __vimax3_s16x2(best, a, b);
    }
    return max((int)(best & 0xFFFF), (int)(best >> 16)) - 257;
}

// pre-test of one word: flag (0x01 per byte) of the pixels that can still be corners
template <bool POW2>
__device__ __forceinline__ uint32_t pretest4b(uint32_t c, uint32_t cm, uint32_t cp, uint32_t up, uint32_t dn, uint32_t K7) {
    const uint32_t dl = __vabsdiffu4(c, __funnelshift_r(cm, c, 8));
    const uint32_t dr = __vabsdiffu4(c, __funnelshift_r(c, cp, 24));
    const uint32_t du = __vabsdiffu4(c, up), dd = __vabsdiffu4(c, dn);
    uint32_t x;
    if (POW2) {
        const uint32_t h = dl | dr, v = du | dd;
        x = (((h & 0x7F7F7F7Fu) + K7) | h) & (((v & 0x7F7F7F7Fu) + K7) | v);
    } else {
        const uint32_t h = ((dl & 0x7F7F7F7Fu) + K7) | dl | ((dr & 0x7F7F7F7Fu) + K7) | dr;
        const uint32_t v = ((du & 0x7F7F7F7Fu) + K7) | du | ((dd & 0x7F7F7F7Fu) + K7) | dd;
        x = h & v;
    }
    return (x >> 7) & 0x01010101u;
}

struct Score4Ctx {
    const uint8_t* pix;
    uint8_t* sc;
    int* ccount;       // [MC4] corners per cell
    uint16_t* clist;   // [MC4][ccap] corners of a cell: row << 6 | column in the cell
    int ccap;
    int sx0, wCell, thr;
    unsigned magic;
};

// queue entry: row << 9 | window column
__device__ __forceinline__ void score_entry4(const Score4Ctx& S, uint32_t e) {
    const int ey = (int)(e >> 9), px = (int)(e & 511);
    const int s = fast9_score_c<PW4>(S.pix + (ey + 3) * PW4 + px);
    if (s >= S.thr && s > 0) {
        const int ex = px - S.sx0;
        const int cidx = (int)__umulhi((unsigned)ex, S.magic), lx = ex - cidx * S.wCell;
        S.sc[(ey + 1) * PW4 + ex + cidx + 1] = (uint8_t)s;
        const int o = atomicAdd(S.ccount + cidx, 1);  // order inside the list is irrelevant
        if (o < S.ccap) S.clist[cidx * S.ccap + o] = (uint16_t)((ey << 6) | lx);
    }
}

__device__ __forceinline__ int lowest_bit64(unsigned long long m) {  // index of the lowest set bit (m != 0)
    const uint32_t lo = (uint32_t)m;
    return lo ? __ffs((int)lo) - 1 : 31 + __ffs((int)(uint32_t)(m >> 32));
}

template <bool POW2>
__global__ void __launch_bounds__(FT) k_fast_cells(const __grid_constant__ FastTmaps maps, const __grid_constant__ FastParams P, SlotMap map,
                                                   const FastTile* __restrict__ tiles, uint32_t* __restrict__ tile_buf, int* __restrict__ tile_cnt,
                                                   int* __restrict__ flags, uint32_t tile_buf_per_slot, int n_tiles) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long mbar;
    const FastTile t = tiles[blockIdx.x];
    const int slot = map.slot(blockIdx.y);
    const FastLevel fl = P.lv[t.level];
    uint8_t* pix = smem;
    uint8_t* sc = smem + P.sc_off;
    uint32_t* hit = reinterpret_cast<uint32_t*>(smem + P.hit_off);  // [row][MC4] 64-bit; aliases the queues (dead once the scores exist)
    uint16_t* queues = reinterpret_cast<uint16_t*>(smem + P.q_off);
    int* ctl = reinterpret_cast<int*>(smem + P.ctl_off);  // [0] pooled leftovers, [1..8] queue fills, [16..23] cell totals
    uint16_t* tail = reinterpret_cast<uint16_t*>(smem + P.tail_off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tw = t.x1 - t.x0, th = t.y1 - t.y0;
    const int xa = (t.x0 - 4) & ~15;  // first column of the window: 16-byte aligned (TMA), 4..19 px left of the tile
    const int sx0 = t.x0 - xa;

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&mbar)), "r"(fl.box_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(pix)),
                     "l"(&maps.m[t.level]), "r"(0), "r"(xa >> 4), "r"(slot * fl.h + t.y0 - 3), "r"((uint32_t)__cvta_generic_to_shared(&mbar))
                     : "memory");
    }
    {  // while the copy is in flight: clear the score map, the hit map and the counters (one contiguous region)
        uint4* z = reinterpret_cast<uint4*>(sc);
        for (int i = threadIdx.x; i < P.zero16; i += FT) z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();  // the mbarrier is initialised for everybody, the maps are clear
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done)
                         : "r"((uint32_t)__cvta_generic_to_shared(&mbar))
                         : "memory");
    }

    // ---- pre-test + exact score
    uint16_t* clist = reinterpret_cast<uint16_t*>(smem + P.cl_off);
    const Score4Ctx S{pix, sc, ctl + 16, clist, P.ccap, sx0, fl.wCell, P.minTh, fl.magic};
    {
        const int T = min(P.minTh + 1, 128);  // necessary: |difference| >= T
        const uint32_t K7 = (uint32_t)(0x80 - T) * 0x01010101u;
        uint16_t* wq = queues + warp * QCAP4;
        volatile int* wcnt = ctl + 1 + warp;
        const uint32_t wcnt_addr = (uint32_t)__cvta_generic_to_shared(const_cast<int*>(wcnt));
        // bit i set: window column colb + i lies in the tile, [sx0, sx0 + tw)
        auto valid_bits = [&](int colb) {
            const int lo = min(max(sx0 - colb, 0), 16), hi = min(max(sx0 + tw - colb, 0), 16);
            return ((1u << hi) - 1u) & ~((1u << lo) - 1u);
        };
        // up to 8 survivors of one byte code (table word l: 3-bit offsets, count << 24) -> queue; the common counts are predicated
        // straight-line code, only a code with more than 3 survivors takes the loop
        auto push8 = [&](uint32_t l, uint32_t e0, uint32_t& pos) {
            const int n = (int)(l >> 24);
            if (n > 0) wq[pos] = (uint16_t)(e0 + (l & 7u));
            if (n > 1) wq[pos + 1] = (uint16_t)(e0 + ((l >> 3) & 7u));
            if (n > 2) wq[pos + 2] = (uint16_t)(e0 + ((l >> 6) & 7u));
            if (n > 3) {
                uint32_t r = l >> 9;
#pragma unroll 1
                for (int k = 3; k < n; ++k) {
                    wq[pos + k] = (uint16_t)(e0 + (r & 7u));
                    r >>= 3;
                }
            }
            pos += n;
        };
        // one step of a warp: every active lane tests the 16 pixels of tile row yy from window column colb on, the survivors go
        // to the warp's queue, and full groups of 32 queued pixels are scored at once
        auto step = [&](bool active, int yy, int colb, uint32_t valid) {
            uint32_t code = 0;
            if (active) {
                const uint8_t* rowp = pix + (yy + 3) * PW4 + colb;
                const uint4 c = *reinterpret_cast<const uint4*>(rowp);
                const uint4 u = *reinterpret_cast<const uint4*>(rowp - 3 * PW4), d = *reinterpret_cast<const uint4*>(rowp + 3 * PW4);
                const uint32_t cm = *reinterpret_cast<const uint32_t*>(rowp - 4), cp = *reinterpret_cast<const uint32_t*>(rowp + 16);
                const uint32_t q0 = pretest4b<POW2>(c.x, cm, c.y, u.x, d.x, K7);
                const uint32_t q1 = pretest4b<POW2>(c.y, c.x, c.z, u.y, d.y, K7);
                const uint32_t q2 = pretest4b<POW2>(c.z, c.y, c.w, u.z, d.z, K7);
                const uint32_t q3 = pretest4b<POW2>(c.w, c.z, cp, u.w, d.w, K7);
                const uint32_t lo = __dp4a(q1, 0x80402010u, __dp4a(q0, 0x08040201u, 0u));
                const uint32_t hi = __dp4a(q3, 0x80402010u, __dp4a(q2, 0x08040201u, 0u));
                code = (lo | (hi << 8)) & valid;
            }
            if (code) {
                const uint32_t la = __ldg(&kBitLut.v[code & 255u]), lb = __ldg(&kBitLut.v[code >> 8]);
                uint32_t pos;
                asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(pos) : "r"(wcnt_addr), "r"((la >> 24) + (lb >> 24)) : "memory");
                const uint32_t e0 = ((uint32_t)yy << 9) + (uint32_t)colb;
                push8(la, e0, pos);
                push8(lb, e0 + 8, pos);
            }
            __syncwarp();
            int qn = *wcnt;
            if (qn >= 32) {
                do {
                    qn -= 32;
                    score_entry4(S, wq[qn + lane]);
                } while (qn >= 32);
                __syncwarp();
                if (lane == 0) *wcnt = qn;
                __syncwarp();
            }
        };
        const int cb = (sx0 >> 4) << 4;  // first window column of lane 0: the 16 lanes of a row cover [cb, cb + 256)
        {
            const int half = lane >> 4, colb = cb + 16 * (lane & 15);
            const uint32_t valid = valid_bits(colb);
            for (int rp = warp; 2 * rp < th; rp += FW) step(2 * rp + half < th, 2 * rp + half, colb, valid);
        }
        if (sx0 + tw > cb + 256) {  // (CTA-uniform) a 17th chunk of up to 9 pixels: one warp takes 32 rows of it in one step, lane = row
            const uint32_t valid = valid_bits(cb + 256);
            const int w0 = (int)(blockIdx.x & 7u);
            if (warp == w0) step(lane < th, lane, cb + 256, valid);
            if (th > 32 && warp == ((w0 + 1) & 7)) step(32 + lane < th, 32 + lane, cb + 256, valid);
        }
        const int qn = *wcnt;
        if (qn) {  // leftovers (< 32 per warp) are pooled and scored by full warps after the barrier
            int base = 0;
            if (lane == 0) base = atomicAdd(ctl, qn);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (lane < qn) tail[base + lane] = wq[lane];
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < ctl[0]) score_entry4(S, tail[threadIdx.x]);
    __syncthreads();

    // ---- warp w = cell w.  Dense suppression over the cell's corner list (one lane per corner, 8 neighbours by a max tree; the
    // cells are one zero column apart in the score map, rows 0 and th + 1 are zero), the cell's "any corner >= iniTh" vote,
    // the survivors' bits into a [row][cell] 64-bit map; then lane = row: popcount, warp scan, ordered emission.
    unsigned long long sel[2] = {0ull, 0ull};
    int rowoff[2] = {0, 0};
    int cell_total = 0;
    bool list_overflow = false;
    const int cidx = warp;
    const int cx0 = cidx * fl.wCell;
    const uint8_t* scell = sc + PW4 + cx0 + cidx + 1;  // score of (row 0, column 0) of this cell
    if (cidx < t.ncells) {
        int n = ctl[16 + cidx];
        list_overflow = n > P.ccap;
        n = min(n, P.ccap);
        uint16_t* cl = clist + cidx * P.ccap;
        unsigned any_strong = 0;
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            bool strong = false;
            if (i < n) {
                const int e = cl[i], ey = e >> 6, lx = e & 63;
                const uint8_t* q = scell + ey * PW4 + lx;
                const int s = q[0];
                const int m1 = __vimax3_s32(q[-PW4 - 1], q[-PW4], q[-PW4 + 1]), m2 = __vimax3_s32(q[-1], q[1], q[PW4 - 1]);
                const int m = max(__vimax3_s32(q[PW4], q[PW4 + 1], m1), m2);
                const bool keep = s > m;
                strong = keep && s >= P.iniTh;
                cl[i] = (uint16_t)(e | (keep ? 0x4000 : 0) | (strong ? 0x8000 : 0));
            }
            any_strong |= __ballot_sync(0xFFFFFFFFu, strong);
        }
        const int need = any_strong ? 0xC000 : 0x4000;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch)  // this cell's column of the survivor map (the queues' memory: nobody reads them any more)
            if (32 * ch + lane < th) *reinterpret_cast<unsigned long long*>(hit + (((32 * ch + lane) * MC4 + cidx) << 1)) = 0ull;
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
            const int e = cl[i];
            if ((e & need) == need) {
                const int ey = (e >> 6) & 63, lx = e & 63;
                atomicOr(hit + ((ey * MC4 + cidx) << 1) + (lx >> 5), 1u << (lx & 31));
            }
        }
        __syncwarp();
        int run = 0;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            const int ly = 32 * ch + lane;
            sel[ch] = ly < th ? *reinterpret_cast<const unsigned long long*>(hit + ((ly * MC4 + cidx) << 1)) : 0ull;
            const int c = __popcll(sel[ch]);
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += v;
            }
            rowoff[ch] = run + incl - c;
            run += __shfl_sync(0xFFFFFFFFu, incl, 31);
            if (th <= 32) break;
        }
        cell_total = run;
        if (lane == 0) ctl[24 + cidx] = run;
    }
    __syncthreads();
    int cell_base = 0, total = 0;
#pragma unroll
    for (int k = 0; k < MC4; ++k) {
        const int v = ctl[24 + k];
        if (k < cidx) cell_base += v;
        total += v;
    }
    if (cidx < t.ncells && cell_total) {
        uint32_t* out = tile_buf + (size_t)slot * tile_buf_per_slot + t.out_off;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            const int ly = 32 * ch + lane;
            unsigned long long m = sel[ch];
            int o = cell_base + rowoff[ch];
            const uint8_t* s1 = scell + ly * PW4;
            while (m) {
                const int b = lowest_bit64(m);
                m &= m - 1;
                if (o < t.out_cap) out[o] = pack_cand(t.x0 + cx0 + b - SVO_MINB, t.y0 + ly - SVO_MINB, s1[b]);
                ++o;
            }
        }
    }
    if (list_overflow && lane == 0) atomicOr(flags + slot, SVO_FLAG_TILE_OVERFLOW);
    if (threadIdx.x == 0) {
        tile_cnt[(size_t)slot * n_tiles + blockIdx.x] = min(total, t.out_cap);
        if (total > t.out_cap) atomicOr(flags + slot, SVO_FLAG_TILE_OVERFLOW);
    }
}

// K2b, two launches (one CTA per (level, image) doing both was 3 % of a 4K step: at 8 frames it is 128 CTAs walking ~100 k candidates each):
//   k_compact_scan    one CTA per (level, image): exclusive scan of the level's tile counts -> tile_pos[], ncand, overflow flag
//   k_compact_gather  one CTA per (tile, image): copies the tile's list to its position in the level's candidate array
__global__ void __launch_bounds__(256) k_compact_scan(PyrGeom g, SlotMap map, const int* __restrict__ tile_cnt, int* __restrict__ tile_pos,
                                                      int* __restrict__ ncand, int* __restrict__ flags, int n_tiles) {
    __shared__ int wsum[8];
    __shared__ int running;
    const int l = blockIdx.x, slot = map.slot(blockIdx.y);
    const LevelGeom& G = g.lvl[l];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    const int* cnt = tile_cnt + (size_t)slot * n_tiles + G.tile_begin;
    int* pos = tile_pos + (size_t)slot * n_tiles + G.tile_begin;
    for (int t0 = 0; t0 < G.tile_count; t0 += 256) {
        const int ti = t0 + threadIdx.x;
        const int v = ti < G.tile_count ? cnt[ti] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const int x = wsum[w];
            if (w < warp) wbase += x;
            total += x;
        }
        const int base = running;
        if (ti < G.tile_count) pos[ti] = base + wbase + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) running = base + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        ncand[slot * SVO_MAX_LEVELS + l] = min(running, G.cand_cap);
        if (running > G.cand_cap) atomicOr(flags + slot, SVO_FLAG_CAND_OVERFLOW);
    }
}

__global__ void __launch_bounds__(64) k_compact_gather(PyrGeom g, SlotMap map, const FastTile* __restrict__ tiles, const uint32_t* __restrict__ tile_buf,
                                                       const int* __restrict__ tile_cnt, const int* __restrict__ tile_pos, uint32_t* __restrict__ cand,
                                                       uint32_t tile_buf_per_slot, uint32_t cand_per_slot, int n_tiles) {
    const int slot = map.slot(blockIdx.y);
    const int n = tile_cnt[(size_t)slot * n_tiles + blockIdx.x];
    if (n == 0) return;
    const FastTile t = tiles[blockIdx.x];
    const LevelGeom& G = g.lvl[t.level];
    const int p0 = tile_pos[(size_t)slot * n_tiles + blockIdx.x];
    const uint32_t* src = tile_buf + (size_t)slot * tile_buf_per_slot + t.out_off;
    uint32_t* dst = cand + (size_t)slot * cand_per_slot + G.cand_off;
    for (int e = threadIdx.x; e < n; e += 64)
        if (p0 + e < G.cand_cap) dst[p0 + e] = src[e];
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
    __host__ __device__ size_t total() const { return queue_off() + (size_t)(FW * QCAP + 4 + FW + FW * 32) * sizeof(uint32_t); }
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
    extern __shared__ __align__(128) uint8_t smem[];
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
        if (threadIdx.x < 4 + FW) queues[FW * QCAP + threadIdx.x] = 0;  // pooled-leftover counter and the per-warp queue fills
    }
    __syncthreads();
    // map row of image row y is y - (t.y0 - 1); the score stage writes row (y - ys0) + 1, hence the shift
    const int shift = ys0 - t.y0;
    const CornerSink sink{nz + shift * L.NWB, L.NWB, nullptr, nullptr, 0};
    score_region<false, false>(pix, L.PW, t.x0 - xa, 3, tw, nsr, thr, sc + shift * L.SW, L.SW, sink, queues + warp * QCAP, queues + FW * QCAP);
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
            if (!c->fast_tma) return c->fail(SVO_ERR_CUDA, "FAST tensor maps are not available (cuTensorMapEncodeTiled failed)");
            FastParams P{};
            int PH = 0, TH = 0, ccell = 0;
            for (const FastTile& ft : c->tiles) {
                const LevelGeom& G = g.lvl[ft.level];
                const int xa = (ft.x0 - 4) & ~15;
                if (ft.x1 + 3 - xa > PW4 || ft.ncells > MC4 || ft.y1 - ft.y0 > 64 || G.wCell > 64 || ft.y1 - ft.y0 > G.hCell)
                    return c->fail(SVO_ERR_INVALID, "FAST tile outside the supported geometry (<= 250 x 64 px, <= 8 cells of <= 64 px)");
                PH = std::max(PH, G.hCell + 6);
                TH = std::max(TH, ft.y1 - ft.y0);
                ccell = std::max(ccell, G.wCell * (ft.y1 - ft.y0));
            }
            for (int l = 0; l < g.nlevels; ++l) {
                const LevelGeom& G = g.lvl[l];
                if (G.tile_count == 0) continue;
                P.lv[l] = FastLevel{G.h, G.wCell, 0xFFFFFFFFu / (unsigned)G.wCell + 1u, PW4 * (G.hCell + 6)};
            }
            P.minTh = g.minTh;
            P.iniTh = g.iniTh;
            P.sc_off = (int)svo_align((size_t)PW4 * PH, 128);
            P.ctl_off = P.sc_off + PW4 * (TH + 2);
            P.q_off = P.ctl_off + 32 * (int)sizeof(int);
            P.hit_off = P.q_off;  // the survivor map reuses the queues' memory after the score stage (TH * 64 <= FW * QCAP4 * 2 bytes)
            P.zero16 = (P.q_off - P.sc_off) / 16;
            P.tail_off = P.q_off + FW * QCAP4 * 2;
            P.cl_off = P.tail_off + FT * 2;
            P.ccap = (int)svo_align(std::max(64, ccell / 2), 8);  // corners (score >= minTh) of one cell; overflow -> SVO_ERR_CAPACITY
            const size_t smem = (size_t)P.cl_off + (size_t)MC4 * P.ccap * 2;
            const bool pow2 = g.minTh + 1 <= 128 && ((g.minTh + 1) & g.minTh) == 0;
            auto kern = pow2 ? k_fast_cells<true> : k_fast_cells<false>;
            SVO_CUDA_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            dim3 grid(c->n_tiles, n_img);
            kern<<<grid, FT, smem, c->stream>>>(*reinterpret_cast<const FastTmaps*>(c->tmap_fast), P, map, c->d_tiles, c->d_tile_buf, c->d_tile_cnt,
                                                c->d_flags, c->tile_buf_per_slot, c->n_tiles);
            ++c->launches;
        }
    }
    {
        StageTimer t(c, 2);
        if (c->n_tiles > 0) {
            k_compact_scan<<<dim3(g.nlevels, n_img), 256, 0, c->stream>>>(g, map, c->d_tile_cnt, c->d_tile_pos, c->d_ncand, c->d_flags, c->n_tiles);
            k_compact_gather<<<dim3(c->n_tiles, n_img), 64, 0, c->stream>>>(g, map, c->d_tiles, c->d_tile_buf, c->d_tile_cnt, c->d_tile_pos, c->d_cand,
                                                                            c->tile_buf_per_slot, c->cand_per_slot, c->n_tiles);
            c->launches += 2;
        }
    }
    SVO_CUDA_TRY(c, cudaGetLastError());
    return SVO_OK;
}
