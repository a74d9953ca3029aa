// synth.cpp -- deterministic synthetic stereo sequences.  Host-only, integer-only, part of the product
// library so that tests, bench.py and the C++ host classes all see the same bytes without touching the oracle.
//
//   canvas  : h x (w + 64), filled with 110; K = round(w*h/900) axis-aligned rectangles painted in order, each
//             (x0, y0, rw in [6,90), rh in [6,60), gray in [20,236)) drawn from splitmix64; per-pixel noise in [-8,8);
//             3x3 box blur (replicated border, round to nearest).  The canvas is re-seeded every 16 frames.
//   views   : left_t  = canvas[:, 2t : 2t+w], right_t = canvas[:, 2t+8 : 2t+8+w]   (t = frame % 16: 8-px disparity,
//             2 px/frame motion) plus independent per-view sensor noise in [-3,3] (below every FAST threshold used).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "svo_cuda.h"

namespace {

inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void make_canvas(uint64_t seed, int epoch, int w, int h, std::vector<uint8_t>& out) {
    const int cw = w + 64;
    const uint64_t cseed = splitmix64(seed * 0x100000001B3ull + (uint64_t)epoch);
    std::vector<uint8_t> raw((size_t)cw * h, 110);
    const int K = (int)(((int64_t)w * h + 450) / 900);
    for (int k = 0; k < K; ++k) {
        const uint64_t r0 = splitmix64(cseed + 5ull * k), r1 = splitmix64(cseed + 5ull * k + 1), r2 = splitmix64(cseed + 5ull * k + 2),
                       r3 = splitmix64(cseed + 5ull * k + 3), r4 = splitmix64(cseed + 5ull * k + 4);
        const int x0 = (int)(r0 % (uint64_t)cw), y0 = (int)(r1 % (uint64_t)h);
        const int x1 = std::min(cw, x0 + 6 + (int)(r2 % 84)), y1 = std::min(h, y0 + 6 + (int)(r3 % 54));
        const uint8_t gray = (uint8_t)(20 + (int)(r4 % 216));
        for (int y = y0; y < y1; ++y) std::memset(&raw[(size_t)y * cw + x0], gray, (size_t)(x1 - x0));
    }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < cw; ++x) {
            const int n = (int)(splitmix64(cseed ^ (0xA5A5A5A5ull + (uint64_t)y * 65536ull + (uint64_t)x)) & 15) - 8;
            raw[(size_t)y * cw + x] = (uint8_t)std::min(255, std::max(0, raw[(size_t)y * cw + x] + n));
        }
    out.resize((size_t)cw * h);
    for (int y = 0; y < h; ++y) {
        const int ym = std::max(y - 1, 0), yp = std::min(y + 1, h - 1);
        for (int x = 0; x < cw; ++x) {
            const int xm = std::max(x - 1, 0), xp = std::min(x + 1, cw - 1);
            const int s = raw[(size_t)ym * cw + xm] + raw[(size_t)ym * cw + x] + raw[(size_t)ym * cw + xp] + raw[(size_t)y * cw + xm] +
                          raw[(size_t)y * cw + x] + raw[(size_t)y * cw + xp] + raw[(size_t)yp * cw + xm] + raw[(size_t)yp * cw + x] +
                          raw[(size_t)yp * cw + xp];
            out[(size_t)y * cw + x] = (uint8_t)((s + 4) / 9);
        }
    }
}

void cut_view(const std::vector<uint8_t>& canvas, int cw, int xoff, uint64_t vseed, int w, int h, uint8_t* dst, size_t stride) {
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const uint64_t r = splitmix64(vseed + (uint64_t)y * 65536ull + (uint64_t)x);
            const int n = (int)(r & 3) + (int)((r >> 2) & 3) - 3;
            dst[(size_t)y * stride + x] = (uint8_t)std::min(255, std::max(0, canvas[(size_t)y * cw + xoff + x] + n));
        }
}

}  // namespace

extern "C" int svo_synth_stereo_frame(uint64_t seed, int frame, int w, int h, uint8_t* left, uint8_t* right, size_t stride) {
    if (w <= 0 || h <= 0 || frame < 0 || !left || !right || stride < (size_t)w) return SVO_ERR_INVALID;
    // one-entry cache: consecutive frames of one sequence share a canvas
    static thread_local uint64_t c_seed = ~0ull;
    static thread_local int c_epoch = -1, c_w = 0, c_h = 0;
    static thread_local std::vector<uint8_t> canvas;
    const int epoch = frame / 16, t = frame % 16;
    if (c_seed != seed || c_epoch != epoch || c_w != w || c_h != h) {
        make_canvas(seed, epoch, w, h, canvas);
        c_seed = seed;
        c_epoch = epoch;
        c_w = w;
        c_h = h;
    }
    const uint64_t vs = splitmix64(seed ^ (0xC3C3C3C3ull + (uint64_t)frame * 2ull));
    cut_view(canvas, w + 64, 2 * t, vs, w, h, left, stride);
    cut_view(canvas, w + 64, 2 * t + 8, splitmix64(vs + 1), w, h, right, stride);
    return SVO_OK;
}
