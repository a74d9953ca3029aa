// svo_ctx.cu -- context, HBM layout, host-side tables and the ORB-extractor entry points of the C ABI.
//
// Host-side restatements (tables only; all pixel work is in the k_*.cu kernels):
//   svo_orb_configure  <- ORBextractor::ORBextractor            (ORBextractor.cpp:359-418)
//   svo_layout         <- level sizes of ComputePyramid          (ORBextractor.cpp:1065-1066), the cell grid of
//                         ComputeKeyPointsOctTree (:721-740), nIni/hX of DistributeOctTree (:491-493) and the
//                         coefficient tables cv::resize builds once per call
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

#include <cuda.h>

#include "svo_common.cuh"

static const int8_t kBriefPattern[1024] = {
#include "brief_pattern.inc"
};

size_t svo_align(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int cv_round_f(float v) { return (int)lrintf(v); }


// ---- device allocator with optional guard bands (SVO_GUARD=1; see svo_common.cuh) ------------------------------------------------
namespace {
constexpr size_t kGuard = 4096;
constexpr unsigned char kGuardByte = 0xA5;
struct GuardedBuf { void* raw; size_t bytes; int device; };
std::mutex g_guard_mu;
std::map<void*, GuardedBuf> g_guarded;  // user pointer -> allocation
bool guard_on() {
    static const bool on = [] { const char* e = getenv("SVO_GUARD"); return e && *e && *e != '0'; }();
    return on;
}
}  // namespace

cudaError_t svo_dev_malloc_bytes(void** p, size_t bytes) {
    if (!guard_on()) return cudaMalloc(p, bytes);
    const size_t body = svo_align(std::max<size_t>(bytes, 1), 256);  // the rear band starts at the next 256-byte boundary behind the buffer
    void* raw = nullptr;
    cudaError_t e = cudaMalloc(&raw, body + 2 * kGuard);
    if (e != cudaSuccess) return e;
    if ((e = cudaMemset(raw, kGuardByte, body + 2 * kGuard)) != cudaSuccess) {  // the slack between bytes and body counts as band too
        cudaFree(raw);
        return e;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    *p = static_cast<unsigned char*>(raw) + kGuard;
    std::lock_guard<std::mutex> lk(g_guard_mu);
    g_guarded[*p] = GuardedBuf{raw, bytes, dev};
    return cudaSuccess;
}

cudaError_t svo_dev_free(void* p) {
    if (!guard_on() || !p) return cudaFree(p);
    void* raw = p;
    {
        std::lock_guard<std::mutex> lk(g_guard_mu);
        auto it = g_guarded.find(p);
        if (it != g_guarded.end()) {
            raw = it->second.raw;
            g_guarded.erase(it);
        }
    }
    return cudaFree(raw);
}

// Number of guard-band bytes (front band, slack behind the buffer, rear band) of all live buffers that no longer hold the fill
// value; -1 when the library runs without SVO_GUARD.  Synchronises the devices it reads.
extern "C" long long svo_guard_check(void) {
    if (!guard_on()) return -1;
    std::vector<std::pair<void*, GuardedBuf>> live;
    {
        std::lock_guard<std::mutex> lk(g_guard_mu);
        live.assign(g_guarded.begin(), g_guarded.end());
    }
    int keep = 0;
    cudaGetDevice(&keep);
    long long bad = 0;
    std::vector<unsigned char> h;
    for (const auto& kv : live) {
        const GuardedBuf& b = kv.second;
        const size_t body = svo_align(std::max<size_t>(b.bytes, 1), 256);
        cudaSetDevice(b.device);
        cudaDeviceSynchronize();
        const unsigned char* raw = static_cast<const unsigned char*>(b.raw);
        const size_t rear = body - b.bytes + kGuard;
        h.resize(kGuard + rear);
        if (cudaMemcpy(h.data(), raw, kGuard, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(h.data() + kGuard, raw + kGuard + b.bytes, rear, cudaMemcpyDeviceToHost) != cudaSuccess)
            return -2;
        long long n = 0;
        for (unsigned char v : h) n += v != kGuardByte;
        if (n) fprintf(stderr, "svo_guard_check: %lld guard bytes overwritten around a buffer of %zu bytes\n", n, b.bytes);
        bad += n;
    }
    cudaSetDevice(keep);
    return bad;
}

// Positive control of the guard bands: overruns a scratch buffer of 1000 bytes by one byte in front and two behind on purpose and
// returns what svo_guard_check() sees for it (3 with SVO_GUARD=1, -1 without); the buffer is released again.
extern "C" long long svo_guard_selftest(void) {
    if (!guard_on()) return -1;
    const long long before = svo_guard_check();
    unsigned char* p = nullptr;
    if (before < 0 || svo_dev_malloc(&p, 1000) != cudaSuccess) return -2;
    cudaMemset(p - 1, 0, 1);
    cudaMemset(p + 1000, 0, 2);
    const long long after = svo_guard_check();
    svo_dev_free(p);
    return after < 0 ? after : after - before;
}

extern "C" int svo_abi_version(void) { return SVO_ABI_VERSION; }

extern "C" const char* svo_last_error(const svo_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int svo_ctx_create(svo_ctx** out, int device, int max_w, int max_h, int max_images) {
    if (!out) return SVO_ERR_INVALID;
    *out = nullptr;
    if (max_w <= 0 || max_h <= 0 || max_images <= 0 || max_w > SVO_MAX_DIM + SVO_MINB || max_h > SVO_MAX_DIM + SVO_MINB)
        return SVO_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return SVO_ERR_CUDA;  // no CPU fallback
    if (cudaSetDevice(device) != cudaSuccess) return SVO_ERR_CUDA;
    svo_ctx* c = new svo_ctx();
    c->device = device;
    c->max_w = max_w;
    c->max_h = max_h;
    c->max_images = max_images;
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return SVO_ERR_CUDA;
    }
    c->stream = c->own_stream;
    *out = c;
    return SVO_OK;
}

static void free_layout(svo_ctx* c) {
    void** ptrs[] = {(void**)&c->d_pyr,      (void**)&c->d_tables,   (void**)&c->d_tile_buf, (void**)&c->d_tile_cnt,
                     (void**)&c->d_cand,     (void**)&c->d_ncand,    (void**)&c->d_knode,    (void**)&c->d_lkp,
                     (void**)&c->d_nlkp,     (void**)&c->d_kps,      (void**)&c->d_desc,     (void**)&c->d_nkp,
                     (void**)&c->d_flags,    (void**)&c->d_tiles,    (void**)&c->d_lkpyr,    (void**)&c->d_ftiles,
                     (void**)&c->d_ftile_buf, (void**)&c->d_ftile_cnt, (void**)&c->d_fast_kps, (void**)&c->d_nfast,
                     (void**)&c->d_lk_pts,   (void**)&c->d_lk_status, (void**)&c->d_lk_err,   (void**)&c->d_match,
                     (void**)&c->d_orb_pts,  (void**)&c->d_norb,      (void**)&c->d_lkf_pts,  (void**)&c->d_nlkf,
                     (void**)&c->d_ptiles,   (void**)&c->d_orb_xyz,   (void**)&c->d_lkf_xyz,   (void**)&c->d_tile_pos,  (void**)&c->d_lkderiv};
    for (void** p : ptrs) {
        if (*p) svo_dev_free(*p);
        *p = nullptr;
    }
    svo_pnp_free_frontend(c);
    c->cur_w = c->cur_h = 0;
    c->aux_ready = false;
    c->fe_configured = false;
}

extern "C" void svo_ctx_destroy(svo_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    free_layout(c);
    if (c->d_pattern) svo_dev_free(c->d_pattern);
    if (c->d_scratch) svo_dev_free(c->d_scratch);
    if (c->d_flag_accum) svo_dev_free(c->d_flag_accum);
    if (c->d_pnp_rng) svo_dev_free(c->d_pnp_rng);
    if (c->d_xdesc) svo_dev_free(c->d_xdesc);
    if (c->d_xpop) svo_dev_free(c->d_xpop);
    if (c->d_xslots) svo_dev_free(c->d_xslots);
    if (c->d2h_stream) {
        cudaStreamSynchronize(c->d2h_stream);
        cudaStreamDestroy(c->d2h_stream);
        for (int i = 0; i < 2; ++i) cudaEventDestroy(c->ev_done[i]);
    }
    if (c->copy_stream) {
        cudaStreamSynchronize(c->copy_stream);
        for (int i = 0; i < 2; ++i) {
            if (c->d_stage[i]) svo_dev_free(c->d_stage[i]);
            cudaEventDestroy(c->ev_copy[i]);
            cudaEventDestroy(c->ev_repitch[i]);
        }
        cudaStreamDestroy(c->copy_stream);
    }
    if (c->side_stream) {
        cudaStreamSynchronize(c->side_stream);
        cudaStreamDestroy(c->side_stream);
        cudaEventDestroy(c->ev_fork);
        cudaEventDestroy(c->ev_join);
    }
    if (c->ev) {
        for (int k = 0; k < SVO_TIMING_SLOTS; ++k)
            for (int s = 0; s < SVO_NUM_STAGES; ++s) {
                cudaEventDestroy(c->ev[k][s][0]);
                cudaEventDestroy(c->ev[k][s][1]);
            }
        delete[] c->ev;
    }
    cudaStreamDestroy(c->own_stream);
    delete c;
}

extern "C" int svo_ctx_set_stream(svo_ctx* c, void* s) {
    if (!c) return SVO_ERR_INVALID;
    cudaStreamSynchronize(c->stream);
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return SVO_OK;
}

extern "C" int svo_ctx_synchronize(svo_ctx* c) {
    if (!c) return SVO_ERR_INVALID;
    SVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SVO_OK;
}

extern "C" int64_t svo_ctx_launch_count(const svo_ctx* c) { return c ? c->launches : 0; }

extern "C" int svo_ctx_enable_stage_timing(svo_ctx* c, int on) {
    if (!c) return SVO_ERR_INVALID;
    cudaSetDevice(c->device);
    if (on && !c->ev) {
        c->ev = new cudaEvent_t[SVO_TIMING_SLOTS][SVO_NUM_STAGES][2];
        for (int k = 0; k < SVO_TIMING_SLOTS; ++k)
            for (int s = 0; s < SVO_NUM_STAGES; ++s) {
                SVO_CUDA_TRY(c, cudaEventCreate(&c->ev[k][s][0]));
                SVO_CUDA_TRY(c, cudaEventCreate(&c->ev[k][s][1]));
            }
    }
    c->timing = on != 0;
    c->tslot = 0;
    std::memset(c->ev_used, 0, sizeof(c->ev_used));
    return SVO_OK;
}

extern "C" int svo_ctx_set_timing_slot(svo_ctx* c, int slot) {
    if (!c || slot < 0 || slot >= SVO_TIMING_SLOTS) return SVO_ERR_INVALID;
    c->tslot = slot;
    for (int s = 0; s < SVO_NUM_STAGES; ++s) c->ev_used[slot][s] = false;
    return SVO_OK;
}

extern "C" float svo_ctx_stage_ms(const svo_ctx* c, int slot, int stage) {
    if (!c || !c->ev || slot < 0 || slot >= SVO_TIMING_SLOTS || stage < 0 || stage >= SVO_NUM_STAGES || !c->ev_used[slot][stage]) return -1.f;
    float ms = -1.f;
    if (cudaEventSynchronize(c->ev[slot][stage][1]) != cudaSuccess) return -1.f;
    if (cudaEventElapsedTime(&ms, c->ev[slot][stage][0], c->ev[slot][stage][1]) != cudaSuccess) return -1.f;
    return ms;
}

// ---------------------------------------------------------------------------------------------------------------
// ORBextractor::ORBextractor (ORBextractor.cpp:359-418)
// ---------------------------------------------------------------------------------------------------------------
int svo_orb_configure_impl(svo_ctx* c, int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, bool by_user);

extern "C" int svo_orb_configure(svo_ctx* c, int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
    return svo_orb_configure_impl(c, nfeatures, scaleFactor, nlevels, iniTh, minTh, true);
}

int svo_orb_configure_impl(svo_ctx* c, int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, bool by_user) {
    if (!c) return SVO_ERR_INVALID;
    if (nfeatures < 0 || nlevels < 1 || nlevels > SVO_MAX_LEVELS || !(scaleFactor > 1.f) || iniTh < 0 || iniTh > 255 ||
        minTh < 0 || minTh > iniTh)
        return c->fail(SVO_ERR_INVALID, "svo_orb_configure: need nfeatures>=0, 1<=nlevels<=12, scaleFactor>1, 0<=minTh<=iniTh<=255");
    cudaSetDevice(c->device);
    c->nfeatures = nfeatures;
    c->scaleFactor = scaleFactor;
    c->nlevels = nlevels;
    c->iniTh = iniTh;
    c->minTh = minTh;
    const double sfd = (double)scaleFactor;  // the reference stores scaleFactor in a double member (ORBextractor.h:85)
    c->scale.assign(nlevels, 1.f);
    c->sigma2.assign(nlevels, 1.f);
    for (int i = 1; i < nlevels; ++i) {
        c->scale[i] = (float)(c->scale[i - 1] * sfd);
        c->sigma2[i] = c->scale[i] * c->scale[i];
    }
    c->inv_scale.resize(nlevels);
    c->inv_sigma2.resize(nlevels);
    for (int i = 0; i < nlevels; ++i) {
        c->inv_scale[i] = 1.0f / c->scale[i];
        c->inv_sigma2[i] = 1.0f / c->sigma2[i];
    }
    c->features_per_level.assign(nlevels, 0);
    const float factor = (float)(1.0f / sfd);
    float desired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        c->features_per_level[l] = cv_round_f(desired);
        sum += c->features_per_level[l];
        desired *= factor;
    }
    c->features_per_level[nlevels - 1] = std::max(nfeatures - sum, 0);
    // umax: row extents of the 31x31 circular patch (:402-417)
    {
        const int HP = 15;
        const int vmax = (int)std::floor(HP * std::sqrt(2.f) / 2 + 1), vmin = (int)std::ceil(HP * std::sqrt(2.f) / 2);
        for (int v = 0; v <= vmax; ++v) c->umax[v] = (int)lrint(std::sqrt((double)HP * HP - v * v));
        for (int v = HP, v0 = 0; v >= vmin; --v) {
            while (c->umax[v0] == c->umax[v0 + 1]) ++v0;
            c->umax[v] = v0;
            ++v0;
        }
    }
    if (!c->d_pattern) {
        SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_pattern, sizeof(kBriefPattern)));
        SVO_CUDA_TRY(c, cudaMemcpy(c->d_pattern, kBriefPattern, sizeof(kBriefPattern), cudaMemcpyHostToDevice));
    }
    c->configured = true;
    c->user_configured = by_user;
    free_layout(c);  // geometry depends on the tables
    return SVO_OK;
}

extern "C" int svo_orb_tables(const svo_ctx* c, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int32_t* fpl) {
    if (!c || !c->configured) return SVO_ERR_NOT_CONFIGURED;
    for (int i = 0; i < c->nlevels; ++i) {
        if (scale) scale[i] = c->scale[i];
        if (inv_scale) inv_scale[i] = c->inv_scale[i];
        if (sigma2) sigma2[i] = c->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = c->inv_sigma2[i];
        if (fpl) fpl[i] = c->features_per_level[i];
    }
    return SVO_OK;
}

static void level_dims(const svo_ctx* c, int w, int h, int l, int* lw, int* lh) {
    *lw = cv_round_f((float)w * c->inv_scale[l]);
    *lh = cv_round_f((float)h * c->inv_scale[l]);
}

extern "C" int svo_orb_level_size(const svo_ctx* c, int w, int h, int level, int* lw, int* lh) {
    if (!c || !c->configured) return SVO_ERR_NOT_CONFIGURED;
    if (level < 0 || level >= c->nlevels || !lw || !lh) return SVO_ERR_INVALID;
    level_dims(c, w, h, level, lw, lh);
    return SVO_OK;
}

static int level_kp_cap(const svo_ctx* c, int lw, int lh, int l) {
    const int Wp = lw - 2 * SVO_MINB, Hp = lh - 2 * SVO_MINB;
    if (Wp < 30 || Hp < 30) return 0;
    const int nIni = (int)std::round((float)Wp / (float)Hp);
    return std::max(c->features_per_level[l] + 3, 4 * std::max(nIni, 1)) + 1;
}

extern "C" int svo_orb_capacity(const svo_ctx* c, int w, int h) {
    if (!c || !c->configured) return SVO_ERR_NOT_CONFIGURED;
    int cap = 0;
    for (int l = 0; l < c->nlevels; ++l) {
        int lw, lh;
        level_dims(c, w, h, l, &lw, &lh);
        cap += level_kp_cap(c, lw, lh, l);
    }
    return cap;
}

// ---------------------------------------------------------------------------------------------------------------
// HBM layout for a given image size
// ---------------------------------------------------------------------------------------------------------------
int svo_layout(svo_ctx* c, int w, int h) {
    if (!c->configured) return c->fail(SVO_ERR_NOT_CONFIGURED, "svo_orb_configure() first");
    if (w == c->cur_w && h == c->cur_h && c->d_pyr) return SVO_OK;
    if (w > c->max_w || h > c->max_h) return c->fail(SVO_ERR_INVALID, "image larger than the context's max_w x max_h");
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    free_layout(c);
    const int L = c->nlevels, S = c->max_images;
    PyrGeom& g = c->geom;
    std::memset(&g, 0, sizeof(g));
    g.nlevels = L;
    g.iniTh = c->iniTh;
    g.minTh = c->minTh;

    // ---- pyramid planes
    std::vector<size_t> lvl_off(L);
    size_t total = 0;
    for (int l = 0; l < L; ++l) {
        LevelGeom& G = g.lvl[l];
        level_dims(c, w, h, l, &G.w, &G.h);
        if (G.w < 1 || G.h < 1) return c->fail(SVO_ERR_INVALID, "image too small for the configured number of levels");
        G.pitch = (int)svo_align(G.w, 64);
        G.plane = (uint32_t)G.pitch * G.h;
        lvl_off[l] = total;
        total += svo_align((size_t)G.plane * S, 256);
        G.scale = c->scale[l];
        G.patch_size = (int)(31 * c->scale[l]);
    }
    c->pyr_bytes = total + 4096;  // slack: vector loads may run past the last row
    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_pyr, c->pyr_bytes));
    SVO_CUDA_TRY(c, cudaMemsetAsync(c->d_pyr, 0, c->pyr_bytes, c->stream));
    for (int l = 0; l < L; ++l) g.lvl[l].base = c->d_pyr + lvl_off[l];

    // ---- resize coefficient tables (what cv::resize computes per call)
    {
        std::vector<uint8_t> host;
        std::vector<size_t> off_x(L), off_a(L), off_y(L), off_b(L), off_xm(L), off_ym(L);
        auto reserve = [&](size_t bytes) {
            size_t o = svo_align(host.size(), 16);
            host.resize(o + bytes, 0);
            return o;
        };
        for (int l = 1; l < L; ++l) {
            const int sw = g.lvl[l - 1].w, sh = g.lvl[l - 1].h, dw = g.lvl[l].w, dh = g.lvl[l].h;
            const int dwp = (int)svo_align(dw, 4);
            off_x[l] = reserve(sizeof(int) * dwp);
            off_a[l] = reserve(sizeof(short) * 2 * dwp);
            off_y[l] = reserve(sizeof(int) * 2 * dh);
            off_b[l] = reserve(sizeof(short) * 2 * dh);
            int* xofs = (int*)(host.data() + off_x[l]);
            short* alpha = (short*)(host.data() + off_a[l]);
            int* yofs = (int*)(host.data() + off_y[l]);
            short* beta = (short*)(host.data() + off_b[l]);
            const double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
            for (int dx = 0; dx < dw; ++dx) {
                float fx = (float)((dx + 0.5) * scale_x - 0.5);
                int sx = (int)std::floor(fx);
                fx -= sx;
                if (sx < 0) { fx = 0; sx = 0; }
                if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
                xofs[dx] = sx;
                alpha[2 * dx] = (short)std::min(32767, std::max(-32768, cv_round_f((1.f - fx) * 2048.f)));
                alpha[2 * dx + 1] = (short)std::min(32767, std::max(-32768, cv_round_f(fx * 2048.f)));
            }
            for (int dy = 0; dy < dh; ++dy) {
                float fy = (float)((dy + 0.5) * scale_y - 0.5);
                int sy = (int)std::floor(fy);
                fy -= sy;
                yofs[2 * dy] = std::min(std::max(sy, 0), sh - 1);
                yofs[2 * dy + 1] = std::min(std::max(sy + 1, 0), sh - 1);
                beta[2 * dy] = (short)std::min(32767, std::max(-32768, cv_round_f((1.f - fy) * 2048.f)));
                beta[2 * dy + 1] = (short)std::min(32767, std::max(-32768, cv_round_f(fy * 2048.f)));
            }
        }
        // the K1-fused form of the same tables (see LevelGeom::xmeta / ymeta)
        c->pyr_meta_ok = true;
        for (int l = 1; l < L; ++l) {
            const int dw = g.lvl[l].w, dh = g.lvl[l].h;
            const int ngr = (dw + 3) / 4 + 2;
            off_xm[l] = reserve(sizeof(int) * 8 * ngr);
            off_ym[l] = reserve(sizeof(int) * 4 * (dh + 1));  // + 1: the kernel requests row y + 1 unconditionally
            const int* xofs = (const int*)(host.data() + off_x[l]);
            const short* alpha = (const short*)(host.data() + off_a[l]);
            const int* yofs = (const int*)(host.data() + off_y[l]);
            const short* beta = (const short*)(host.data() + off_b[l]);
            int* xm = (int*)(host.data() + off_xm[l]);
            int* ym = (int*)(host.data() + off_ym[l]);
            for (int gi = 0; gi < ngr; ++gi) {
                int sx[4];
                uint32_t al[4];
                for (int i = 0; i < 4; ++i) {
                    const int dx = std::min(4 * gi + i, dw - 1);  // pixels in the pitch padding repeat the last column
                    sx[i] = xofs[dx];
                    al[i] = (uint32_t)(uint16_t)alpha[2 * dx] | ((uint32_t)(uint16_t)alpha[2 * dx + 1] << 16);
                }
                uint32_t sel = 0;
                for (int i = 1; i < 4; ++i) {
                    const int d = sx[i] - sx[0];
                    if (d < 0 || d > 6) c->pyr_meta_ok = false;  // taps of a group must lie within 8 source bytes (scale < ~1.3)
                    const int dc = std::min(std::max(d, 0), 6);
                    sel |= (uint32_t)(dc | ((dc + 1) << 4)) << (8 * (i - 1));
                }
                xm[8 * gi + 0] = sx[0];
                xm[8 * gi + 1] = (int)sel;
                for (int i = 0; i < 4; ++i) xm[8 * gi + 2 + i] = (int)al[i];
                xm[8 * gi + 6] = xm[8 * gi + 7] = 0;
            }
            for (int dy = 0; dy < dh; ++dy) {
                ym[4 * dy + 0] = yofs[2 * dy];
                ym[4 * dy + 1] = yofs[2 * dy + 1];
                ym[4 * dy + 2] = beta[2 * dy];
                ym[4 * dy + 3] = beta[2 * dy + 1];
                if (beta[2 * dy] < 0 || beta[2 * dy + 1] < 0) c->pyr_meta_ok = false;
            }
        }
        host.resize(svo_align(host.size() + 16, 16), 0);
        SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_tables, host.size()));
        SVO_CUDA_TRY(c, cudaMemcpy(c->d_tables, host.data(), host.size(), cudaMemcpyHostToDevice));
        for (int l = 1; l < L; ++l) {
            g.lvl[l].xofs = (const int*)((uint8_t*)c->d_tables + off_x[l]);
            g.lvl[l].alpha = (const short*)((uint8_t*)c->d_tables + off_a[l]);
            g.lvl[l].yofs = (const int*)((uint8_t*)c->d_tables + off_y[l]);
            g.lvl[l].beta = (const short*)((uint8_t*)c->d_tables + off_b[l]);
            g.lvl[l].xmeta = (const int4*)((uint8_t*)c->d_tables + off_xm[l]);
            g.lvl[l].ymeta = (const int4*)((uint8_t*)c->d_tables + off_ym[l]);
        }
    }

    // ---- one-pass pyramid tiling (K1 fused).  Ownership: a level-l pixel belongs to the tile that owns its top-left source
    // tap at level l-1, so only the +1 taps (right / bottom) fall outside the owned source region; the region a tile NEEDS at
    // level l-1 is its owned region plus that small right/bottom fringe, computed redundantly by the neighbours.
    c->fused_pyr = false;
    c->ptiles.clear();
    if (L >= 2 && c->pyr_meta_ok && !getenv("SVO_PYRAMID_PER_LEVEL")) {
        std::vector<std::vector<int>> hx(L), hy(L);  // host copies of the left / top tap index per destination pixel
        bool ok = true;
        for (int l = 1; l < L && ok; ++l) {
            const int sw = g.lvl[l - 1].w, sh = g.lvl[l - 1].h, dw = g.lvl[l].w, dh = g.lvl[l].h;
            if (sw == 2 * dw && sh == 2 * dh) ok = false;  // INTER_AREA case: per-level kernels only
            const double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
            hx[l].resize(dw);
            hy[l].resize(dh);
            for (int dx = 0; dx < dw; ++dx) {
                int sx = (int)std::floor((float)((dx + 0.5) * scale_x - 0.5));
                hx[l][dx] = std::min(std::max(sx, 0), sw - 1);
            }
            for (int dy = 0; dy < dh; ++dy) {
                int sy = (int)std::floor((float)((dy + 0.5) * scale_y - 0.5));
                hy[l][dy] = std::min(std::max(sy, 0), sh - 1);
            }
            // the scheme needs monotone taps that start at 0 (down-scaling)
            if (hx[l][0] != 0 || hy[l][0] != 0) ok = false;
        }
        auto first_ge = [](const std::vector<int>& v, int val) { return (int)(std::lower_bound(v.begin(), v.end(), val) - v.begin()); };
        for (int TW = 192, TH = 64; ok && c->ptiles.empty(); ) {
            std::vector<PyrTile> tiles;
            bool fits = true;
            int maxB = 0;
            const int w1 = g.lvl[1].w, h1 = g.lvl[1].h;
            for (int ty = 0; ty * TH < h1 && fits; ++ty)
                for (int tx = 0; tx * TW < w1 && fits; ++tx) {
                    PyrTile t;
                    std::memset(&t, 0, sizeof(t));
                    int ox0 = tx * TW, ox1 = std::min(ox0 + TW, w1), oy0 = ty * TH, oy1 = std::min(oy0 + TH, h1);
                    t.lv[1] = PyrRect{(short)ox0, (short)ox1, (short)ox1, (short)oy0, (short)oy1, (short)oy1};
                    for (int l = 2; l < L; ++l) {
                        const PyrRect& p = t.lv[l - 1];
                        const int x0 = first_ge(hx[l], p.x0), x1 = first_ge(hx[l], p.x1o), y0 = first_ge(hy[l], p.y0), y1 = first_ge(hy[l], p.y1o);
                        t.lv[l] = PyrRect{(short)x0, (short)x1, (short)x1, (short)y0, (short)y1, (short)y1};
                    }
                    for (int l = L - 1; l >= 1; --l) {  // needed regions, deepest level first
                        PyrRect& d = t.lv[l];
                        // a 4-pixel group is stored (as one word) by the tile that owns its first pixel, so that tile needs
                        // valid values for the whole group
                        if (d.x1o > d.x0 && d.y1o > d.y0) d.x1n = (short)std::max<int>(d.x1n, std::min((d.x1o + 3) & ~3, g.lvl[l].w));
                        const bool empty = d.x1n <= d.x0 || d.y1n <= d.y0;
                        if (l == 1) {
                            PyrRect& s0 = t.lv[0];
                            if (empty) { s0 = PyrRect{0, 0, 0, 0, 0, 0}; break; }
                            s0.x0 = (short)hx[1][d.x0];
                            s0.y0 = (short)hy[1][d.y0];
                            s0.x1o = s0.x1n = (short)std::min(hx[1][d.x1n - 1] + 2, g.lvl[0].w);
                            s0.y1o = s0.y1n = (short)std::min(hy[1][d.y1n - 1] + 2, g.lvl[0].h);
                            // the TMA box starts at a 16-byte aligned column (hardware requirement for 1-byte elements)
                            if (s0.x1n - (s0.x0 & ~15) > SVO_PYR_BOX_W || s0.y1n - s0.y0 > SVO_PYR_BOX_H) fits = false;
                        } else if (!empty) {
                            PyrRect& sr = t.lv[l - 1];
                            sr.x1n = (short)std::max<int>(sr.x1n, std::min(hx[l][d.x1n - 1] + 2, g.lvl[l - 1].w));
                            sr.y1n = (short)std::max<int>(sr.y1n, std::min(hy[l][d.y1n - 1] + 2, g.lvl[l - 1].h));
                            if (hx[l][d.x0] < sr.x0 || hy[l][d.y0] < sr.y0) fits = false;  // cannot happen by construction
                        }
                    }
                    for (int l = 1; l < L; ++l) {  // thread layout and shared-memory footprint of every level
                        PyrRect& r = t.lv[l];
                        const int nw = r.x1n - r.x0, nh = r.y1n - r.y0;
                        if (nw <= 0 || nh <= 0) { r.ng = 0; continue; }
                        const int ng = (r.x1n - (r.x0 & ~3) + 3) >> 2;
                        if (ng > 256) { fits = false; break; }
                        r.ng = (short)ng;
                        r.nrl = (short)(256 / ng);
                        r.chunk = (short)((nh + r.nrl - 1) / r.nrl);
                        r.pitch = (short)(4 * ng + 12);
                        r.magic = 65536 / ng + 1;
                        // the source row origin of level l >= 2 is 4 px left of the source tile's first group: the first group's
                        // left tap must not lie further left
                        if (l >= 2 && hx[l][r.x0 & ~3] < (t.lv[l - 1].x0 & ~3) - 4) fits = false;
                        const int bytes = r.pitch * nh + 16;
                        if (l & 1) maxB = std::max(maxB, bytes);  // odd levels live in buffer B, even levels reuse the level-0 box
                        else if (bytes > SVO_PYR_BOX_W * SVO_PYR_BOX_H) fits = false;
                    }
                    tiles.push_back(t);
                }
            if (fits) {
                c->ptiles = tiles;
                c->pyr_bufB_bytes = (int)svo_align(maxB + 16, 128);
            } else if (TW > 64) {
                TW -= 32;
                TH = std::max(32, TH - 8);
            } else {
                ok = false;
            }
        }
        if (ok && !c->ptiles.empty()) {
            // TMA descriptor: level 0 of all slots as one 2-D u8 tensor [max_images * h0 rows][pitch0 cols]
            typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                         const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult qres;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn && qres == cudaDriverEntryPointSuccess) {
                const cuuint64_t dims[2] = {(cuuint64_t)g.lvl[0].pitch, (cuuint64_t)g.lvl[0].h * (cuuint64_t)S};
                const cuuint64_t strides[1] = {(cuuint64_t)g.lvl[0].pitch};
                const cuuint32_t box[2] = {SVO_PYR_BOX_W, SVO_PYR_BOX_H};
                const cuuint32_t estr[2] = {1, 1};
                const CUresult r = ((EncodeFn)fn)(reinterpret_cast<CUtensorMap*>(c->tmap_l0), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, g.lvl[0].base, dims, strides,
                                                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r == CUDA_SUCCESS) {
                    c->n_ptiles = (int)c->ptiles.size();
                    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_ptiles, sizeof(PyrTile) * c->n_ptiles));
                    SVO_CUDA_TRY(c, cudaMemcpy(c->d_ptiles, c->ptiles.data(), sizeof(PyrTile) * c->n_ptiles, cudaMemcpyHostToDevice));
                    c->fused_pyr = true;
                }
            }
        }
    }

    // ---- cell grid, FAST tiles, octree roots, capacities
    c->tiles.clear();
    uint32_t tile_off = 0, cand_off = 0, lkp_off = 0;
    int ncap = 0;
    for (int l = 0; l < L; ++l) {
        LevelGeom& G = g.lvl[l];
        G.tile_begin = (int)c->tiles.size();
        G.cand_off = (int)cand_off;
        G.lkp_off = (int)lkp_off;
        const int minB = SVO_MINB, maxBX = G.w - minB, maxBY = G.h - minB;
        G.Wp = maxBX - minB;
        G.Hp = maxBY - minB;
        G.N = c->features_per_level[l];
        if (G.Wp >= 30 && G.Hp >= 30) {
            const float width = (float)G.Wp, height = (float)G.Hp;
            G.nCols = (int)(width / 30.f);
            G.nRows = (int)(height / 30.f);
            G.wCell = (int)std::ceil(width / G.nCols);
            G.hCell = (int)std::ceil(height / G.nRows);
            G.nIni = (int)std::round((float)G.Wp / (float)G.Hp);
            if (G.nIni < 1) return c->fail(SVO_ERR_INVALID, "portrait images (width/height < 0.5) are undefined in the reference (nIni == 0)");
            G.hX = (float)G.Wp / G.nIni;
            const int xe = G.w - SVO_EDGE, ye = G.h - SVO_EDGE;  // exclusive end of the FAST inner domain
            const int cells_per_tile = std::max(1, std::min(32, 249 / G.wCell));  // <= 249 px: one full 64-word pre-test step per row in every alignment
            uint32_t lvl_cap = 0;
            for (int i = 0; i < G.nRows; ++i) {
                const int y0 = SVO_EDGE + i * G.hCell, y1 = std::min(y0 + G.hCell, ye);
                if (y0 >= y1) continue;
                for (int j0 = 0; j0 < G.nCols; j0 += cells_per_tile) {
                    const int x0 = SVO_EDGE + j0 * G.wCell;
                    if (x0 >= xe) break;
                    int nc = 0;
                    while (nc < cells_per_tile && j0 + nc < G.nCols && SVO_EDGE + (j0 + nc) * G.wCell < xe) ++nc;
                    FastTile t;
                    t.level = (short)l;
                    t.cell_i = (short)i;
                    t.cell_j0 = (short)j0;
                    t.ncells = (short)nc;
                    t.x0 = x0;
                    t.x1 = std::min(SVO_EDGE + (j0 + nc) * G.wCell, xe);
                    t.y0 = y0;
                    t.y1 = y1;
                    t.out_cap = std::max(64, (t.x1 - t.x0) * (t.y1 - t.y0) / 8);
                    t.out_off = (int)tile_off;
                    tile_off += (uint32_t)t.out_cap;
                    lvl_cap += (uint32_t)t.out_cap;
                    c->tiles.push_back(t);
                }
            }
            G.cand_cap = (int)std::min<uint32_t>(lvl_cap, (1u << 24) - 1);
            G.lkp_cap = level_kp_cap(c, G.w, G.h, l);
        }
        G.tile_count = (int)c->tiles.size() - G.tile_begin;
        cand_off += (uint32_t)svo_align(G.cand_cap, 4);
        lkp_off += (uint32_t)G.lkp_cap;
        ncap = std::max(ncap, G.lkp_cap);
    }
    c->n_tiles = (int)c->tiles.size();
    // K2 tensor maps: level l of all slots as [S * h rows][pitch / 16 chunks][16 bytes]; the 17-chunk box gives a 272-byte wide
    // window (a 2-D box could be 256 bytes at most); the start column of a box is a chunk index, i.e. 16-byte aligned
    c->fast_tma = false;
    if (c->n_tiles) {
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn && qres == cudaDriverEntryPointSuccess) {
            bool ok = true;
            for (int l = 0; l < L && ok; ++l) {
                const LevelGeom& G = g.lvl[l];
                if (G.tile_count == 0) continue;
                const cuuint64_t dims[3] = {16, (cuuint64_t)G.pitch / 16, (cuuint64_t)G.h * (cuuint64_t)S};
                const cuuint64_t strides[2] = {16, (cuuint64_t)G.pitch};
                const cuuint32_t box[3] = {16, 17, (cuuint32_t)(G.hCell + 6)};
                const cuuint32_t estr[3] = {1, 1, 1};
                ok = ((EncodeFn)fn)(reinterpret_cast<CUtensorMap*>(c->tmap_fast[l]), CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, G.base, dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
            }
            c->fast_tma = ok;
        }
    }
    c->tile_buf_per_slot = tile_off;
    c->cand_per_slot = cand_off;
    c->lkp_per_slot = lkp_off;
    c->kp_cap = (int)lkp_off;
    c->octree_ncap = (int)svo_align(std::max(ncap, 32), 32);
    if (c->octree_ncap > 4096) return c->fail(SVO_ERR_INVALID, "per-level feature quota above 4092 is not supported");

    if (c->n_tiles) {
        SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_tiles, sizeof(FastTile) * c->n_tiles));
        SVO_CUDA_TRY(c, cudaMemcpy(c->d_tiles, c->tiles.data(), sizeof(FastTile) * c->n_tiles, cudaMemcpyHostToDevice));
    }
    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_tile_buf, sizeof(uint32_t) * std::max<size_t>(1, (size_t)c->tile_buf_per_slot * S)));
    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_tile_cnt, sizeof(int) * std::max<size_t>(1, (size_t)c->n_tiles * S)));
    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_tile_pos, sizeof(int) * std::max<size_t>(1, (size_t)c->n_tiles * S)));
    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_cand, sizeof(uint32_t) * std::max<size_t>(1, (size_t)c->cand_per_slot * S)));
    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_knode, sizeof(uint16_t) * std::max<size_t>(1, (size_t)c->cand_per_slot * S)));
    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_ncand, sizeof(int) * SVO_MAX_LEVELS * S));
    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_lkp, sizeof(uint32_t) * std::max<size_t>(1, (size_t)c->lkp_per_slot * S)));
    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_nlkp, sizeof(int) * SVO_MAX_LEVELS * S));
    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_kps, sizeof(svo_keypoint) * std::max<size_t>(1, (size_t)c->kp_cap * S)));
    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_desc, std::max<size_t>(1, (size_t)32 * c->kp_cap * S)));
    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_nkp, sizeof(int) * S));
    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_flags, sizeof(int) * S));
    SVO_CUDA_TRY(c, cudaMemsetAsync(c->d_ncand, 0, sizeof(int) * SVO_MAX_LEVELS * S, c->stream));
    SVO_CUDA_TRY(c, cudaMemsetAsync(c->d_nlkp, 0, sizeof(int) * SVO_MAX_LEVELS * S, c->stream));
    SVO_CUDA_TRY(c, cudaMemsetAsync(c->d_nkp, 0, sizeof(int) * S, c->stream));
    SVO_CUDA_TRY(c, cudaMemsetAsync(c->d_flags, 0, sizeof(int) * S, c->stream));
    c->cur_w = w;
    c->cur_h = h;
    return SVO_OK;
}

int svo_check_flags(svo_ctx* c, const SlotMap& map, int n_img) {
    // reads the per-slot error flags (after the stream has been synchronised by the caller's D2H copies)
    std::vector<int> flags(c->max_images);
    SVO_CUDA_TRY(c, cudaMemcpyAsync(flags.data(), c->d_flags, sizeof(int) * c->max_images, cudaMemcpyDeviceToHost, c->stream));
    SVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (int b = 0; b < n_img; ++b) {
        const int f = flags[map.slot(b)];
        if (f) {
            cudaMemsetAsync(c->d_flags, 0, sizeof(int) * c->max_images, c->stream);
            char msg[160];
            snprintf(msg, sizeof msg, "device capacity exceeded on image %d (flags 0x%x: 1 tile, 2 candidates, 4 octree nodes, 8 keypoints, 16 fast)", b, f);
            return c->fail(SVO_ERR_CAPACITY, msg);
        }
    }
    return SVO_OK;
}

// runs K1..K5 for n_img images already resident in level 0 of their slots
int svo_orb_run(svo_ctx* c, const SlotMap& map, int n_img) {
    int rc;
    if ((rc = launch_pyramid(c, map, n_img)) != SVO_OK) return rc;
    if ((rc = launch_fast_cells(c, map, n_img)) != SVO_OK) return rc;
    if ((rc = launch_octree(c, map, n_img)) != SVO_OK) return rc;
    if ((rc = launch_describe(c, map, n_img)) != SVO_OK) return rc;
    return SVO_OK;
}

extern "C" int svo_orb_extract_batch(svo_ctx* c, const uint8_t* const* imgs, int n_imgs, int w, int h, size_t stride,
                                     svo_keypoint* kps, uint8_t* desc, int cap_per_img, int* n_out) {
    if (!c) return SVO_ERR_INVALID;
    if (!c->user_configured) return c->fail(SVO_ERR_NOT_CONFIGURED, "svo_orb_configure() first");
    if (n_imgs < 0 || n_imgs > c->max_images || !n_out) return c->fail(SVO_ERR_INVALID, "svo_orb_extract_batch: bad n_imgs / n_out");
    for (int i = 0; i < n_imgs; ++i) n_out[i] = 0;
    if (n_imgs == 0 || w == 0 || h == 0 || !imgs) return SVO_OK;  // empty image: silent return (ORBextractor.cpp:993-994)
    if (w < 0 || h < 0 || stride < (size_t)w || !kps || !desc) return c->fail(SVO_ERR_INVALID, "svo_orb_extract_batch: bad geometry / null outputs");
    cudaSetDevice(c->device);
    int rc = svo_standalone_guard(c);
    if (rc != SVO_OK) return rc;
    if ((rc = svo_layout(c, w, h)) != SVO_OK) return rc;
    if (cap_per_img < c->kp_cap) return c->fail(SVO_ERR_CAPACITY, "cap_per_img < svo_orb_capacity()");
    const LevelGeom& L0 = c->geom.lvl[0];
    for (int i = 0; i < n_imgs; ++i) {
        if (!imgs[i]) return c->fail(SVO_ERR_INVALID, "null image pointer");
        SVO_CUDA_TRY(c, cudaMemcpy2DAsync(L0.base + (size_t)i * L0.plane, L0.pitch, imgs[i], stride, w, h, cudaMemcpyHostToDevice, c->stream));
    }
    SlotMap map{1, 0, 0, c->max_images};
    if ((rc = svo_orb_run(c, map, n_imgs)) != SVO_OK) return rc;
    std::vector<int> n(n_imgs);
    SVO_CUDA_TRY(c, cudaMemcpyAsync(n.data(), c->d_nkp, sizeof(int) * n_imgs, cudaMemcpyDeviceToHost, c->stream));
    if ((rc = svo_check_flags(c, map, n_imgs)) != SVO_OK) return rc;  // synchronises
    for (int i = 0; i < n_imgs; ++i) {
        if (n[i] > cap_per_img) return c->fail(SVO_ERR_CAPACITY, "keypoint count above cap_per_img");
        if (n[i] > 0) {
            SVO_CUDA_TRY(c, cudaMemcpyAsync(kps + (size_t)i * cap_per_img, c->d_kps + (size_t)i * c->kp_cap, sizeof(svo_keypoint) * n[i],
                                            cudaMemcpyDeviceToHost, c->stream));
            SVO_CUDA_TRY(c, cudaMemcpyAsync(desc + (size_t)i * cap_per_img * 32, c->d_desc + (size_t)i * c->kp_cap * 32, (size_t)32 * n[i],
                                            cudaMemcpyDeviceToHost, c->stream));
        }
        n_out[i] = n[i];
    }
    SVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SVO_OK;
}

extern "C" int svo_orb_extract(svo_ctx* c, const uint8_t* img, int w, int h, size_t stride, svo_keypoint* kps, uint8_t* desc,
                               int cap, int* n_out) {
    if (!c) return SVO_ERR_INVALID;
    if (!n_out) return c->fail(SVO_ERR_INVALID, "n_out is null");
    *n_out = 0;
    if (!img || w == 0 || h == 0) return SVO_OK;
    const uint8_t* one[1] = {img};
    return svo_orb_extract_batch(c, one, 1, w, h, stride, kps, desc, cap, n_out);
}

extern "C" int svo_orb_read_level(svo_ctx* c, int img_slot, int level, uint8_t* dst, size_t dst_stride, int* w, int* h) {
    if (!c) return SVO_ERR_INVALID;
    if (!c->d_pyr) return c->fail(SVO_ERR_INVALID, "no image has been processed yet");
    if (img_slot < 0 || img_slot >= c->max_images || level < 0 || level >= c->nlevels) return c->fail(SVO_ERR_INVALID, "bad slot / level");
    const LevelGeom& G = c->geom.lvl[level];
    if (w) *w = G.w;
    if (h) *h = G.h;
    if (!dst) return SVO_OK;
    if (dst_stride < (size_t)G.w) return c->fail(SVO_ERR_INVALID, "dst_stride < level width");
    cudaSetDevice(c->device);
    SVO_CUDA_TRY(c, cudaMemcpy2DAsync(dst, dst_stride, G.base + (size_t)img_slot * G.plane, G.pitch, G.w, G.h, cudaMemcpyDeviceToHost, c->stream));
    SVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SVO_OK;
}

extern "C" int svo_orb_read_candidates(svo_ctx* c, int img_slot, int level, svo_keypoint* out, int cap, int* n_out) {
    if (!c) return SVO_ERR_INVALID;
    if (!c->d_cand || !n_out) return c->fail(SVO_ERR_INVALID, "no image has been processed yet / n_out null");
    if (img_slot < 0 || img_slot >= c->max_images || level < 0 || level >= c->nlevels) return c->fail(SVO_ERR_INVALID, "bad slot / level");
    cudaSetDevice(c->device);
    int n = 0;
    SVO_CUDA_TRY(c, cudaMemcpyAsync(&n, c->d_ncand + img_slot * SVO_MAX_LEVELS + level, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    *n_out = n;
    if (!out || n == 0) return SVO_OK;
    if (n > cap) return c->fail(SVO_ERR_CAPACITY, "candidate count above cap");
    std::vector<uint32_t> packed(n);
    SVO_CUDA_TRY(c, cudaMemcpyAsync(packed.data(), c->d_cand + (size_t)img_slot * c->cand_per_slot + c->geom.lvl[level].cand_off,
                                    sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    SVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; ++i)
        out[i] = svo_keypoint{(float)cand_x(packed[i]), (float)cand_y(packed[i]), 7.f, -1.f, (float)cand_r(packed[i]), 0, -1};
    return SVO_OK;
}
