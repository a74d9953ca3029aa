// svo_frontend.cu -- host orchestration of the Tracking front-end entry points of the C ABI:
//   svo_fast_detect   <- Tracking::Detect_OpenCVFASTFeatures            (tracking.cpp:94-113)
//   svo_lk_track      <- one cv::calcOpticalFlowPyrLK call               (tracking.cpp:593-618)
//   svo_frontend_*    <- Detect_MyORBFeatures + the two BF matches (tracking.cpp:502-544) and/or FAST + the 4-call LK
//                        chain (tracking.cpp:258-271, 583-618) for a batch of consecutive stereo frames, with the
//                        previous frame (Tracking::last_frame_) carried in a ring of frame slots on the device.
// No pixel arithmetic happens here; see k_*.cu.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "svo_common.cuh"

int launch_lk_standalone(svo_ctx* c, int n, int max_level, int max_iter, double eps, double min_eig, int want_err);
int svo_orb_configure_impl(svo_ctx* c, int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, bool by_user);

static int ensure_config(svo_ctx* c) {
    if (c->configured) return SVO_OK;
    return svo_orb_configure_impl(c, 0, 1.2f, 1, 20, 7, false);  // detector-only layout: level 0 only
}

// The standalone entry points (svo_orb_extract*, svo_fast_detect, svo_lk_track) use image slots 0 / 1, which are also ring slots of
// the throughput mode: on a frontend-configured context they refuse to run while results are still queued, and they drop the
// carried frame (the next svo_frontend_run starts like after svo_frontend_reset).
int svo_standalone_guard(svo_ctx* c) {
    if (!c->fe_configured) return SVO_OK;
    if (c->fe_nq > 0) return c->fail(SVO_ERR_INVALID, "standalone call on a frontend context with undelivered runs: svo_frontend_download() them first");
    if (c->d2h_stream) cudaStreamSynchronize(c->d2h_stream);
    c->fe_has_last = false;
    c->fe_pending = 0;
    return SVO_OK;
}

// LK pyramid levels 1..3 and the full-frame FAST / LK point buffers for the current image size
static int layout_aux(svo_ctx* c, int pt_cap) {
    if (c->aux_ready && c->pt_cap >= pt_cap) return SVO_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    void** ptrs[] = {(void**)&c->d_lkderiv,  (void**)&c->d_lkpyr,    (void**)&c->d_ftiles,  (void**)&c->d_ftile_buf, (void**)&c->d_ftile_cnt, (void**)&c->d_fast_kps,
                     (void**)&c->d_nfast,    (void**)&c->d_lk_pts,  (void**)&c->d_lk_status, (void**)&c->d_lk_err};
    for (void** p : ptrs) {
        if (*p) svo_dev_free(*p);
        *p = nullptr;
    }
    c->aux_ready = false;
    // the throughput-mode buffers (d_lkf_pts, d_lkf_xyz, ...) were sized with the old pt_cap: svo_frontend_configure() must run again
    c->fe_configured = false;
    const int S = c->max_images, w = c->cur_w, h = c->cur_h;
    if (w > SVO_MAX_DIM || h > SVO_MAX_DIM) return c->fail(SVO_ERR_INVALID, "full-frame FAST / LK support images up to 4096 x 4096");
    const LevelGeom& L0 = c->geom.lvl[0];
    c->lk[0] = LkLevelGeom{L0.base, L0.w, L0.h, L0.pitch, L0.plane};
    c->lk_levels = 1;
    // padded planes (REFLECT_101 border of SVO_LK_PADX columns / SVO_LK_PADY rows): a copy of level 0 and levels 1..3
    size_t total = 0;
    std::vector<size_t> off(SVO_LK_LEVELS, 0);
    auto padded = [&](int lw, int lh, size_t& o) {
        LkLevelGeom g{};
        g.w = lw;
        g.h = lh;
        g.pitch = (int)svo_align(lw + 2 * SVO_LK_PADX, 64);
        g.plane = (uint32_t)g.pitch * (uint32_t)(lh + 2 * SVO_LK_PADY);
        o = total;
        total += svo_align((size_t)g.plane * S, 256);
        return g;
    };
    c->lkp0 = padded(L0.w, L0.h, off[0]);
    for (int l = 1; l < SVO_LK_LEVELS; ++l) {
        const int lw = (c->lk[l - 1].w + 1) / 2, lh = (c->lk[l - 1].h + 1) / 2;
        if (lw <= SVO_LK_WIN || lh <= SVO_LK_WIN) break;  // buildOpticalFlowPyramid stops before a level <= winSize
        c->lk[l] = padded(lw, lh, off[l]);
        c->lk_levels = l + 1;
    }
    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_lkpyr, total + 4096));
    SVO_CUDA_TRY(c, cudaMemsetAsync(c->d_lkpyr, 0, total + 4096, c->stream));
    c->lkp0.base = c->d_lkpyr + off[0] + (size_t)SVO_LK_PADY * c->lkp0.pitch + SVO_LK_PADX;
    for (int l = 1; l < c->lk_levels; ++l) c->lk[l].base = c->d_lkpyr + off[l] + (size_t)SVO_LK_PADY * c->lk[l].pitch + SVO_LK_PADX;

    // Scharr planes: 4 bytes per pixel, the same padding as the image planes, border zero (set once here: the kernel writes the interior)
    c->lk_deriv = getenv("SVO_LK_DERIV") != nullptr;  // measured alternative, off by default (see k_lk.cu / DESIGN.md)
    if (const char* gx = getenv("SVO_LK_GRID")) c->lk_grid_x = std::max(32, atoi(gx));  // A/B measurements of the chain launch's grid
    if (c->lk_deriv) {
        size_t dtotal = 0;
        std::vector<size_t> doff(SVO_LK_LEVELS, 0);
        for (int l = 0; l < c->lk_levels; ++l) {
            const LkLevelGeom& G = l == 0 ? c->lkp0 : c->lk[l];
            LkLevelGeom d{};
            d.w = G.w;
            d.h = G.h;
            d.pitch = 4 * (int)svo_align(G.w + 2 * SVO_LK_PADX, 16);
            d.plane = (uint32_t)d.pitch * (uint32_t)(G.h + 2 * SVO_LK_PADY);
            doff[l] = dtotal;
            dtotal += svo_align((size_t)d.plane * S, 256);
            c->lkd[l] = d;
        }
        SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_lkderiv, dtotal + 4096));
        SVO_CUDA_TRY(c, cudaMemsetAsync(c->d_lkderiv, 0, dtotal + 4096, c->stream));
        for (int l = 0; l < c->lk_levels; ++l) c->lkd[l].base = c->d_lkderiv + doff[l] + (size_t)SVO_LK_PADY * c->lkd[l].pitch + (size_t)SVO_LK_PADX * 4;
    }

    // full-frame FAST strips over the inner domain [3, w-3) x [3, h-3)
    c->ftiles.clear();
    uint32_t boff = 0;
    if (w >= 7 && h >= 7) {
        const int R = (int)svo_align(w, 16) <= 2048 ? 16 : 8;
        for (int y0 = 3; y0 < h - 3; y0 += R) {
            FastTile t{};
            t.x0 = 3;
            t.x1 = w - 3;
            t.y0 = y0;
            t.y1 = std::min(y0 + R, h - 3);
            t.out_cap = std::max(64, (t.x1 - t.x0) * (t.y1 - t.y0) / 8);
            t.out_off = (int)boff;
            boff += (uint32_t)t.out_cap;
            c->ftiles.push_back(t);
        }
    }
    c->n_ftiles = (int)c->ftiles.size();
    c->ftile_buf_per_slot = boff;
    c->pt_cap = std::max(pt_cap, 256);
    if (c->n_ftiles) {
        SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_ftiles, sizeof(FastTile) * c->n_ftiles));
        SVO_CUDA_TRY(c, cudaMemcpy(c->d_ftiles, c->ftiles.data(), sizeof(FastTile) * c->n_ftiles, cudaMemcpyHostToDevice));
    }
    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_ftile_buf, sizeof(uint32_t) * std::max<size_t>(1, (size_t)boff * S)));
    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_ftile_cnt, sizeof(int) * std::max<size_t>(1, (size_t)c->n_ftiles * S)));
    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_fast_kps, sizeof(svo_keypoint) * (size_t)c->pt_cap * S));
    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_nfast, sizeof(int) * S));
    SVO_CUDA_TRY(c, cudaMemsetAsync(c->d_nfast, 0, sizeof(int) * S, c->stream));
    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_lk_pts, sizeof(float) * 2 * 5 * (size_t)c->pt_cap * S));
    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_lk_status, (size_t)4 * c->pt_cap * S));
    SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_lk_err, sizeof(float) * 4 * (size_t)c->pt_cap * S));
    c->aux_ready = true;
    return SVO_OK;
}

extern "C" int svo_fast_detect(svo_ctx* c, const uint8_t* img, int w, int h, size_t stride, int threshold, int nms, svo_keypoint* kps,
                               int cap, int* n_out) {
    if (!c) return SVO_ERR_INVALID;
    if (!n_out) return c->fail(SVO_ERR_INVALID, "n_out is null");
    *n_out = 0;
    if (!img || w == 0 || h == 0) return SVO_OK;
    if (!nms) return c->fail(SVO_ERR_INVALID, "svo_fast_detect: only nonmaxSuppression=true is on the reference path (tracking.cpp:100)");
    if (w < 0 || h < 0 || stride < (size_t)w || cap < 0 || (cap > 0 && !kps)) return c->fail(SVO_ERR_INVALID, "svo_fast_detect: bad arguments");
    cudaSetDevice(c->device);
    int rc;
    if ((rc = svo_standalone_guard(c)) != SVO_OK) return rc;
    if ((rc = ensure_config(c)) != SVO_OK) return rc;
    if ((rc = svo_layout(c, w, h)) != SVO_OK) return rc;
    if ((rc = layout_aux(c, std::max(cap, w * h / 64))) != SVO_OK) return rc;
    const LevelGeom& L0 = c->geom.lvl[0];
    SVO_CUDA_TRY(c, cudaMemcpy2DAsync(L0.base, L0.pitch, img, stride, w, h, cudaMemcpyHostToDevice, c->stream));
    SlotMap map{1, 0, 0, c->max_images};
    if ((rc = launch_fast_full(c, map, 1, threshold)) != SVO_OK) return rc;
    int n = 0;
    SVO_CUDA_TRY(c, cudaMemcpyAsync(&n, c->d_nfast, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if ((rc = svo_check_flags(c, map, 1)) != SVO_OK) return rc;
    *n_out = n;
    if (n > cap) return c->fail(SVO_ERR_CAPACITY, "svo_fast_detect: more corners than cap");
    if (n > 0) SVO_CUDA_TRY(c, cudaMemcpyAsync(kps, c->d_fast_kps, sizeof(svo_keypoint) * n, cudaMemcpyDeviceToHost, c->stream));
    SVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SVO_OK;
}

extern "C" int svo_lk_track(svo_ctx* c, const uint8_t* prev, const uint8_t* next, int w, int h, size_t stride, const float* prev_xy, int n,
                            float* next_xy, uint8_t* status, float* err, int win, int max_level, int max_iter, double eps, double min_eig) {
    if (!c) return SVO_ERR_INVALID;
    if (win != SVO_LK_WIN) return c->fail(SVO_ERR_INVALID, "svo_lk_track: winSize must be 21x21 (tracking.cpp:594)");
    if (max_level < 0 || max_level > SVO_LK_LEVELS - 1) return c->fail(SVO_ERR_INVALID, "svo_lk_track: maxLevel must be 0..3");
    if (!prev || !next || w <= 0 || h <= 0 || stride < (size_t)w || n < 0 || (n > 0 && (!prev_xy || !next_xy || !status)))
        return c->fail(SVO_ERR_INVALID, "svo_lk_track: bad arguments");
    if (c->max_images < 2) return c->fail(SVO_ERR_INVALID, "svo_lk_track needs a context with max_images >= 2");
    if (n == 0) return SVO_OK;
    cudaSetDevice(c->device);
    int rc;
    if ((rc = svo_standalone_guard(c)) != SVO_OK) return rc;
    if ((rc = ensure_config(c)) != SVO_OK) return rc;
    if ((rc = svo_layout(c, w, h)) != SVO_OK) return rc;
    if ((rc = layout_aux(c, std::max(n, w * h / 64))) != SVO_OK) return rc;
    const LevelGeom& L0 = c->geom.lvl[0];
    SVO_CUDA_TRY(c, cudaMemcpy2DAsync(L0.base, L0.pitch, prev, stride, w, h, cudaMemcpyHostToDevice, c->stream));
    SVO_CUDA_TRY(c, cudaMemcpy2DAsync(L0.base + L0.plane, L0.pitch, next, stride, w, h, cudaMemcpyHostToDevice, c->stream));
    SVO_CUDA_TRY(c, cudaMemcpyAsync(c->d_lk_pts, prev_xy, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, c->stream));
    SlotMap map{1, 0, 0, c->max_images};
    if ((rc = launch_lk_pyramid(c, map, 2)) != SVO_OK) return rc;
    if ((rc = launch_lk_standalone(c, n, max_level, max_iter, eps, min_eig, err != nullptr)) != SVO_OK) return rc;
    SVO_CUDA_TRY(c, cudaMemcpyAsync(next_xy, c->d_lk_pts + (size_t)1 * c->pt_cap * 2, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, c->stream));
    SVO_CUDA_TRY(c, cudaMemcpyAsync(status, c->d_lk_status, n, cudaMemcpyDeviceToHost, c->stream));
    if (err) SVO_CUDA_TRY(c, cudaMemcpyAsync(err, c->d_lk_err, sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
    SVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SVO_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// throughput mode
// ---------------------------------------------------------------------------------------------------------------
extern "C" int svo_frontend_configure(svo_ctx* c, int w, int h, int mode, int max_frames) {
    if (!c) return SVO_ERR_INVALID;
    if (!(mode & (SVO_MODE_ORB | SVO_MODE_LK)) || (mode & ~(SVO_MODE_ORB | SVO_MODE_LK)) || w <= 0 || h <= 0 || max_frames < 1)
        return c->fail(SVO_ERR_INVALID, "svo_frontend_configure: bad mode / size / max_frames");
    if (2 * (max_frames + 1) > c->max_images) return c->fail(SVO_ERR_INVALID, "svo_frontend_configure: need max_images >= 2 * (max_frames + 1)");
    if (c->fe_configured) cudaStreamSynchronize(c->stream);
    svo_pnp_free_frontend(c);  // sized by the ring and the capacities below: svo_frontend_set_pnp() must follow a (re)configure
    if ((mode & SVO_MODE_ORB) && !c->user_configured) return c->fail(SVO_ERR_NOT_CONFIGURED, "svo_orb_configure() first");
    cudaSetDevice(c->device);
    int rc;
    if ((rc = ensure_config(c)) != SVO_OK) return rc;
    if ((rc = svo_layout(c, w, h)) != SVO_OK) return rc;
    if (mode & SVO_MODE_LK)
        if ((rc = layout_aux(c, std::max(1024, w * h / 64))) != SVO_OK) return rc;
    // ring of frame slots: previous frame + one batch, or + two batches when the context has room -- then the results of
    // batch k can be downloaded (own stream) while batch k+1 is already running
    c->fe_depth = (2 * (2 * max_frames + 1) <= c->max_images) ? 2 : 1;
    c->fe_ring = c->fe_depth * max_frames + 1;
    if (mode & SVO_MODE_ORB) {
        void** ptrs[] = {(void**)&c->d_match, (void**)&c->d_xdesc, (void**)&c->d_xpop, (void**)&c->d_xslots};
        for (void** p : ptrs) {
            if (*p) svo_dev_free(*p);
            *p = nullptr;
        }
        SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_match, sizeof(int) * 4 * (size_t)c->kp_cap * c->fe_ring));
        c->match_mma = getenv("SVO_MATCH_POPC") == nullptr;
        c->match_tc = c->match_mma && getenv("SVO_MATCH_IMMA") == nullptr;  // default: the tcgen05 form, which needs no expanded copies
        if (c->match_mma && !c->match_tc) {
            SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_xdesc, (size_t)256 * c->kp_cap * c->max_images));
            SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_xpop, sizeof(int) * (size_t)c->kp_cap * c->max_images));
            c->xslots_cap = (size_t)2 * max_frames;
            SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_xslots, sizeof(int) * c->xslots_cap));
        }
    }
    if (!c->copy_stream) {
        SVO_CUDA_TRY(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            SVO_CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_copy[i], cudaEventDisableTiming));
            SVO_CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_repitch[i], cudaEventDisableTiming));
        }
    }
    SVO_CUDA_TRY(c, cudaStreamSynchronize(c->copy_stream));
    for (int i = 0; i < 2; ++i) {
        if (c->d_stage[i]) svo_dev_free(c->d_stage[i]);
        c->d_stage[i] = nullptr;
        c->ev_repitch_valid[i] = false;
    }
    c->stage_bytes = 2 * svo_align((size_t)max_frames * h * (2 * (size_t)w), 256);  // 2 sides, stride up to 2 x width
    for (int i = 0; i < 2; ++i) SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_stage[i], c->stage_bytes));
    c->stage_idx = 0;
    {
        void** ptrs[] = {(void**)&c->d_orb_pts, (void**)&c->d_norb, (void**)&c->d_lkf_pts, (void**)&c->d_nlkf, (void**)&c->d_orb_xyz, (void**)&c->d_lkf_xyz};
        for (void** p : ptrs) {
            if (*p) svo_dev_free(*p);
            *p = nullptr;
        }
    }
    if (mode & SVO_MODE_ORB) {
        SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_orb_pts, sizeof(float) * 2 * 3 * (size_t)c->kp_cap * c->fe_ring));
        SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_norb, sizeof(int) * c->fe_ring));
        SVO_CUDA_TRY(c, cudaMemsetAsync(c->d_norb, 0, sizeof(int) * c->fe_ring, c->stream));
        SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_orb_xyz, sizeof(float) * 3 * (size_t)c->kp_cap * c->fe_ring));
    }
    if (mode & SVO_MODE_LK) {
        SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_lkf_pts, sizeof(float) * 2 * 5 * (size_t)c->pt_cap * c->fe_ring));
        SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_nlkf, sizeof(int) * c->fe_ring));
        SVO_CUDA_TRY(c, cudaMemsetAsync(c->d_nlkf, 0, sizeof(int) * c->fe_ring, c->stream));
        SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_lkf_xyz, sizeof(float) * 3 * (size_t)c->pt_cap * c->fe_ring));
    }
    if (!c->d_flag_accum) {
        SVO_CUDA_TRY(c, svo_dev_malloc(&c->d_flag_accum, sizeof(int)));
        SVO_CUDA_TRY(c, cudaMemsetAsync(c->d_flag_accum, 0, sizeof(int), c->stream));
    }
    c->fe_mode = mode;
    c->fe_max_frames = max_frames;
    if (!c->d2h_stream) {
        SVO_CUDA_TRY(c, cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) SVO_CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
    }
    c->fe_head = c->fe_ring - 1;
    c->fe_has_last = false;
    c->fe_pending = 0;
    c->fe_nq = 0;
    c->fe_configured = true;
    return SVO_OK;
}

extern "C" int svo_frontend_set_match_error(svo_ctx* c, double fme) {
    if (!c) return SVO_ERR_INVALID;
    c->fe_match_error = fme;
    return SVO_OK;
}

extern "C" int svo_frontend_set_lk_err(svo_ctx* c, int enable) {
    if (!c) return SVO_ERR_INVALID;
    c->fe_lk_err = enable != 0;
    return SVO_OK;
}

extern "C" int svo_frontend_set_overlap(svo_ctx* c, int enable) {
    if (!c) return SVO_ERR_INVALID;
    c->fe_overlap = enable != 0;
    return SVO_OK;
}

extern "C" int svo_frontend_set_projection(svo_ctx* c, const double* p1, const double* p2) {
    if (!c || !p1 || !p2) return SVO_ERR_INVALID;
    for (int i = 0; i < 12; ++i) {
        c->proj1[i] = p1[i];
        c->proj2[i] = p2[i];
    }
    c->have_proj = true;
    return SVO_OK;
}

extern "C" int svo_frontend_reset(svo_ctx* c) {
    if (!c) return SVO_ERR_INVALID;
    if (!c->fe_configured) return c->fail(SVO_ERR_INVALID, "svo_frontend_configure() first");
    if (c->d2h_stream) cudaStreamSynchronize(c->d2h_stream);
    cudaStreamSynchronize(c->stream);
    c->fe_has_last = false;
    c->fe_pending = 0;
    c->fe_nq = 0;
    return SVO_OK;
}

static int frontend_upload_impl(svo_ctx* c, const uint8_t* left, const uint8_t* right, int n_frames, size_t stride, cudaMemcpyKind kind);

extern "C" int svo_frontend_upload(svo_ctx* c, const uint8_t* left, const uint8_t* right, int n_frames, size_t stride) {
    return frontend_upload_impl(c, left, right, n_frames, stride, cudaMemcpyHostToDevice);
}

extern "C" int svo_frontend_wait_upload(svo_ctx* c) {
    if (!c) return SVO_ERR_INVALID;
    if (!c->fe_configured || !c->copy_stream) return c->fail(SVO_ERR_INVALID, "svo_frontend_configure() first");
    cudaSetDevice(c->device);
    SVO_CUDA_TRY(c, cudaStreamSynchronize(c->copy_stream));
    return SVO_OK;
}

extern "C" int svo_frontend_upload_dev(svo_ctx* c, const uint8_t* left, const uint8_t* right, int n_frames, size_t stride) {
    return frontend_upload_impl(c, left, right, n_frames, stride, cudaMemcpyDeviceToDevice);
}

static int frontend_upload_impl(svo_ctx* c, const uint8_t* left, const uint8_t* right, int n_frames, size_t stride, cudaMemcpyKind kind) {
    // Host images: ONE linear H2D copy per side into a staging buffer on the context's copy stream (full PCIe rate; a 2-D
    // copy of 1241-byte rows is several times slower), double-buffered so that the upload of batch k+1 overlaps the kernels
    // of batch k.  Device images: nothing is copied here.  In both cases svo_frontend_run() moves the rows into the pitched
    // level-0 planes of the frame ring with a kernel (k_repitch).
    if (!c) return SVO_ERR_INVALID;
    if (!c->fe_configured) return c->fail(SVO_ERR_INVALID, "svo_frontend_configure() first");
    if (!left || !right || n_frames < 1 || n_frames > c->fe_max_frames || stride < (size_t)c->cur_w)
        return c->fail(SVO_ERR_INVALID, "svo_frontend_upload: bad arguments");
    cudaSetDevice(c->device);
    const int h = c->cur_h;
    if (kind == cudaMemcpyHostToDevice) {
        const int sb = c->stage_idx ^ 1;  // the buffer the previous run did not read
        const size_t bytes = (size_t)n_frames * h * stride;
        if (bytes > c->stage_bytes / 2) return c->fail(SVO_ERR_INVALID, "svo_frontend_upload: stride larger than 2 x width is not supported");
        if (c->ev_repitch_valid[sb]) SVO_CUDA_TRY(c, cudaStreamWaitEvent(c->copy_stream, c->ev_repitch[sb], 0));
        uint8_t* dl = c->d_stage[sb];
        uint8_t* dr = dl + c->stage_bytes / 2;
        SVO_CUDA_TRY(c, cudaMemcpyAsync(dl, left, bytes, cudaMemcpyHostToDevice, c->copy_stream));
        SVO_CUDA_TRY(c, cudaMemcpyAsync(dr, right, bytes, cudaMemcpyHostToDevice, c->copy_stream));
        SVO_CUDA_TRY(c, cudaEventRecord(c->ev_copy[sb], c->copy_stream));
        c->stage_idx = sb;
        c->pend_left = dl;
        c->pend_right = dr;
        c->pend_staged = true;
    } else {
        c->pend_left = left;
        c->pend_right = right;
        c->pend_staged = false;
    }
    c->pend_stride = stride;
    c->fe_pending = n_frames;
    return SVO_OK;
}

extern "C" int svo_frontend_run(svo_ctx* c) {
    if (!c) return SVO_ERR_INVALID;
    if (!c->fe_configured || c->fe_pending < 1) return c->fail(SVO_ERR_INVALID, "svo_frontend_run: nothing uploaded");
    if (c->fe_nq >= c->fe_depth) return c->fail(SVO_ERR_INVALID, "svo_frontend_run: download the previous results first (ring depth reached)");
    cudaSetDevice(c->device);
    const int ring = c->fe_ring, first = (c->fe_head + 1) % ring, n = c->fe_pending, skip_first = c->fe_has_last ? 0 : 1;
    int rc;
    const SlotMap both{2, ring, first, ring}, leftonly{1, 0, first, ring};
    {  // the capacity flags are per run: clear those of the slots this run writes (left and right side, <= 2 runs of slots each)
        const int run1 = std::min(n, ring - first), run2 = n - run1;
        for (int side = 0; side < 2; ++side) {
            SVO_CUDA_TRY(c, cudaMemsetAsync(c->d_flags + side * ring + first, 0, sizeof(int) * run1, c->stream));
            if (run2 > 0) SVO_CUDA_TRY(c, cudaMemsetAsync(c->d_flags + side * ring, 0, sizeof(int) * run2, c->stream));
        }
    }
    if (c->pend_staged) SVO_CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_copy[c->stage_idx], 0));
    if ((rc = launch_repitch(c, c->pend_left, c->pend_right, c->pend_stride, first, ring, n)) != SVO_OK) return rc;
    if (c->pend_staged) {
        SVO_CUDA_TRY(c, cudaEventRecord(c->ev_repitch[c->stage_idx], c->stream));
        c->ev_repitch_valid[c->stage_idx] = true;
    }
    const bool overlap = c->fe_overlap && (c->fe_mode & SVO_MODE_ORB) && (c->fe_mode & SVO_MODE_LK);
    cudaStream_t main_stream = c->stream;
    if (c->fe_mode & SVO_MODE_ORB) {
        if ((rc = launch_pyramid(c, both, 2 * n)) != SVO_OK) return rc;
        if ((rc = launch_fast_cells(c, both, 2 * n)) != SVO_OK) return rc;
        if (overlap) {  // the ORB tail goes to a side stream; the FAST + LK chain below stays on the caller's stream
            if (!c->side_stream) {
                // default (lowest) priority, like the caller's stream: measured 1.2 % faster than one stream; the highest
                // priority only moves the whole ORB tail in front of the LK chain (0.9 %)
                SVO_CUDA_TRY(c, cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking));
                SVO_CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
                SVO_CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
            }
            SVO_CUDA_TRY(c, cudaEventRecord(c->ev_fork, main_stream));
            SVO_CUDA_TRY(c, cudaStreamWaitEvent(c->side_stream, c->ev_fork, 0));
            c->stream = c->side_stream;  // launchers and stage timers follow c->stream
        }
        rc = launch_octree(c, both, 2 * n);
        if (rc == SVO_OK) rc = launch_describe(c, both, 2 * n);
        if (rc == SVO_OK) rc = launch_match_frontend(c, first, ring, n, skip_first);
        if (rc == SVO_OK) rc = launch_orb_filter(c, first, ring, n, skip_first);
        if (overlap) {
            if (rc == SVO_OK && cudaEventRecord(c->ev_join, c->side_stream) != cudaSuccess) rc = c->fail(SVO_ERR_CUDA, "cudaEventRecord(ev_join)");
            c->stream = main_stream;
        }
        if (rc != SVO_OK) return rc;
    }
    if (c->fe_mode & SVO_MODE_LK) {
        if ((rc = launch_fast_full(c, leftonly, n, 20)) != SVO_OK) return rc;  // int fast_threshold = 20 (tracking.cpp:99)
        if ((rc = launch_lk_pyramid(c, both, 2 * n)) != SVO_OK) return rc;
        // the 4-call chain (tracking.cpp:593-618): L(t-1)->R(t-1), R(t-1)->R(t), R(t)->L(t), L(t)->L(t-1)
        const LkCall calls[4] = {{0, 1, 1, 1, 0, 1, 0}, {1, 1, 1, 0, 1, 2, 1}, {1, 0, 0, 0, 2, 3, 2}, {0, 0, 0, 1, 3, 4, 3}};
        if ((rc = launch_lk_chain(c, first, ring, n, skip_first, calls, 4, 30, 0.01, 0.001, c->fe_lk_err ? 1 : 0)) != SVO_OK) return rc;
        if ((rc = launch_lk_filter(c, first, ring, n, skip_first)) != SVO_OK) return rc;
    }
    if (overlap) SVO_CUDA_TRY(c, cudaStreamWaitEvent(main_stream, c->ev_join, 0));  // join before anything reads the ORB results
    if (c->have_proj && (rc = launch_triangulate_frontend(c, first, ring, n)) != SVO_OK) return rc;
    if (c->have_proj && c->fe_pnp && (rc = launch_pnp_frontend(c, first, ring, n)) != SVO_OK) return rc;
    {  // completed-run queue (depth <= 2): what svo_frontend_download delivers, oldest first
        svo_ctx::DoneRun& d = c->fe_q[(c->fe_qhead + c->fe_nq) % 2];
        d.first = first;
        d.n = n;
        d.skip = skip_first;
        d.ev = (c->fe_qhead + c->fe_nq) % 2;
        SVO_CUDA_TRY(c, cudaEventRecord(c->ev_done[d.ev], c->stream));
        ++c->fe_nq;
    }
    c->fe_head = (first + n - 1) % ring;
    c->fe_has_last = true;
    c->fe_pending = 0;
    return SVO_OK;
}

namespace {
// copy `rows_per_frame` rows of `row_bytes` bytes (device row pitch dev_pitch, host row pitch host_pitch) for the frames of
// the last run; frames occupy at most two contiguous runs of ring slots.
struct RingCopy {
    svo_ctx* c;
    int first, n, ring;
    int copy(void* host, size_t host_frame_bytes, const void* dev, size_t dev_slot_bytes, size_t width_bytes) {
        // one "row" per frame
        const int run1 = std::min(n, ring - first), run2 = n - run1;
        if (width_bytes == 0) return SVO_OK;
        SVO_CUDA_TRY(c, cudaMemcpy2DAsync(host, host_frame_bytes, (const uint8_t*)dev + (size_t)first * dev_slot_bytes, dev_slot_bytes,
                                          width_bytes, run1, cudaMemcpyDeviceToHost, c->stream));
        if (run2 > 0)
            SVO_CUDA_TRY(c, cudaMemcpy2DAsync((uint8_t*)host + (size_t)run1 * host_frame_bytes, host_frame_bytes, dev, dev_slot_bytes,
                                              width_bytes, run2, cudaMemcpyDeviceToHost, c->stream));
        return SVO_OK;
    }
};
}  // namespace

// ORs the capacity flags of a retired run into the context's accumulator (runs retired without a download)
__global__ void k_accum_flags(const int* __restrict__ flags, int first, int n, int ring, int* __restrict__ accum) {
    int v = 0;
    for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) {
        const int f = i >> 1, side = i & 1;
        int fs = first + f;
        if (fs >= ring) fs -= ring;
        v |= flags[side * ring + fs];
    }
    if (v) atomicOr(accum, v);
}

extern "C" int svo_frontend_check(svo_ctx* c) {
    if (!c) return SVO_ERR_INVALID;
    if (!c->fe_configured || !c->d_flag_accum) return c->fail(SVO_ERR_INVALID, "svo_frontend_configure() first");
    cudaSetDevice(c->device);
    int v = 0;
    SVO_CUDA_TRY(c, cudaMemcpyAsync(&v, c->d_flag_accum, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (v) {
        SVO_CUDA_TRY(c, cudaMemsetAsync(c->d_flag_accum, 0, sizeof(int), c->stream));
        char msg[160];
        snprintf(msg, sizeof msg, "device capacity exceeded in a run retired without download (flags 0x%x: 1 tile, 2 candidates, 4 octree nodes, 8 keypoints, 16 fast)", v);
        return c->fail(SVO_ERR_CAPACITY, msg);
    }
    return SVO_OK;
}

extern "C" int svo_frontend_discard(svo_ctx* c) {
    if (!c) return SVO_ERR_INVALID;
    if (!c->fe_configured || c->fe_nq < 1) return c->fail(SVO_ERR_INVALID, "svo_frontend_discard: nothing to discard");
    c->fe_qhead = (c->fe_qhead + 1) % 2;
    --c->fe_nq;
    return SVO_OK;
}

extern "C" int svo_frontend_download(svo_ctx* c, svo_frontend_out* out) {
    if (!c) return SVO_ERR_INVALID;
    if (!c->fe_configured || c->fe_nq < 1) return c->fail(SVO_ERR_INVALID, "svo_frontend_download: nothing to download");
    cudaSetDevice(c->device);
    // the run is only PEEKED here: it leaves the queue when it has been delivered (or when the device reported an overflow, in
    // which case its results are truncated and useless); a caller buffer that is too small leaves it queued for a retry
    const svo_ctx::DoneRun done = c->fe_q[c->fe_qhead];
    auto pop = [&]() {
        c->fe_qhead = (c->fe_qhead + 1) % 2;
        --c->fe_nq;
    };
    const int ring = c->fe_ring, first = done.first, n = done.n, skip = done.skip;
    if (!out) {
        // results stay on the device (device-resident pipelines): retire the run; its capacity flags are OR-ed into the
        // context's accumulator on the compute stream (which is ordered behind the run), read by svo_frontend_check()
        k_accum_flags<<<1, 128, 0, c->stream>>>(c->d_flags, first, n, ring, c->d_flag_accum);
        ++c->launches;
        SVO_CUDA_TRY(c, cudaGetLastError());
        pop();
        return SVO_OK;
    }
    // all copies go on the download stream, behind the event that marks the end of that run's kernels
    cudaStream_t const compute_stream = c->stream;
    SVO_CUDA_TRY(c, cudaStreamWaitEvent(c->d2h_stream, c->ev_done[done.ev], 0));
    struct StreamSwap {
        svo_ctx* c;
        cudaStream_t saved;
        ~StreamSwap() { c->stream = saved; }
    } swap{c, compute_stream};
    c->stream = c->d2h_stream;
    const int S = c->max_images;
    int rc;
    // ---- counts first (one small copy + sync), then exact-width 2-D copies
    std::vector<int> nkp(S, 0), nfast(S, 0), flags(S, 0), norb(ring, 0), nlkf(ring, 0);
    if (c->fe_mode & SVO_MODE_ORB) SVO_CUDA_TRY(c, cudaMemcpyAsync(norb.data(), c->d_norb, sizeof(int) * ring, cudaMemcpyDeviceToHost, c->stream));
    if (c->fe_mode & SVO_MODE_LK) SVO_CUDA_TRY(c, cudaMemcpyAsync(nlkf.data(), c->d_nlkf, sizeof(int) * ring, cudaMemcpyDeviceToHost, c->stream));
    if (c->fe_mode & SVO_MODE_ORB) SVO_CUDA_TRY(c, cudaMemcpyAsync(nkp.data(), c->d_nkp, sizeof(int) * S, cudaMemcpyDeviceToHost, c->stream));
    if (c->fe_mode & SVO_MODE_LK) SVO_CUDA_TRY(c, cudaMemcpyAsync(nfast.data(), c->d_nfast, sizeof(int) * S, cudaMemcpyDeviceToHost, c->stream));
    SVO_CUDA_TRY(c, cudaMemcpyAsync(flags.data(), c->d_flags, sizeof(int) * S, cudaMemcpyDeviceToHost, c->stream));
    SVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (int f = 0; f < n; ++f) {
        const int fs = (first + f) % ring;
        if (flags[fs] | flags[ring + fs]) {  // (the flags are cleared by the next run that writes these slots)
            pop();
            char msg[128];
            snprintf(msg, sizeof msg, "device capacity exceeded on frame %d (flags 0x%x / 0x%x)", f, flags[fs], flags[ring + fs]);
            return c->fail(SVO_ERR_CAPACITY, msg);
        }
    }
    // ---- validate the caller's capacities BEFORE anything is written or retired
    if ((c->fe_mode & SVO_MODE_ORB) && out->kp_cap < c->kp_cap) return c->fail(SVO_ERR_CAPACITY, "svo_frontend_out.kp_cap < svo_orb_capacity() (run kept queued)");
    if (c->fe_mode & SVO_MODE_LK) {
        if (out->pt_cap < 1) return c->fail(SVO_ERR_CAPACITY, "svo_frontend_out.pt_cap < 1 (run kept queued)");
        for (int f = 0; f < n; ++f) {
            const int fs = (first + f) % ring, ps = (first + f - 1 + ring) % ring;
            const int ns = (f == 0 && skip) ? 0 : nfast[ps];
            if (nfast[fs] > out->pt_cap || ns > out->pt_cap)
                return c->fail(SVO_ERR_CAPACITY, "more FAST corners / LK seeds than svo_frontend_out.pt_cap (run kept queued)");
        }
    }
    RingCopy rcpy{c, first, n, ring};
    if (c->fe_mode & SVO_MODE_ORB) {
        const int K = c->kp_cap;
        int maxL = 0, maxR = 0, maxQ = 0;
        for (int f = 0; f < n; ++f) {
            const int fs = (first + f) % ring, ps = (first + f - 1 + ring) % ring;
            maxL = std::max(maxL, nkp[fs]);
            maxR = std::max(maxR, nkp[ring + fs]);
            const int nq = (f == 0 && skip) ? 0 : nkp[ps];
            maxQ = std::max(maxQ, nq);
            if (out->n_left) out->n_left[f] = nkp[fs];
            if (out->n_right) out->n_right[f] = nkp[ring + fs];
            if (out->n_query) out->n_query[f] = nq;
        }
        const size_t hk = (size_t)out->kp_cap;
        if (out->kps_left && (rc = rcpy.copy(out->kps_left, hk * 28, c->d_kps, (size_t)K * 28, (size_t)maxL * 28))) return rc;
        if (out->desc_left && (rc = rcpy.copy(out->desc_left, hk * 32, c->d_desc, (size_t)K * 32, (size_t)maxL * 32))) return rc;
        if (out->kps_right && (rc = rcpy.copy(out->kps_right, hk * 28, c->d_kps + (size_t)ring * K, (size_t)K * 28, (size_t)maxR * 28))) return rc;
        if (out->desc_right && (rc = rcpy.copy(out->desc_right, hk * 32, c->d_desc + (size_t)ring * K * 32, (size_t)K * 32, (size_t)maxR * 32))) return rc;
        int maxM = 0;
        for (int f = 0; f < n; ++f) {
            const int fs = (first + f) % ring;
            maxM = std::max(maxM, norb[fs]);
            if (out->n_orb_matched) out->n_orb_matched[f] = norb[fs];
        }
        float* oh[3] = {out->orb_t2_left, out->orb_t1_left, out->orb_t1_right};
        for (int r = 0; r < 3; ++r)
            if (oh[r] && (rc = rcpy.copy(oh[r], hk * 8, c->d_orb_pts + (size_t)r * K * 2, (size_t)3 * K * 8, (size_t)maxM * 8))) return rc;
        if (out->orb_points3d && c->have_proj && (rc = rcpy.copy(out->orb_points3d, hk * 12, c->d_orb_xyz, (size_t)K * 12, (size_t)maxM * 12))) return rc;
        if (c->fe_pnp) {
            if (out->orb_pose && (rc = rcpy.copy(out->orb_pose, 128, c->d_pnp_pose, 128, 128))) return rc;
            if (out->orb_pnp_info && (rc = rcpy.copy(out->orb_pnp_info, 16, c->d_pnp_info, 16, 16))) return rc;
            if (out->orb_inlier_mask && (rc = rcpy.copy(out->orb_inlier_mask, hk, c->d_pnp_mask_orb, (size_t)K, (size_t)maxM))) return rc;
        }
        int32_t* mh[4] = {out->stereo_idx, out->stereo_dist, out->temporal_idx, out->temporal_dist};
        for (int r = 0; r < 4; ++r)
            if (mh[r] && (rc = rcpy.copy(mh[r], hk * 4, c->d_match + (size_t)r * K, (size_t)4 * K * 4, (size_t)maxQ * 4))) return rc;
    }
    if (c->fe_mode & SVO_MODE_LK) {
        const int P = c->pt_cap;
        int maxF = 0, maxS = 0;
        for (int f = 0; f < n; ++f) {
            const int fs = (first + f) % ring, ps = (first + f - 1 + ring) % ring;
            const int ns = (f == 0 && skip) ? 0 : nfast[ps];
            maxF = std::max(maxF, nfast[fs]);
            maxS = std::max(maxS, ns);
            if (out->n_fast) out->n_fast[f] = nfast[fs];
            if (out->n_seed) out->n_seed[f] = ns;
        }
        const size_t hp = (size_t)out->pt_cap;
        if (out->fast_kps && (rc = rcpy.copy(out->fast_kps, hp * 28, c->d_fast_kps, (size_t)P * 28, (size_t)maxF * 28))) return rc;
        float* ph[4] = {out->lk_p1, out->lk_p2, out->lk_p3, out->lk_p0r};
        for (int r = 0; r < 4; ++r)
            if (ph[r] && (rc = rcpy.copy(ph[r], hp * 8, c->d_lk_pts + (size_t)(r + 1) * P * 2, (size_t)5 * P * 8, (size_t)maxS * 8))) return rc;
        int maxK = 0;
        for (int f = 0; f < n; ++f) {
            const int fs = (first + f) % ring;
            maxK = std::max(maxK, nlkf[fs]);
            if (out->n_lk_matched) out->n_lk_matched[f] = nlkf[fs];
        }
        float* fh[4] = {out->lkf_t1_left, out->lkf_t1_right, out->lkf_t2_right, out->lkf_t2_left};
        for (int r = 0; r < 4; ++r)
            if (fh[r] && (rc = rcpy.copy(fh[r], hp * 8, c->d_lkf_pts + (size_t)r * P * 2, (size_t)5 * P * 8, (size_t)maxK * 8))) return rc;
        if (out->lk_points3d && c->have_proj && (rc = rcpy.copy(out->lk_points3d, hp * 12, c->d_lkf_xyz, (size_t)P * 12, (size_t)maxK * 12))) return rc;
        if (c->fe_pnp) {
            if (out->lk_pose && (rc = rcpy.copy(out->lk_pose, 128, c->d_pnp_pose + (size_t)16 * ring, 128, 128))) return rc;
            if (out->lk_pnp_info && (rc = rcpy.copy(out->lk_pnp_info, 16, c->d_pnp_info + (size_t)4 * ring, 16, 16))) return rc;
            if (out->lk_inlier_mask && (rc = rcpy.copy(out->lk_inlier_mask, hp, c->d_pnp_mask_lk, (size_t)P, (size_t)maxK))) return rc;
        }
        if (out->lk_status)
            for (int r = 0; r < 4; ++r)
                if ((rc = rcpy.copy(out->lk_status + (size_t)r * hp, 4 * hp, c->d_lk_status + (size_t)r * P, (size_t)4 * P, (size_t)maxS))) return rc;
        if (out->lk_err && c->fe_lk_err)
            for (int r = 0; r < 4; ++r)
                if ((rc = rcpy.copy(out->lk_err + (size_t)r * hp, 16 * hp, c->d_lk_err + (size_t)r * P, (size_t)16 * P, (size_t)maxS * 4))) return rc;
    }
    SVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    pop();
    return SVO_OK;
}

extern "C" int svo_frontend_batch(svo_ctx* c, const uint8_t* left, const uint8_t* right, int n_frames, size_t stride, svo_frontend_out* out) {
    int rc;
    if ((rc = svo_frontend_upload(c, left, right, n_frames, stride)) != SVO_OK) return rc;
    if ((rc = svo_frontend_run(c)) != SVO_OK) return rc;
    return svo_frontend_download(c, out);
}
