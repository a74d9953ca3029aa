// svo_common.cuh -- shared definitions of libsvo_b200.so (B200 / sm_100a stereo front-end).
//
// Data layout in HBM (all buffers are allocated once per context, sized for `max_images` image slots):
//   level l of slot s        : lvl[l].base + s * lvl[l].plane, rows of lvl[l].pitch bytes (pitch multiple of 64,
//                              plane == pitch * h so that consecutive slots form one pitched 2-D array)
//   FAST tile staging        : tile_buf + s * tile_buf_per_slot + tile.out_off   (packed u32 candidates per tile)
//   candidates (cell loop)   : cand + s * cand_per_slot + lvl[l].cand_off        (packed u32, reference order)
//   level keypoints (octree) : lkp  + s * lkp_per_slot  + lvl[l].lkp_off         (packed u32, list order)
//   keypoints / descriptors  : kps + s * kp_cap (28-byte cv::KeyPoint), desc + s * kp_cap * 32
// A packed candidate is  response << 24 | y << 12 | x  with x, y relative to minBorder (= pixel - 16).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "svo_cuda.h"

#define SVO_MAX_LEVELS 12
#define SVO_EDGE 19        // EDGE_THRESHOLD, ORBextractor.cpp:19
#define SVO_MINB 16        // minBorder = EDGE_THRESHOLD - 3, ORBextractor.cpp:726
#define SVO_MAX_DIM 4096   // packed candidate coordinates are 12 bits (relative to minBorder)
#define SVO_LK_LEVELS 4    // maxLevel 3 (tracking.cpp:594)
#define SVO_LK_WIN 21
// REFLECT_101 border materialised around every LK plane (buildOpticalFlowPyramid pads by winSize): the tracker reads up to
// 22 px beyond the image and stages aligned 32-byte rows, hence 32 columns; 24 rows
#define SVO_LK_PADX 32
#define SVO_LK_PADY 24

struct LevelGeom {
    uint8_t* base;  // slot 0 of this level
    int w, h, pitch;
    uint32_t plane;  // bytes per slot
    // cell grid (ORBextractor.cpp:731-740)
    int nCols, nRows, wCell, hCell;
    // octree (ORBextractor.cpp:491-493)
    int N, nIni, Wp, Hp;
    float hX;
    // per-slot offsets (in elements) into cand / lkp arrays
    int cand_off, cand_cap, lkp_off, lkp_cap;
    int tile_begin, tile_count;
    float scale;     // mvScaleFactor[l]
    int patch_size;  // int(31 * scale)
    // resize tables (device), describing level l as resized from level l-1
    const int* xofs;
    const short* alpha;  // 2 per dst x
    const int* yofs;     // 2 per dst y (clamped row indices)
    const short* beta;   // 2 per dst y
    // the same tables in the form the one-pass kernel reads (K1 fused): per 4-pixel group {left tap of pixel 0, byte selectors
    // of pixels 1..3 relative to it, alpha pairs of the 4 pixels, 2 pad words}; per row {top tap row, bottom tap row, beta0, beta1}
    const int4* xmeta;   // 2 per group
    const int4* ymeta;   // 1 per dst y
};

struct PyrGeom {
    int nlevels;
    int iniTh, minTh;
    LevelGeom lvl[SVO_MAX_LEVELS];
};

// One tile of the one-pass pyramid (K1 fused): for every level the pixels this CTA OWNS (writes to HBM) and the slightly
// larger region it NEEDS in shared memory as the source of the next level.  lv[0] is the level-0 region to stage (TMA box).
struct PyrRect {
    short x0, x1o, x1n;  // owned [x0, x1o), needed [x0, x1n)
    short y0, y1o, y1n;
    // thread layout of the level inside the CTA, tabulated on the host: 4-pixel groups per row (from x0 & ~3), row lanes
    // (256 / ng), rows per row lane, shared-memory pitch (4-byte front margin + groups + 8), 65536 / ng + 1
    short ng, nrl, chunk, pitch;
    int magic;
};
struct PyrTile {
    PyrRect lv[SVO_MAX_LEVELS];
};
#define SVO_PYR_BOX_W 256  // TMA box of the staged level-0 region (bytes x rows); boxDim <= 256 per dimension
#define SVO_PYR_BOX_H 96

// One FAST tile = one cell row x a run of cells of one level (K2), or a full-width strip of rows (K9).
struct alignas(16) FastTile {
    short level, cell_i;
    short cell_j0, ncells;
    int x0, x1, y0, y1;  // inner pixel domain [x0,x1) x [y0,y1), level coordinates
    int out_off, out_cap;
};

// batch image index -> image slot (see svo_ctx.cu: plain batches use slots 0..n-1; the frontend uses a ring of frames)
struct SlotMap {
    int sides, side_stride, first, ring;
    // sides is 1 or 2 and first, b / sides < ring (see the four constructions in svo_ctx.cu / svo_frontend.cu): no division
    __host__ __device__ inline int slot(int b) const {
        const int side = sides == 2 ? (b & 1) : 0, f = sides == 2 ? (b >> 1) : b;
        int r = first + f;
        if (r >= ring) r -= ring;
        return side * side_stride + r;
    }
};

#define SVO_FLAG_TILE_OVERFLOW 1
#define SVO_FLAG_CAND_OVERFLOW 2
#define SVO_FLAG_NODE_OVERFLOW 4
#define SVO_FLAG_KP_OVERFLOW 8
#define SVO_FLAG_FAST_OVERFLOW 16

__host__ __device__ inline uint32_t pack_cand(int x, int y, int r) { return ((uint32_t)r << 24) | ((uint32_t)y << 12) | (uint32_t)x; }
__host__ __device__ inline int cand_x(uint32_t c) { return (int)(c & 0xFFFu); }
__host__ __device__ inline int cand_y(uint32_t c) { return (int)((c >> 12) & 0xFFFu); }
__host__ __device__ inline int cand_r(uint32_t c) { return (int)(c >> 24); }

struct LkLevelGeom {
    uint8_t* base;
    int w, h, pitch;
    uint32_t plane;
};

struct svo_ctx {
    int device = 0, max_w = 0, max_h = 0, max_images = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::string err;
    int64_t launches = 0;
    bool timing = false;
    cudaEvent_t (*ev)[SVO_NUM_STAGES][2] = nullptr;  // [SVO_TIMING_SLOTS], created when timing is first enabled
    bool ev_used[SVO_TIMING_SLOTS][SVO_NUM_STAGES] = {};
    int tslot = 0;

    // ---- ORB configuration (host copies of the reference's tables)
    bool configured = false;
    int nfeatures = 0, nlevels = 0, iniTh = 0, minTh = 0;
    float scaleFactor = 0;
    std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
    std::vector<int> features_per_level;
    int umax[16];

    // ---- geometry for the currently laid-out image size
    int cur_w = 0, cur_h = 0;
    PyrGeom geom;
    std::vector<FastTile> tiles;  // all levels
    FastTile* d_tiles = nullptr;
    int n_tiles = 0;
    uint32_t tile_buf_per_slot = 0, cand_per_slot = 0, lkp_per_slot = 0;
    int kp_cap = 0;
    int octree_ncap = 0;

    // ---- one-pass pyramid (K1 fused): tiles, TMA descriptor of level 0, shared-memory budget
    bool fused_pyr = false;
    bool pyr_meta_ok = false;  // the xmeta / ymeta tables fit the one-pass kernel's assumptions
    std::vector<PyrTile> ptiles;
    PyrTile* d_ptiles = nullptr;
    int n_ptiles = 0;
    int pyr_bufB_bytes = 0;
    alignas(64) unsigned char tmap_l0[128] = {};  // CUtensorMap
    // K2: one 3-D tensor map per level ([rows][pitch / 16][16 bytes], box 17 chunks x (hCell + 6) rows)
    alignas(64) unsigned char tmap_fast[SVO_MAX_LEVELS][128] = {};
    bool fast_tma = false;

    // ---- device memory
    uint8_t* d_pyr = nullptr;  // all levels, all slots
    size_t pyr_bytes = 0;
    void* d_tables = nullptr;  // resize tables
    uint32_t* d_tile_buf = nullptr;
    int* d_tile_cnt = nullptr;
    int* d_tile_pos = nullptr;  // position of every tile's list in its level's candidate array (k_compact_scan)
    uint32_t* d_cand = nullptr;
    int* d_ncand = nullptr;     // slot * SVO_MAX_LEVELS + l
    uint16_t* d_knode = nullptr;  // octree scratch, same shape as cand
    uint32_t* d_lkp = nullptr;
    int* d_nlkp = nullptr;
    svo_keypoint* d_kps = nullptr;
    uint8_t* d_desc = nullptr;
    int* d_nkp = nullptr;
    int* d_flags = nullptr;  // per slot
    int* d_flag_accum = nullptr;  // flags of runs retired without a download (svo_frontend_check)
    int8_t* d_pattern = nullptr;

    // generic scratch for the standalone calls (match / fast / lk)
    void* d_scratch = nullptr;
    size_t scratch_bytes = 0;

    // ---- LK / full-frame FAST
    LkLevelGeom lk[SVO_LK_LEVELS];  // [0]: the ORB level-0 plane (unpadded); [1..3]: padded planes, base = interior origin
    LkLevelGeom lkp0;               // padded copy of level 0 (written by k_lk_pad)
    int lk_levels = 0;
    uint8_t* d_lkpyr = nullptr;  // padded planes: level-0 copy + levels 1..3
    // Scharr planes of the 4 LK levels (Ix | Iy << 16, zero border): the tracker's template setup interpolates them (K7c k_lk_deriv)
    LkLevelGeom lkd[SVO_LK_LEVELS];
    uint8_t* d_lkderiv = nullptr;
    // one-warp CTAs per frame of the chain launch; fewer than pt_cap makes the kernel grid-stride over the frame's points
    // (SVO_LK_GRID; measured: 2048 is -0.4 % at 1241x376 / 128 frames but +8 % at 3840x2160 / 8 frames, where a CTA then walks ~10
    // points and the last wave of CTAs runs half empty -- empty CTAs cost nothing measurable, so the default is one CTA per slot)
    int lk_grid_x = 1 << 30;
    bool lk_deriv = false;  // SVO_LK_DERIV=1 in the environment: template setup from Scharr planes (parity-green, but 4 % slower: more L1 / L2 traffic)
    std::vector<FastTile> ftiles;
    FastTile* d_ftiles = nullptr;
    int n_ftiles = 0;
    uint32_t* d_ftile_buf = nullptr;
    int* d_ftile_cnt = nullptr;
    uint32_t ftile_buf_per_slot = 0;
    svo_keypoint* d_fast_kps = nullptr;  // per frame slot, pt_cap each
    int* d_nfast = nullptr;
    int pt_cap = 0;
    float* d_lk_pts = nullptr;      // per frame slot: 5 lists (p0,p1,p2,p3,p0r) x pt_cap x 2
    uint8_t* d_lk_status = nullptr;  // per frame slot: 4 x pt_cap
    float* d_lk_err = nullptr;

    // ---- device-side post-filters (K10 / K11): per frame slot
    float* d_orb_pts = nullptr;  // [frame slot][3: t2_left, t1_left, t1_right][kp_cap][2]
    int* d_norb = nullptr;
    float* d_lkf_pts = nullptr;  // [frame slot][5: t1_left, t1_right, t2_right, t2_left, t1_left_returned][pt_cap][2]
    int* d_nlkf = nullptr;
    double fe_match_error = 3.0;  // feature_match_error (config/default.yaml:66)
    // K12: triangulated points of the filtered lists (sensors_->projMatr1_/2_, sensors.cpp:8-24)
    bool have_proj = false;
    double proj1[12] = {}, proj2[12] = {};
    float* d_orb_xyz = nullptr;  // [frame slot][kp_cap][3]
    float* d_lkf_xyz = nullptr;  // [frame slot][pt_cap][3]

    // K13: pose of every unit frame from its filtered lists (cv::solvePnPRansac, tracking.cpp:464-501); parameters as in
    // config/default.yaml:80-82.  Results per list kind (0 ORB, 1 LK) and frame slot.
    bool fe_pnp = false;
    int pnp_iters = 500;
    float pnp_reproj = 0.5f;
    double pnp_conf = (double)0.99f;
    uint32_t* d_pnp_rng = nullptr;   // the raw stream of cv::RNG((uint64)-1): the same for every problem
    int pnp_rng_n = 0;
    uint64_t pnp_rng_end = 0;
    double* d_pnp_pose = nullptr;    // [kind][frame slot][16]: rvec, tvec, R row-major, spare
    int* d_pnp_info = nullptr;       // [kind][frame slot][4]: ok, inliers, best hypothesis, hypotheses evaluated
    uint8_t* d_pnp_mask_orb = nullptr;  // [frame slot][kp_cap]
    uint8_t* d_pnp_mask_lk = nullptr;   // [frame slot][pt_cap]

    // ---- matches (frontend): per frame slot, kp_cap each
    int* d_match = nullptr;  // [frame slot][4][kp_cap]: stereo idx, stereo dist, temporal idx, temporal dist
    // K6 on the int8 tensor-core path: descriptors expanded to one byte per bit in IMMA fragment order, their popcounts, slot list
    uint32_t* d_xdesc = nullptr;  // [image slot][kp_cap][64 words]
    int* d_xpop = nullptr;        // [image slot][kp_cap]
    int* d_xslots = nullptr;
    size_t xslots_cap = 0;
    bool match_mma = true;        // SVO_MATCH_POPC=1 in the environment selects the POPC kernel (A/B measurements)
    bool match_tc = true;         // K6 on tcgen05 (k_match_tc.cu); SVO_MATCH_IMMA=1 / SVO_MATCH_POPC=1 select the older forms

    // ---- frontend state
    int fe_mode = 0, fe_max_frames = 0, fe_ring = 0, fe_head = -1, fe_pending = 0, fe_first = 0;
    struct DoneRun {
        int first, n, skip, ev;
    };
    DoneRun fe_q[2] = {};  // completed, not yet downloaded runs (oldest at fe_qhead)
    int fe_qhead = 0, fe_nq = 0, fe_depth = 1;
    cudaStream_t d2h_stream = nullptr;
    // ORB tail || FAST + LK chain: side stream and the fork / join events
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool fe_overlap = true;
    bool fe_lk_err = true;  // the reference passes `err` to calcOpticalFlowPyrLK (tracking.cpp:593): the chain computes it too
    cudaEvent_t ev_done[2] = {};
    bool fe_has_last = false;
    bool fe_configured = false;
    // upload staging: linear H2D copies on a copy stream, double-buffered; rows are re-pitched into the ring by a kernel
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copy[2] = {}, ev_repitch[2] = {};
    bool ev_repitch_valid[2] = {false, false};
    uint8_t* d_stage[2] = {nullptr, nullptr};
    size_t stage_bytes = 0;
    int stage_idx = 0;
    const uint8_t* pend_left = nullptr;
    const uint8_t* pend_right = nullptr;
    size_t pend_stride = 0;
    bool pend_staged = false;
    bool aux_ready = false;        // LK pyramid / full-frame FAST buffers laid out for the current size
    bool user_configured = false;  // svo_orb_configure() was called by the user (not the detector-only default)

    int fail(int code, const std::string& msg) {
        err = msg;
        return code;
    }
};

#define SVO_CUDA_TRY(ctx, expr)                                                                          \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            return (ctx)->fail(SVO_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));        \
        }                                                                                                \
    } while (0)

struct StageTimer {
    svo_ctx* c;
    int s;
    StageTimer(svo_ctx* ctx, int stage) : c(ctx), s(stage) {
        if (c->timing) cudaEventRecord(c->ev[c->tslot][s][0], c->stream);
    }
    ~StageTimer() {
        if (c->timing) {
            cudaEventRecord(c->ev[c->tslot][s][1], c->stream);
            c->ev_used[c->tslot][s] = true;
        }
    }
};

// ---- kernel launchers (one per stage; defined in k_*.cu).  n_img batch images mapped to slots through `map`.
int launch_pyramid(svo_ctx* c, const SlotMap& map, int n_img);                       // K1
int launch_repitch(svo_ctx* c, const uint8_t* left, const uint8_t* right, size_t stride, int first, int ring, int n_frames);
int launch_fast_cells(svo_ctx* c, const SlotMap& map, int n_img);                    // K2 + K2b
int launch_octree(svo_ctx* c, const SlotMap& map, int n_img);                        // K3
int launch_describe(svo_ctx* c, const SlotMap& map, int n_img);                      // K4 + K5
int launch_match_frontend(svo_ctx* c, int first, int ring, int n_frames, int skip_first);  // K6
int launch_match_tc_frontend(svo_ctx* c, int first, int ring, int n_frames, int skip_first);             // K6, tcgen05 form (k_match_tc.cu)
int launch_match_tc_one(svo_ctx* c, const uint8_t* dq, int nq, const uint8_t* dt, int nt, int* di, int* dd);
int launch_orb_filter(svo_ctx* c, int first, int ring, int n_frames, int skip_first);      // K10
int launch_lk_filter(svo_ctx* c, int first, int ring, int n_frames, int skip_first);       // K11
int launch_triangulate_frontend(svo_ctx* c, int first, int ring, int n_frames);            // K12
int launch_pnp_frontend(svo_ctx* c, int first, int ring, int n_frames);                    // K13
int svo_pnp_free_frontend(svo_ctx* c);
int launch_fast_full(svo_ctx* c, const SlotMap& map, int n_img, int threshold);      // K9 (+ ordered compaction)
int launch_lk_pyramid(svo_ctx* c, const SlotMap& map, int n_img);                    // K7
// the 4-call LK chain for n_frames unit frames in ONE launch (K8); see k_lk.cu
struct LkCall {
    int prev_side, prev_back;  // prev image: side (0 left, 1 right) of frame (current - prev_back)
    int next_side, next_back;
    int src_list, dst_list, status_row;
};
int launch_lk_chain(svo_ctx* c, int first, int ring, int n_frames, int skip_first, const LkCall* calls, int ncalls, int max_iter, double eps,
                    double min_eig, int want_err);

int svo_orb_run(svo_ctx* c, const SlotMap& map, int n_img);
int svo_check_flags(svo_ctx* c, const SlotMap& map, int n_img);
int svo_ensure_scratch(svo_ctx* c, size_t bytes);
int svo_standalone_guard(svo_ctx* c);

// ---- device allocations -------------------------------------------------------------------------------------------------------
// Every device buffer of the library comes from svo_dev_malloc.  With SVO_GUARD=1 in the environment each buffer gets a 4 KB band
// of 0xA5 bytes on both sides and svo_guard_check() (C ABI, debug) counts the band bytes that changed: the write-overrun check the
// GPU tests run (the pool's compute-sanitizer is closed).  Without the variable it is plain cudaMalloc / cudaFree.
cudaError_t svo_dev_malloc_bytes(void** p, size_t bytes);
cudaError_t svo_dev_free(void* p);
template <class T>
inline cudaError_t svo_dev_malloc(T** p, size_t bytes) { return svo_dev_malloc_bytes(reinterpret_cast<void**>(p), bytes); }


// host helpers
int svo_layout(svo_ctx* c, int w, int h);
size_t svo_align(size_t v, size_t a);
