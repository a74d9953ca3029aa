#!/usr/bin/env python
"""bench.py -- stereo pairs/s of the front-end (ORB + Hamming match + FAST + pyramidal LK).

One unit of work = one new stereo frame t:
  ORB mode : ORBextractor::operator() on left and right (8 levels, FAST 20/7), match(last_left, last_right),
             match(last_left, cur_left)                                              (BASELINE.json configs[0])
  LK mode  : cv::FAST(left, 20) and the four chained calcOpticalFlowPyrLK calls      (BASELINE.json configs[1])
and every step runs BOTH modes on its frames.

--config c1 (default): 1241x376, 1000 features, one step = one batch of 128 consecutive frames per GPU, weak scaling  (configs[0]+[1])
--config c3          : 1920x1080, 2000 features, 32 frames per step per GPU, weak scaling                              (configs[2])
--config c4          : 3840x2160, 4000 features, 8 frames per step per GPU, weak scaling                               (configs[3])
--config c5          : ONE 10,001-frame 1241x376 sequence (10,000 units); the units are cut into contiguous chunks, one per GPU,
                       overlapping by one frame (sharding.shard_units); one step = one sweep over the whole sequence, STRONG scaling;
                       a per-unit signature of the results is gathered to rank 0 and hashed, so runs at different N can be compared
                                                                                                                      (configs[4])

`value`  : device-timed (CUDA events on the launching stream), images already resident in HBM, max over ranks.
`e2e`    : the same work through svo_frontend_upload / run / download with pinned HOST buffers, H2D and D2H inside the timed region.
`--impl reference` times the reference's CPU implementation of the same path on all host cores (oracle/ref_pipeline.py).
Multi-GPU: one process per GPU (torchrun), frames sharded, no collective on the data path.
"""
import argparse
import hashlib
import importlib
import json
import os
import subprocess
import sys
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    "c1": dict(width=1241, height=376, nfeatures=1000, frames=128, seed=1000, scaling="weak", ref_units_per_core=32,
               name="C1+C2", metric="stereo pairs/s (ORB+match+LK) @1241x376"),
    "c3": dict(width=1920, height=1080, nfeatures=2000, frames=32, seed=2000, scaling="weak", ref_units_per_core=8,
               name="C3", metric="stereo pairs/s (ORB+match+LK) @1920x1080"),
    "c4": dict(width=3840, height=2160, nfeatures=4000, frames=8, seed=3000, scaling="weak", ref_units_per_core=3,
               name="C4", metric="stereo pairs/s (ORB+match+LK) @3840x2160"),
    "c5": dict(width=1241, height=376, nfeatures=1000, frames=128, seed=1000, scaling="strong", ref_units_per_core=32, sweep_units=10000,
               name="C5", metric="stereo pairs/s (ORB+match+LK) @1241x376, one 10k-pair sequence sharded over the GPUs"),
}


KITTI_P1 = [718.856, 0, 607.1928, 0, 0, 718.856, 185.2157, 0, 0, 0, 1, 0]          # P0 / P1 of KITTI odometry sequence 00
KITTI_P2 = [718.856, 0, 607.1928, -386.1448, 0, 718.856, 185.2157, 0, 0, 0, 1, 0]


def workload_config(a):
    c = {"workload": f"{a.cfg['name']}: ORB({a.nfeatures} feats, 8 levels, FAST 20/7) on L+R + 2x BF-Hamming + full-frame FAST(20) + "
                     f"4-call pyrLK(21x21, 4 levels, 30 it, err computed like the reference's call), synthetic stereo sequence seed {a.seed}",
         "config": a.config, "width": a.width, "height": a.height, "frames_per_step_per_gpu": a.frames, "nfeatures": a.nfeatures,
         "l2_policy": "inputs larger than L2: each step streams a different batch of images (>= 59 MB at the smallest config) plus the "
                      "pyramid levels and scratch it writes (~250 MB per step at c1)"}
    if getattr(a, "pnp", False):
        c["workload"] += " + triangulation + PnP-RANSAC (500 it, 0.5 px, 0.99; EPnP hypotheses, LM refinement) on both filtered lists"
        c["pnp"] = True
    if a.cfg["scaling"] == "strong":
        c["sweep_units"] = a.sweep_units
        c["frames_per_step_per_gpu"] = None
        c["batch_frames"] = a.frames
        c["sharding"] = "contiguous unit ranges per rank, chunks overlap by one frame (re-extracted), no collective on the data path"
    return c


def algorithmic_bytes_per_image(w, h, nlevels=8, scale=1.2):
    """read L0 + write L1..7 + FAST reads every level once."""
    s, px = np.float32(1.0), []
    for _ in range(nlevels):
        inv = np.float32(1.0) / s
        px.append(int(np.rint(np.float32(w) * inv)) * int(np.rint(np.float32(h) * inv)))
        s = np.float32(s * np.float32(scale))
    return px[0] + sum(px[1:]) + sum(px)


class ClockSampler(threading.Thread):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag, self.proc = gpu_index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i",
                                          str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        busy = [v for v in sm if v > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "covers": "device-timed loop and e2e loop"}


def pinned_u8(shape):
    import torch
    return torch.empty(shape, dtype=torch.uint8, pin_memory=True)


def synth_into(pkg, seed, first, count, w, h, nl, nr, threads=16):
    """frames [first, first + count) of the synthetic sequence into nl / nr (count, h, w); the generator releases the GIL"""
    S = pkg.synth_lib()

    def work(t):
        for i in range(t, count, threads):
            rc = S.svo_synth_stereo_frame(seed, first + i, w, h, nl[i].ctypes.data, nr[i].ctypes.data, w)
            assert rc == 0
    ts = [threading.Thread(target=work, args=(t,)) for t in range(min(threads, count))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()


def unit_signature(out, i):
    """CRC of everything one unit delivers to the host (counts, keypoints, descriptors, matches, tracks): runs with different
    shardings must produce the same per-unit signatures."""
    nl, nr, nq = int(out.n_left[i]), int(out.n_right[i]), int(out.n_query[i])
    ns, nm, nk, nf = int(out.n_seed[i]), int(out.n_orb_matched[i]), int(out.n_lk_matched[i]), int(out.n_fast[i])
    c = zlib.crc32(np.array([nl, nr, nq, ns, nm, nk, nf], np.int32).tobytes())
    for arr, n in ((out.kps_left[i], nl), (out.desc_left[i], nl), (out.kps_right[i], nr), (out.desc_right[i], nr), (out.stereo_idx[i], nq),
                   (out.stereo_dist[i], nq), (out.temporal_idx[i], nq), (out.temporal_dist[i], nq), (out.fast_kps[i], nf),
                   (out.lk_p1[i], ns), (out.lk_p2[i], ns), (out.lk_p3[i], ns), (out.lk_p0r[i], ns), (out.orb_t2_left[i], nm),
                   (out.lkf_t2_left[i], nk)):
        c = zlib.crc32(np.ascontiguousarray(arr[:n]).tobytes(), c)
    return zlib.crc32(np.ascontiguousarray(out.lk_status[i][:, :ns]).tobytes(), c)


def run_b200(a):
    import torch
    import __graft_entry__ as g
    pkg = g.load_package()
    sh = importlib.import_module("svo_b200.sharding")
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # the version banner goes to stdout, in front of the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))  # barrier / max-reduce / signature gather only
    torch.cuda.set_device(local)
    w, h, B, K, W = a.width, a.height, a.frames, a.steps, a.warmup
    strong = a.cfg["scaling"] == "strong"

    # ---- CPU baseline on rank 0 at N=1, before the GPU is busy (bounded sample, separate process)
    cpu_baseline = None
    if world == 1 and not a.no_cpu_baseline:
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3", "--warmup", "1",
                                  "--config", a.config, "--width", str(w), "--height", str(h), "--nfeatures", str(a.nfeatures)] + (["--pnp"] if a.pnp else []),
                                 capture_output=True, text=True, timeout=900)
            cpu_baseline = json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as e:  # the baseline is a reported number, never a reason to lose the GPU measurement
            cpu_baseline = {"value": None, "unit": "stereo pairs/s", "error": repr(e)[:200]}

    # ---- synthetic input pool (pinned host + device copies)
    #  weak  : NB batches of B consecutive frames per rank, cycled through
    #  strong: the rank's chunk of the ONE sequence (its units plus the overlap frame)
    if strong:
        first_frame, n_frames = sh.shard_units(a.sweep_units + 1, world, rank)
        runs = sh.batches(first_frame, n_frames, B)  # (first frame, count) of every svo_frontend batch of a sweep
        units_rank = max(n_frames - 1, 0)
    else:
        NB = 3
        first_frame, n_frames = rank * 4096, NB * B
        units_rank = B
    host_l, host_r = pinned_u8((n_frames, h, w)), pinned_u8((n_frames, h, w))
    nl, nr = host_l.numpy(), host_r.numpy()
    synth_into(pkg, a.seed, first_frame, n_frames, w, h, nl, nr)
    dev_l, dev_r = host_l.cuda(non_blocking=True), host_r.cuda(non_blocking=True)
    torch.cuda.synchronize()

    ctx = pkg.Context(local, w, h, 2 * (2 * B + 1))  # ring of 2 batches + the carried frame: downloads overlap the next run
    ctx.orb_configure(a.nfeatures, 1.2, 8, 20, 7)
    ctx.frontend_configure(w, h, pkg.MODE_ORB | pkg.MODE_LK, B)
    ctx.frontend_set_overlap(not a.no_overlap)
    if a.pnp:  # the pose step of the reference loop (tracking.cpp:190-200): KITTI sequence-00 projection matrices
        ctx.frontend_set_projection(KITTI_P1, KITTI_P2)
        ctx.frontend_set_pnp(True, 500, 0.5)
    stream = torch.cuda.Stream()  # a real (non-null) stream shared by torch events and the library's kernels
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    kp_cap, pt_cap = ctx.orb_capacity(w, h), max(1024, w * h // 64)
    plane = w * h

    def batch_of(k):  # weak scaling: the k-th step's batch
        b = k % NB
        return b * B, B

    def step_resident(k):
        if strong:
            ctx.frontend_reset()  # every sweep starts with the rank's overlap frame (detect only)
            for (f, n) in runs:
                o = (f - first_frame) * plane
                ctx.frontend_upload_dev(dev_l.data_ptr() + o, dev_r.data_ptr() + o, n, w)
                ctx.frontend_run()
                ctx.frontend_download(None)
        else:
            f, n = batch_of(k)
            ctx.frontend_upload_dev(dev_l.data_ptr() + f * plane, dev_r.data_ptr() + f * plane, n, w)
            ctx.frontend_run()
            ctx.frontend_download(None)  # results stay in HBM: retire the run (host bookkeeping + the flag accumulation kernel)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- value: device-timed, inputs resident in HBM
    for k in range(W):
        step_resident(k)
    barrier()
    ctx.frontend_check()
    sampler = ClockSampler(local)
    sampler.start()
    t_wait = time.time()
    while not sampler.rows and time.time() - t_wait < 3.0:  # nvidia-smi takes a moment to produce its first sample
        time.sleep(0.05)
    ctx.enable_stage_timing(True)
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for k in range(K):
        ctx.set_timing_slot(k % pkg.TIMING_SLOTS)
        step_resident(W + k)
    e1.record(stream)
    barrier()
    ms_dev = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    ctx.frontend_check()  # the runs above were retired without a download: a capacity overflow must not go unnoticed
    stage = {s: 0.0 for s in pkg.STAGES}
    nslots = min(K, pkg.TIMING_SLOTS)
    for k in range(nslots):
        for s, v in ctx.stage_ms(k).items():
            stage[s] += max(v, 0.0) / nslots
    ctx.enable_stage_timing(False)
    runs_per_slot = len(runs) if strong else 1  # strong: the stage events of a slot hold the LAST batch of the sweep

    # ---- e2e: host buffers in, host buffers out, through the public upload / run / download calls
    out = pkg.FrontendBuffers(B, kp_cap, pt_cap, alloc=lambda shape, dt: pinned_u8((int(np.prod(shape)) * np.dtype(dt).itemsize,)).numpy()
                              .view(dt).reshape(shape))
    out.pnp_enabled = bool(a.pnp)
    d2h_total = [0]
    sigs = {}

    def views(f, n):
        return nl[f:f + n], nr[f:f + n]

    def sweep_e2e(collect=False):
        # software-pipelined use of the public API (pinned host memory): while batch k+1 runs on the GPU, batch k+2 is uploaded
        # (copy stream) and the results of batch k are downloaded (download stream).  Every batch is uploaded, processed and
        # downloaded inside the timed region.
        ctx.frontend_reset()
        seq = [(f - first_frame, n) for (f, n) in runs]
        ctx.frontend_upload(*views(*seq[0]))
        ctx.frontend_run()
        for i in range(len(seq)):
            if i + 1 < len(seq):
                ctx.frontend_upload(*views(*seq[i + 1]))
                ctx.frontend_run()
            ctx.frontend_download(out)
            d2h_total[0] += out.copied_bytes()
            if collect:
                f, n = seq[i]
                for j in range(n):
                    if f + j > 0:  # the chunk's first frame only warms up the carried frame
                        sigs[first_frame + f + j - 1] = unit_signature(out, j)

    if strong:
        sweep_e2e(collect=True)  # also the warm-up; the signatures are gathered after the timed region
        for k in range(max(W - 1, 0)):
            sweep_e2e()
        barrier()
        d2h_total[0] = 0
        t0 = time.perf_counter()
        for k in range(K):
            sweep_e2e()
        barrier()
        ms_e2e = (time.perf_counter() - t0) * 1e3
        d2h = d2h_total[0] // max(K, 1)
        h2d = 2 * n_frames * plane
    else:
        ctx.frontend_reset()

        def step_e2e(k):
            ctx.frontend_run()                      # batch k + 1
            ctx.frontend_upload(*views(*batch_of(k + 2)))
            ctx.frontend_download(out)              # batch k

        ctx.frontend_upload(*views(*batch_of(0)))
        ctx.frontend_run()
        ctx.frontend_upload(*views(*batch_of(1)))
        for k in range(W):
            step_e2e(k)
        barrier()
        t0 = time.perf_counter()
        for k in range(K):
            step_e2e(W + k)
        barrier()
        ms_e2e = (time.perf_counter() - t0) * 1e3
        d2h = out.copied_bytes()
        h2d = 2 * B * plane
    clocks = sampler.finish()
    tracks_last = 4 * int(np.sum(out.n_seed))  # points handed to each of the 4 calls in the last downloaded batch

    signature = None
    if world > 1:
        t = torch.tensor([ms_dev, ms_e2e], device="cuda", dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms_dev, ms_e2e = float(t[0]), float(t[1])
    if strong:  # per-unit signatures -> rank 0 (bookkeeping after the timed region; the data path itself has no collective)
        mine = sorted(sigs.items())
        if world > 1:
            gathered = [None] * world if rank == 0 else None
            torch.distributed.gather_object(mine, gathered, dst=0)
        else:
            gathered = [mine]
        if rank == 0:
            merged = {}
            for part in gathered:
                merged.update(dict(part))
            assert sorted(merged) == list(range(a.sweep_units)), "units missing after the merge"
            signature = hashlib.sha256(np.array([merged[u] for u in range(a.sweep_units)], np.uint32).tobytes()).hexdigest()[:32]
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak, peak_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        bpi = algorithmic_bytes_per_image(w, h)
        t_pf = stage["pyramid"] + stage["fast_cells"] + stage["compact"]
        imgs_per_launch = 2 * (runs[-1][1] if strong else B)
        achieved = bpi * imgs_per_launch / (t_pf * 1e-3) / 1e9 if t_pf > 0 else None
        t_lk = sum(stage[f"lk_track{i}"] for i in (1, 2, 3, 4))
        traffic, traffic_src, lk_instr = None, None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if (tj.get("width"), tj.get("height")) in ((w, h), (None, None)):
                traffic = tj.get("pyramid_fast_dram_bytes_per_step") * imgs_per_launch / tj.get("images_per_step", 128)
                traffic_src = tj.get("source")
            lk_instr = tj.get("lk_warp_instr_per_track")
        except Exception:
            pass
        units_per_step = a.sweep_units if strong else world * B
        ms_step = ms_dev / K
        issue_peak = 148 * 4 * (clocks.get("sm_mhz") or 1965.0) * 1e6  # warp-instructions per second: 4 schedulers per SM, 1 per clock
        tracks_per_s = tracks_last / (t_lk * 1e-3) if (t_lk > 0 and not strong) else None
        line = {
            "metric": a.cfg["metric"], "value": K * units_per_step / (ms_dev * 1e-3), "unit": "stereo pairs/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": a.cfg["scaling"], "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(a),
            "e2e": {"value": K * units_per_step / (ms_e2e * 1e-3), "unit": "stereo pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / K, "bytes_note": "per rank" if world > 1 else "whole job"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src + " (of measured)" if "hbm_gbs" in peaks else peak_src,
                         "kernels": "K1 k_pyramid_fused (one-pass, TMA-staged) + K2 k_fast_cells (3-D TMA window) + K2b k_compact (pyramid + FAST stages)",
                         "algorithmic_bytes_per_launch": bpi * imgs_per_launch, "ms_per_launch": t_pf,
                         "launch_note": ("strong scaling: the stage events of a timing slot hold the LAST batch of each sweep" if strong else None)},
            "stage_ms_per_step": {k: round(v, 4) for k, v in stage.items()},
            "stage_note": ("event-to-event times per kernel group; pyramid / fast_cells / compact run before the stream fork and lk_track1 "
                           "(the whole 4-call chain) on the main stream; octree / describe / match (side stream) and fast_full / lk_pyramid "
                           "run concurrently, so their figures include waiting for SMs" if not a.no_overlap else "single stream")
                          + (f"; strong scaling: per-batch figures of the last of {runs_per_slot} batches per sweep" if strong else ""),
            # the kernel with the largest share of the step is the LK tracker; it is bound by instruction issue and the
            # shared-memory pipe, not by HBM or tensor throughput (profiles/*_ncu_full_all_kernels.txt), so it is reported as
            # tracks/s and as a fraction of the machine's instruction-issue peak next to the HBM roofline of the pyramid + FAST stages
            "dominant_kernel": {"name": "K8 k_lk_track (the 4-call chain, one launch per batch)", "ms_per_step": round(t_lk * (runs_per_slot if strong else 1), 4),
                                "share_of_step": round(t_lk * (runs_per_slot if strong else 1) / ms_step, 3),
                                "tracks_per_step": tracks_last if not strong else None, "tracks_per_s": tracks_per_s,
                                "warp_instr_per_track": lk_instr,
                                "issue_roofline_frac": (lk_instr * tracks_per_s / issue_peak) if (lk_instr and tracks_per_s) else None,
                                "issue_peak_note": "148 SMs x 4 schedulers x SM clock warp-instructions/s; warp_instr_per_track from the ncu capture "
                                                   "named in roofline.traffic_source",
                                "bound": "instruction issue + shared-memory pipe (integer; no HBM or tensor bound)"},
            "cpu_baseline": cpu_baseline,
        }
        if strong:
            line["merged_signature"] = signature
            line["units_per_rank"] = units_rank
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        torch.distributed.destroy_process_group()


def run_reference(a):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import __graft_entry__ as g
    pkg = g.load_package()  # python module only: the frames come from libsvo_synth.so (host code), the CUDA library is never loaded
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    g.load_oracle().lib()
    import ref_pipeline
    cores = len(os.sched_getaffinity(0))
    w, h = a.width, a.height
    upc = a.cfg["ref_units_per_core"]
    units = a.sample_units or cores * upc  # bounded sample; >= 32 units per worker at c1 so that the warm-up frame is ~3 % of the work
    left = np.empty((units + 1, h, w), np.uint8)
    right = np.empty((units + 1, h, w), np.uint8)
    synth_into(pkg, a.seed, 0, units + 1, w, h, left, right, threads=min(cores, 16))
    # three variants, the arm reports the fastest.  A short calibration pass picks it.
    variants = {
        "ref_allcores": dict(variant="ref", workers=cores, cv_threads=1),          # (3) one single-threaded worker per core, reference-source ORB
        "cv2prim_allcores": dict(variant="cv2prim", workers=cores, cv_threads=1),  # (3) with real-OpenCV image primitives inside ORB
        "port_allcores": dict(variant="port", workers=cores, cv_threads=1),        # (1) x cores: the C++ oracle port of ORB
        "cv2prim_cvthreads": dict(variant="cv2prim", workers=1, cv_threads=cores), # (2) one process, OpenCV's own thread pool on all cores
    }
    cal_units = min(units, max(cores * 2, 4))
    speeds = {}
    for name, v in variants.items():
        n = cal_units if v["workers"] > 1 else min(cal_units, 6)
        pool = ref_pipeline.ReferencePool(left[:n + 1], right[:n + 1], 7 if a.pnp else 3, a.nfeatures, v["workers"], variant=v["variant"], cv_threads=v["cv_threads"])
        dt, _ = pool.run()
        pool.close()
        speeds[name] = n / dt
    best = max(speeds, key=speeds.get)
    v = variants[best]
    pool = ref_pipeline.ReferencePool(left, right, 7 if a.pnp else 3, a.nfeatures, v["workers"], variant=v["variant"], cv_threads=v["cv_threads"])
    for _ in range(a.warmup):
        pool.run()
    total = 0.0
    for _ in range(a.steps):
        dt, _cs = pool.run()
        total += dt
    workers = pool.workers
    pool.close()
    val = a.steps * units / total
    kind = "reference" if (best == "ref_allcores" and ref_pipeline.orb_kind() == "reference") else "port"
    sample = (f"{units} consecutive stereo frames of the same synthetic sequence per step ({units // max(workers, 1)} per worker + 1 warm-up frame each), "
              f"variant {best}: {workers} worker processes x {v['cv_threads']} OpenCV thread(s); cv2 4.13 FAST / BFMatcher / calcOpticalFlowPyrLK "
              f"(err computed) + ORBextractor as in the variant's name (ref = the reference's own src/ORBextractor.cpp compiled into oracle/_ref, "
              f"cv2prim = cv2.resize / cv2.FAST per cell / cv2.GaussianBlur + the oracle's loops, port = C++ oracle port)"
              + ("; pose step with the real cv2.triangulatePoints / convertPointsFromHomogeneous / solvePnPRansac on both filtered lists" if a.pnp else ""))
    cpu_model = ""
    try:
        cpu_model = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        pass
    line = {"impl": "reference", "metric": a.cfg["metric"], "value": val, "unit": "stereo pairs/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": total / a.steps * 1e3, "higher_is_better": True, "scaling": a.cfg["scaling"], "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": workload_config(a),
            "cpu_baseline": {"value": val, "unit": "stereo pairs/s", "cores": cores, "kind": kind, "sample": sample,
                             "pairs_per_s_per_core": val / cores, "cpu_model": cpu_model, "variant": best,
                             "variants_calibration_pairs_per_s": {k: round(s, 2) for k, s in speeds.items()}},
            "e2e": {"value": val, "unit": "stereo pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=None)
    p.add_argument("--warmup", type=int, default=None)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--config", default="c1", choices=sorted(CONFIGS))
    p.add_argument("--frames", type=int, default=None, help="stereo frames per batch per GPU")
    p.add_argument("--width", type=int, default=None)
    p.add_argument("--height", type=int, default=None)
    p.add_argument("--nfeatures", type=int, default=None)
    p.add_argument("--sweep-units", type=int, default=None, help="c5: units (stereo pairs) of the one sequence")
    p.add_argument("--sample-units", type=int, default=0, help="reference arm: stereo frames per step (0 = per-config default per core)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--pnp", action="store_true", help="also run triangulation + PnP-RANSAC on the filtered lists (SURVEY 8f rows 1-2)")
    p.add_argument("--no-overlap", action="store_true", help="run the ORB tail and the FAST + LK chain on one stream")
    a = p.parse_args()
    a.cfg = CONFIGS[a.config]
    for k in ("width", "height", "nfeatures", "frames"):
        if getattr(a, k) is None:
            setattr(a, k, a.cfg[k])
    a.seed = a.cfg["seed"]
    if a.sweep_units is None:
        a.sweep_units = a.cfg.get("sweep_units", 0)
    strong = a.cfg["scaling"] == "strong"
    if a.steps is None:
        a.steps = 3 if strong else 50
    if a.warmup is None:
        a.warmup = 5 if not strong else 3
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
