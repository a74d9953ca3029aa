#!/usr/bin/env python
"""bench.py -- stereo pairs/s of the front-end (ORB + Hamming match + FAST + pyramidal LK) at 1241x376.

One step = one pass of the hot path over one batch of `--frames` consecutive synthetic stereo frames per GPU:
  ORB mode : ORBextractor::operator() on left and right (1000 features, 8 levels), match(last_left, last_right),
             match(last_left, cur_left)                                              (BASELINE.json configs[0])
  LK mode  : cv::FAST(left, 20) and the four chained calcOpticalFlowPyrLK calls      (BASELINE.json configs[1])
`value`  : device-timed (CUDA events on the launching stream), images already resident in HBM, max over ranks.
`e2e`    : the same work through svo_frontend_batch with pinned HOST buffers, H2D and D2H inside the timed region.
`--impl reference` times the reference's CPU implementation of the same path on all host cores (oracle/ref_pipeline.py).
Multi-GPU: one process per GPU (torchrun), frames sharded, no collective on the data path (weak scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "stereo pairs/s (ORB+match+LK) @1241x376"
SEED = 1000


def workload_config(a):
    return {"workload": f"C1+C2: ORB({a.nfeatures} feats, 8 levels, FAST 20/7) on L+R + 2x BF-Hamming + full-frame FAST(20) + "
                        f"4-call pyrLK(21x21, 4 levels, 30 it), synthetic KITTI-shaped sequence seed {SEED}",
            "width": a.width, "height": a.height, "frames_per_step_per_gpu": a.frames, "nfeatures": a.nfeatures,
            "l2_policy": "inputs larger than L2: each step streams a different batch (>= 59 MB of images, ~250 MB of pyramid "
                         "levels and scratch per step at the default size)"}


def algorithmic_bytes_per_image(w, h, nlevels=8, scale=1.2):
    """read L0 + write L1..7 + FAST reads every level once."""
    s, inv, px = np.float32(1.0), [], []
    for l in range(nlevels):
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
                "reasons": sorted(reasons), "samples": len(sm)}


def pinned_u8(shape):
    import torch
    return torch.empty(shape, dtype=torch.uint8, pin_memory=True)


def run_b200(a):
    import torch
    import __graft_entry__ as g
    pkg = g.load_package()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    w, h, B, K, W = a.width, a.height, a.frames, a.steps, a.warmup

    # ---- CPU baseline on rank 0 at N=1, before the GPU is busy (bounded sample, separate process)
    cpu_baseline = None
    if world == 1 and not a.no_cpu_baseline:
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
                                  "--width", str(w), "--height", str(h), "--nfeatures", str(a.nfeatures)],
                                 capture_output=True, text=True, timeout=600)
            cpu_baseline = json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as e:  # the baseline is a reported number, never a reason to lose the GPU measurement
            cpu_baseline = {"value": None, "unit": "stereo pairs/s", "error": repr(e)[:200]}

    # ---- synthetic input pool: NB batches of B consecutive frames per rank (pinned host + device copies)
    NB = 3
    first_frame = rank * 4096
    host_l, host_r = pinned_u8((NB * B, h, w)), pinned_u8((NB * B, h, w))
    nl, nr = host_l.numpy(), host_r.numpy()
    for i in range(NB * B):
        rc = pkg.lib().svo_synth_stereo_frame(SEED, first_frame + i, w, h, nl[i].ctypes.data, nr[i].ctypes.data, w)
        assert rc == 0
    dev_l, dev_r = host_l.cuda(non_blocking=True), host_r.cuda(non_blocking=True)
    torch.cuda.synchronize()

    ctx = pkg.Context(local, w, h, 2 * (2 * B + 1))  # ring of 2 batches + the carried frame: downloads overlap the next run
    ctx.orb_configure(a.nfeatures, 1.2, 8, 20, 7)
    ctx.frontend_configure(w, h, pkg.MODE_ORB | pkg.MODE_LK, B)
    ctx.frontend_set_overlap(not a.no_overlap)
    stream = torch.cuda.Stream()  # a real (non-null) stream shared by torch events and the library's kernels
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    kp_cap, pt_cap = ctx.orb_capacity(w, h), max(1024, w * h // 64)
    plane = w * h

    def step_resident(k):
        b = k % NB
        ctx.frontend_upload_dev(dev_l.data_ptr() + b * B * plane, dev_r.data_ptr() + b * B * plane, B, w)
        ctx.frontend_run()
        ctx.frontend_download(None)  # results stay in HBM: retire the run (host bookkeeping only)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- value: device-timed, inputs resident in HBM
    for k in range(W):
        step_resident(k)
    barrier()
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
    stage = {s: 0.0 for s in pkg.STAGES}
    nslots = min(K, pkg.TIMING_SLOTS)
    for k in range(nslots):
        for s, v in ctx.stage_ms(k).items():
            stage[s] += max(v, 0.0) / nslots
    ctx.enable_stage_timing(False)
    clocks = sampler.finish()

    # ---- e2e: host buffers in, host buffers out, through svo_frontend_batch
    out = pkg.FrontendBuffers(B, kp_cap, pt_cap, alloc=lambda shape, dt: pinned_u8((int(np.prod(shape)) * np.dtype(dt).itemsize,)).numpy()
                              .view(dt).reshape(shape))
    ctx.frontend_reset()

    def batch_views(k):
        b = k % NB
        return nl[b * B:(b + 1) * B], nr[b * B:(b + 1) * B]

    def step_e2e(k):
        # software-pipelined use of the public API (pinned host memory): while batch k+1 runs on the GPU, batch k+2 is
        # uploaded (copy stream) and the results of batch k are downloaded (download stream).  Every batch is uploaded,
        # processed and downloaded inside the timed region.
        ctx.frontend_run()                      # batch k + 1
        ctx.frontend_upload(*batch_views(k + 2))
        ctx.frontend_download(out)              # batch k

    ctx.frontend_upload(*batch_views(0))
    ctx.frontend_run()
    ctx.frontend_upload(*batch_views(1))
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

    if world > 1:
        t = torch.tensor([ms_dev, ms_e2e], device="cuda", dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms_dev, ms_e2e = float(t[0]), float(t[1])
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak, peak_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        bpi = algorithmic_bytes_per_image(w, h)
        t_pf = stage["pyramid"] + stage["fast_cells"] + stage["compact"]
        achieved = bpi * 2 * B / (t_pf * 1e-3) / 1e9 if t_pf > 0 else None
        t_lk = sum(stage[f"lk_track{i}"] for i in (1, 2, 3, 4))
        tracks = 4 * int(np.sum(out.n_seed))  # points handed to each of the 4 calls in the last e2e step
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj.get("pyramid_fast_dram_bytes_per_step") * (2 * B) / tj.get("images_per_step", 128)  # measured per image, scaled to this launch
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": world * K * B / (ms_dev * 1e-3), "unit": "stereo pairs/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(a),
            "e2e": {"value": world * K * B / (ms_e2e * 1e-3), "unit": "stereo pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / K},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
                         "traffic": traffic, "peak_source": peak_src + " (of measured)" if "hbm_gbs" in peaks else peak_src,
                         "kernels": "K1 k_pyramid_fused (one-pass, TMA-staged) + K2 k_fast_cells + K2b k_compact (pyramid + FAST stages)",
                         "algorithmic_bytes_per_launch": bpi * 2 * B, "ms_per_launch": t_pf},
            "stage_ms_per_step": {k: round(v, 4) for k, v in stage.items()},
            "stage_note": ("event-to-event times per kernel group; pyramid / fast_cells / compact run before the stream fork and lk_track1 "
                           "(the whole 4-call chain) on the main stream; octree / describe / match (side stream) and fast_full / lk_pyramid "
                           "run concurrently, so their figures include waiting for SMs" if not a.no_overlap else "single stream"),
            # the kernel with the largest share of the step is the LK tracker; it is bound by instruction issue and the
            # shared-memory pipe, not by HBM or tensor throughput (profiles/*_ncu_full_all_kernels.txt), so it is reported as
            # tracks/s next to the HBM roofline of the pyramid + FAST stages that north_star asks for
            "dominant_kernel": {"name": "K8 k_lk_track (the 4-call chain, one launch per step)", "ms_per_step": round(t_lk, 4), "share_of_step": round(t_lk / (ms_dev / K), 3),
                                "tracks_per_step": tracks, "tracks_per_s": tracks / (t_lk * 1e-3) if t_lk > 0 else None,
                                "bound": "instruction issue + shared-memory pipe (integer; no HBM or tensor bound)"},
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        torch.distributed.destroy_process_group()


def run_reference(a):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import __graft_entry__ as g
    pkg = g.load_package()
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    g.load_oracle().lib()
    import ref_pipeline
    cores = len(os.sched_getaffinity(0))
    w, h = a.width, a.height
    units = a.sample_units or max(cores * 6, 12)  # bounded sample: a few units per core
    left, right = pkg.synth_sequence(SEED, 0, units + 1, w, h)
    pool = ref_pipeline.ReferencePool(left, right, 3, a.nfeatures, cores)
    for _ in range(a.warmup):
        pool.run()
    total = 0.0
    for _ in range(a.steps):
        dt, _cs = pool.run()
        total += dt
    pool.close()
    v = a.steps * units / total
    kind = ref_pipeline.orb_kind()
    orb_desc = ("the reference's own src/ORBextractor.cpp compiled into oracle/_ref (its OpenCV calls served by cv2-pinned restatements)"
                if kind == "reference" else "C++ oracle port of ORBextractor (AVX2 -O3)")
    sample = (f"{units} consecutive stereo frames of the same synthetic sequence per step, split over {pool.workers} worker processes "
              f"(1 OpenCV thread each); cv2 4.13 FAST/BFMatcher/calcOpticalFlowPyrLK + {orb_desc}")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "stereo pairs/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": total / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": workload_config(a),
            "cpu_baseline": {"value": v, "unit": "stereo pairs/s", "cores": pool.workers, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "stereo pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=50)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--frames", type=int, default=128, help="stereo frames per step per GPU")
    p.add_argument("--width", type=int, default=1241)
    p.add_argument("--height", type=int, default=376)
    p.add_argument("--nfeatures", type=int, default=1000)
    p.add_argument("--sample-units", type=int, default=0, help="reference arm: stereo frames per step (0 = 3 per core)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-overlap", action="store_true", help="run the ORB tail and the FAST + LK chain on one stream")
    a = p.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
