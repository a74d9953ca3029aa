#!/usr/bin/env python
"""Frame source measurement: a synthetic KITTI-layout sequence on disk (image_0/%06d.png|pgm, image_1/...)
-> svo_source (decoder threads, pinned ring) -> svo_frontend_upload / run / download.  Prints one JSON line with the decode-only
rate and the end-to-end rate next to the in-memory e2e rate of the same batches.
usage: python tools/bench_source.py [--fmt png|pgm] [--frames 1024] [--batch 128] [--threads N] [--dir /tmp/svo_seq]"""
import argparse
import json
import os
import shutil
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--fmt", default="png", choices=["png", "pgm"])
    p.add_argument("--frames", type=int, default=1024)
    p.add_argument("--batch", type=int, default=128)
    p.add_argument("--threads", type=int, default=0)
    p.add_argument("--dir", default="/tmp/svo_seq")
    p.add_argument("--width", type=int, default=1241)
    p.add_argument("--height", type=int, default=376)
    a = p.parse_args()
    import __graft_entry__ as g
    import bench
    pkg = g.load_package()
    cores = len(os.sched_getaffinity(0))
    threads = a.threads or max(1, cores - 2)
    w, h, B, F = a.width, a.height, a.batch, a.frames
    left = np.empty((F, h, w), np.uint8)
    right = np.empty((F, h, w), np.uint8)
    bench.synth_into(pkg, 1000, 0, F, w, h, left, right, threads=min(cores, 16))
    shutil.rmtree(a.dir, ignore_errors=True)
    t0 = time.perf_counter()
    pkg.write_sequence(a.dir, left, right, fmt=a.fmt)
    t_write = time.perf_counter() - t0
    disk_bytes = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(a.dir) for f in fs)

    # decode only
    src = pkg.FrameSource(a.dir, w, h, B, ring_batches=3, threads=threads)
    assert src.frame_count() == F
    t0 = time.perf_counter()
    n = 0
    while True:
        l, r, k = src.next_ptr()
        if k == 0:
            break
        n += k
    t_dec = time.perf_counter() - t0
    pinned = src.is_pinned()
    src.close()

    # source -> frontend, results to host; checked against the in-memory path on the same frames
    ctx = pkg.Context(0, w, h, 2 * (2 * B + 1))
    ctx.orb_configure(1000, 1.2, 8, 20, 7)
    ctx.frontend_configure(w, h, pkg.MODE_ORB | pkg.MODE_LK, B)
    kp_cap, pt_cap = ctx.orb_capacity(w, h), max(1024, w * h // 64)
    out = pkg.FrontendBuffers(B, kp_cap, pt_cap)

    def run_from_source():
        ctx.frontend_reset()
        src = pkg.FrameSource(a.dir, w, h, B, ring_batches=3, threads=threads)
        sig = []
        t0 = time.perf_counter()
        l, r, k = src.next_ptr()
        ctx.frontend_upload_ptr(l, r, k, w)
        ctx.frontend_run()
        while True:
            l, r, k2 = src.next_ptr()  # the previous batch's ring slot is recycled here: its H2D copy must be complete
            if k2:
                ctx.frontend_upload_ptr(l, r, k2, w)
                ctx.frontend_run()
            ctx.frontend_download(out)
            sig.append((int(out.n_left[:k].sum()), int(out.n_lk_matched[:k].sum()), int(out.n_orb_matched[:k].sum())))
            if not k2:
                break
            k = k2
        dt = time.perf_counter() - t0
        src.close()
        return dt, sig

    def run_from_memory():
        ctx.frontend_reset()
        sig = []
        t0 = time.perf_counter()
        for f0 in range(0, F, B):
            k = min(B, F - f0)
            ctx.frontend_batch(left[f0:f0 + k], right[f0:f0 + k], out)
            sig.append((int(out.n_left[:k].sum()), int(out.n_lk_matched[:k].sum()), int(out.n_orb_matched[:k].sum())))
        return time.perf_counter() - t0, sig

    run_from_source()
    t_src, sig_src = run_from_source()
    _, sig_mem = run_from_memory()
    ok = sig_src == sig_mem
    ctx.close()
    print(json.dumps({"what": "frame source: KITTI-layout directory -> decoder threads -> pinned ring -> svo_frontend_upload/run/download",
                      "format": a.fmt, "frames": F, "width": w, "height": h, "batch": B, "decoder_threads": threads, "host_cores": cores,
                      "bytes_on_disk": disk_bytes, "write_s": round(t_write, 2), "ring_pinned": pinned,
                      "decode_only_pairs_per_s": n / t_dec, "source_to_results_pairs_per_s": F / t_src,
                      "results_equal_in_memory_path": ok}))
    shutil.rmtree(a.dir, ignore_errors=True)
    assert ok


if __name__ == "__main__":
    main()
