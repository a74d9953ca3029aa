#!/usr/bin/env python
"""profiles/traffic.json from an ncu --set full summary (tools/ncu_summary.py output): DRAM bytes per kernel per step, and the LK tracker's
warp-instructions per track.
usage: python tools/make_traffic.py profiles/r02f_ncu_full_all_kernels.txt [images_per_step] [tracks_per_step] [width] [height]"""
import collections
import json
import re
import subprocess
import sys


def main(path, images=256, tracks=None, width=1241, height=376):
    txt = open(path).read()
    per, inst = collections.OrderedDict(), {}
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}
    for blk in txt.split("\n== ")[1:]:
        name = blk.split()[0]
        if name == "void":
            name = blk.split()[1]
        name = name.split("<")[0]

        def val(k):
            m = re.search(re.escape(k) + r"\s+([\d.,]+)\s+(\w+)", blk)
            return float(m.group(1).replace(",", "")) * scale.get(m.group(2), 1)
        per[name] = per.get(name, 0) + val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        inst[name] = inst.get(name, 0) + val("smsp__inst_executed.sum")
    pf = per["k_pyramid_fused"] + per["k_fast_cells"] + per.get("k_compact", 0) + per.get("k_compact_scan", 0) + per.get("k_compact_gather", 0)
    try:
        commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    except Exception:
        commit = "?"
    out = {"pyramid_fast_dram_bytes_per_step": pf, "images_per_step": images, "width": width, "height": height,
           "source": f"{path} (ncu --set full --clock-control none, one step of {images // 2} stereo frames = {images} images; code state {commit})",
           "lk_warp_instr_per_track": (inst["k_lk_track"] / tracks) if tracks else None, "lk_tracks_per_step": tracks,
           "per_kernel_dram_bytes_per_step": per, "per_kernel_warp_instructions_per_step": inst}
    json.dump(out, open("profiles/traffic.json", "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    a = sys.argv
    main(a[1], int(a[2]) if len(a) > 2 else 256, int(a[3]) if len(a) > 3 else None, int(a[4]) if len(a) > 4 else 1241, int(a[5]) if len(a) > 5 else 376)
