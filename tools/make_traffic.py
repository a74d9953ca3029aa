#!/usr/bin/env python
"""profiles/traffic.json from an ncu --set full summary (tools/ncu_summary.py output): DRAM bytes per kernel per step.
usage: python tools/make_traffic.py profiles/r01d_ncu_full_all_kernels.txt [images_per_step]"""
import collections
import json
import re
import sys


def main(path, images=128):
    txt = open(path).read()
    per = collections.OrderedDict()
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}
    for blk in txt.split("\n== ")[1:]:
        name = blk.split()[0]
        if name == "void":
            name = blk.split()[1]
        name = name.split("<")[0]

        def val(k):
            m = re.search(re.escape(k) + r"\s+([\d.,]+)\s+(\w+)", blk)
            return float(m.group(1).replace(",", "")) * scale[m.group(2)]
        per[name] = per.get(name, 0) + val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    out = {"pyramid_fast_dram_bytes_per_step": per["k_pyramid_fused"] + per["k_fast_cells"] + per["k_compact"], "images_per_step": images,
           "source": f"{path} (ncu --set full, one step of {images // 2} stereo frames = {images} images)", "per_kernel_dram_bytes_per_step": per}
    json.dump(out, open("profiles/traffic.json", "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 128)
