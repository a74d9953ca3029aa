"""GPU debug helper: pyramid levels and FAST candidates of the CUDA path against the oracle, with mismatch details."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as g
pkg, orc = g.load_package(), g.load_oracle()
cfgs = [(640, 200, 500, 7), (1241, 376, 1000, 1000), (333, 257, 300, 11)]
if len(sys.argv) > 1:
    cfgs = [tuple(int(v) for v in sys.argv[1].split(","))]
for (w, h, nfeat, seed) in cfgs:
    left, right = pkg.synth_stereo_frame(seed, 0, w, h)
    ctx = pkg.Context(0, w, h, 2)
    ctx.orb_configure(nfeat, 1.2, 8, 20, 7)
    try:
        kps, desc = ctx.orb_extract_batch([left, right])
    except Exception as e:
        print("extract failed:", e)
    for slot, img in enumerate((left, right)):
        o = orc.OrbOracle(nfeat, 1.2, 8, 20, 7)
        ok, od = o.extract(img)
        for l in range(8):
            a, b = ctx.orb_read_level(slot, l), o.level(l)
            if not np.array_equal(a, b):
                ys, xs = np.nonzero(a != b)
                print(f"{w}x{h} slot {slot} level {l} ({b.shape[1]}x{b.shape[0]}): {len(ys)} pixels differ; x range {xs.min()}..{xs.max()} y range {ys.min()}..{ys.max()}; "
                      f"first (x={xs[0]}, y={ys[0]}) gpu {a[ys[0], xs[0]]} ref {b[ys[0], xs[0]]}; distinct x%4: {sorted(set((xs % 4).tolist()))}; cols sample {sorted(set(xs.tolist()))[:12]}")
            gc, c = ctx.orb_read_candidates(slot, l), o.candidates(l)
            sg = set(zip(gc["x"].tolist(), gc["y"].tolist(), gc["response"].tolist())); sc = set(zip(c["x"].tolist(), c["y"].tolist(), c["response"].tolist()))
            same_order = len(gc) == len(c) and all(np.array_equal(gc[f], c[f]) for f in ("x", "y", "response"))
            if not same_order:
                miss, extra = sorted(sc - sg), sorted(sg - sc)
                print(f"{w}x{h} slot {slot} level {l}: gpu {len(gc)} ref {len(c)} candidates; missing {len(miss)} extra {len(extra)}; missing[:6] {miss[:6]} extra[:6] {extra[:6]}")
        print(f"{w}x{h} slot {slot}: keypoints gpu {len(kps[slot]) if 'kps' in dir() else -1} ref {len(ok)}")
    ctx.close()
print("done")
