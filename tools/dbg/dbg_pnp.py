import sys, numpy as np, ctypes as C
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import __graft_entry__ as g
pkg = g.load_package(); orc = g.load_oracle()
gold = np.load("/root/repo/tests/golden/cv2_pnp.npz")
ctx = pkg.Context(0, 64, 64, 2)
K = gold["K"]
for k in (0, 7, 20):
    p3, p2 = gold[f"ransac_{k}_p3"], gold[f"ransac_{k}_p2"]
    # raw call to see the spare slot
    Kf = np.ascontiguousarray(K, np.float64).reshape(9)
    n = len(p3)
    cnt = np.array([n], np.int32); rvec = np.zeros(3); tvec = np.zeros(3); R = np.zeros(9); mask = np.zeros(n, np.uint8)
    ninl = np.zeros(1, np.int32); info = np.zeros(2, np.int32); ok = np.zeros(1, np.int32)
    r = ctx.pnp_ransac(K, p3, p2)
    o = orc.pnp_ransac(K, p3, p2)
    r2, t2, it = orc.pnp_refine(K, p3, p2, r["rvec"], r["tvec"], o["mask"].astype(np.uint8))
    print(k, "dev-oracle", np.abs(r["rvec"] - o["rvec"]).max(), np.abs(r["tvec"] - o["tvec"]).max(), "oracle-refine-from-dev iters", it, np.abs(t2 - o["tvec"]).max())
