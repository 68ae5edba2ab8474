import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from oracle import orc
from svo_pro_universal_b200 import capi
ctx = capi.Context(0)
c = helpers.point_opt_cases()
for n_iter in (1, 2, 5):
    pos = c["pos0"].copy()
    iters = capi.optimize_points(ctx, pos, c["obs_begin"], c["obs_frame"], c["obs_f"], c["T_f_w"], n_iter, False)
    o, oi = [], []
    for i in range(len(pos)):
        lo, hi = c["obs_begin"][i], c["obs_begin"][i + 1]
        p, it = orc.point_optimize(c["T_f_w"][c["obs_frame"][lo:hi]], c["obs_f"][lo:hi], c["pos0"][i], n_iter, False)
        o.append(p); oi.append(it)
    o = np.array(o); oi = np.array(oi)
    d = np.abs(pos - o).max(1)
    bad = np.flatnonzero(d > 0)
    print("n_iter", n_iter, "mismatching points", len(bad), "max", d.max(), "iters differ", (iters != oi).sum())
    for b in bad[:6]:
        print("  point", b, "n_obs", c["obs_begin"][b + 1] - c["obs_begin"][b], "diff", d[b], "iters gpu/orc", iters[b], oi[b])
