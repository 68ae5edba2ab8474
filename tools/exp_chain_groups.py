"""Experiment: the configs[4] front-end chain cut into G groups of stereo pairs, each a StereoFrontendBatch on its own context and
stream, stepped back to back from one host thread (every call is asynchronous, so the groups' kernels overlap on the GPU).
    python tools/exp_chain_groups.py [pairs] [groups ...]
Wall clock between two device synchronisations around `steps` passes over all groups (the region is >= 100 ms); an experiment
driver, not a bench value."""
import gc
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from svo_pro_universal_b200 import capi, frontend, shard  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
groups = [int(a) for a in sys.argv[2:]] or [1, 2, 4]
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
scenes = [frontend.make_stereo_scene(81 + s) for s in range(4)]
steps, warmup = 6, 2
for G in groups:
    ctxs = [capi.Context(0) for _ in range(G)]
    fbs = []
    for g, ctx in enumerate(ctxs):
        lo, hi = shard.partition(pairs, G, g)
        fbs.append(frontend.StereoFrontendBatch(ctx, scenes, hi - lo, dev))
    for _ in range(warmup):
        for fb in fbs:
            fb.step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        for fb in fbs:
            fb.step()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / steps * 1e3
    ok = all((fb.results()["align"]["n_tracked"] > 250).all() for fb in fbs)
    poses = np.concatenate([fb.results()["align"]["T_icur_iref"] for fb in fbs])
    print(f"groups={G} pairs={pairs} ms_per_step={ms:.3f} pairs/s={pairs / (ms * 1e-3):.0f} host_issue_ms_per_step={t_issue / steps * 1e3:.2f} "
          f"tracked_ok={ok} pose_sum={float(np.abs(poses).sum()):.9f}", flush=True)
    if G != groups[-1]:  # (one group count per process is the safe way to run this; several only if memory allows)
        for fb in fbs:
            fb.release()
        del fbs, ctxs
        gc.collect()
        torch.cuda.empty_cache()
