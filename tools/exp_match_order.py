"""Experiment: does grouping the features of a matcher call by type (edgelets -> align1D, corners -> align2D) pay? Times
svo_cuda_find_match_direct on the same 512 k features in the given order and sorted by type (CUDA events, device arrays)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from svo_pro_universal_b200 import capi, synth

dev = torch.device("cuda", 0)
ctx = capi.Context(0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
NP, NF, NU = 256, 2000, 8
sets = [synth.make_match_set(300 + s, n_features=NF) for s in range(NU)]
ref = capi.Pyramid(ctx, NU, 752, 480, 5); cur = capi.Pyramid(ctx, NU, 752, 480, 5)
ref.upload(np.stack([m["ref_img"] for m in sets])); cur.upload(np.stack([m["cur_img"] for m in sets])); ref.build(); cur.build()
cam = capi.Camera.from_dict(sets[0]["cam"])
pid = np.arange(NP) % NU
cat = lambda k: np.concatenate([sets[i][k] for i in pid])
ft = capi.make_features(cat("px"), cat("f"), cat("grad"), cat("type"), cat("level"))
fidx = np.concatenate([np.full(len(sets[i]["px"]), i, np.int32) for i in pid])
T = np.stack([m["T_cur_ref"] for m in sets])
depth, guess = cat("depth"), cat("px_guess")
mopt = capi.matcher_options()
M = len(ft)
inv = 1.0 / depth; rng = np.random.default_rng(1); est = inv * rng.uniform(0.7, 1.4, M); spread = rng.uniform(0.1, 0.8, M) * inv
dinv = np.stack([est, est + spread, np.maximum(est - spread, 1e-8)], 1)

def run(order, name):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a[order])).to(dev)
    d_ft = torch.from_numpy(np.ascontiguousarray(ft[order]).view(np.uint8)).to(dev)
    d = dict(fidx=t(fidx), depth=t(depth), guess=t(guess), dinv=t(dinv))
    d_T = torch.from_numpy(T).to(dev)
    out = torch.zeros(M * capi.MATCH_OUT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    for which in ("direct", "epipolar"):
        def call():
            if which == "direct":
                capi.find_match_direct(ctx, ref, cur, cam, cam, d_T, d_ft, d["depth"], d["guess"], mopt, ref_frame_idx=d["fidx"], cur_frame_idx=d["fidx"], T_idx=d["fidx"], out=out)
            else:
                capi.find_epipolar_match_direct(ctx, ref, cur, cam, cam, d_T, d_ft, d["dinv"], mopt, ref_frame_idx=d["fidx"], cur_frame_idx=d["fidx"], T_idx=d["fidx"], out=out)
        for _ in range(3): call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10): call()
        e1.record(stream); torch.cuda.synchronize()
        r = out.cpu().numpy().view(capi.MATCH_OUT_DTYPE)
        print(f"{name:28s} {which:9s} {e0.elapsed_time(e1) / 10:.4f} ms  success {float((r['result'] == 0).mean()):.4f}")

typ = ft["type"]
run(np.arange(M), "given order")
run(np.argsort(typ, kind="stable"), "sorted by type")
run(np.lexsort((ft["level"], typ)), "sorted by type, level")
run(np.random.default_rng(3).permutation(M), "random permutation")
