"""Experiment: one step (pyramid build + sparse align for B pairs) on one stream vs split over K contexts/streams."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from svo_pro_universal_b200 import capi, synth, batch
B = int(os.environ.get("PROF_PAIRS", "4096"))
dev = torch.device("cuda", 0)
uniq = [synth.make_align_pair(5000 + s) for s in range(32)]
cam = capi.Camera.from_dict(uniq[0]["cam"])
opt = capi.sparse_align_options()
main = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(main)

def setup(K):
    parts = []
    for k in range(K):
        b = B // K
        pk = batch.tile_batch(batch.pack_align_batch(uniq, max_features=180), b)
        ctx = capi.Context(0)
        st = torch.cuda.Stream(device=dev)
        ctx.set_stream(st.cuda_stream)
        ref = capi.Pyramid(ctx, b, 752, 480, 5); cur = capi.Pyramid(ctx, b, 752, 480, 5)
        ref.upload(torch.from_numpy(pk["ref_imgs"]).to(dev)); cur.upload(torch.from_numpy(pk["cur_imgs"]).to(dev)); ref.build()
        d = {k2: torch.from_numpy(np.ascontiguousarray(pk[k2])).to(dev) for k2 in ("T_imu_world_ref", "T_imu_world_cur", "n_features", "px", "f", "depth", "eligible")}
        res = torch.zeros(b * capi.ALIGN_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        parts.append((ctx, st, ref, cur, pk, d, res))
    torch.cuda.synchronize()
    return parts

def step(parts):
    e0 = torch.cuda.Event(); e0.record(main)
    for ctx, st, ref, cur, pk, d, res in parts:
        st.wait_event(e0)
        cur.build()
        capi.sparse_align(ctx, [ref], [cur], [cam], pk["T_cam_imu"], d["T_imu_world_ref"], d["T_imu_world_cur"], d["n_features"], d["px"], d["f"], d["depth"], d["eligible"], opt, results=res)
        e = torch.cuda.Event(); e.record(st); main.wait_event(e)

for K in (1, 2, 4, 8):
    parts = setup(K)
    for _ in range(3): step(parts)
    torch.cuda.synchronize()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(main)
    for _ in range(10): step(parts)
    b_.record(main); torch.cuda.synchronize()
    print(f"K={K} step_ms={a.elapsed_time(b_)/10:.4f} pairs/s={B/(a.elapsed_time(b_)/10*1e-3):.0f}")
    del parts
