"""Experiment driver: sparse_align kernel time (CUDA events) for B pairs; env SVO_ALIGN_PAD_SMEM varies CTAs/SM."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from svo_pro_universal_b200 import capi, synth, batch
B = int(os.environ.get("PROF_PAIRS", "4096"))
dev = torch.device("cuda", 0)
ctx = capi.Context(0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
uniq = [synth.make_align_pair(5000 + s) for s in range(32)]
pk = batch.tile_batch(batch.pack_align_batch(uniq, max_features=180), B)
ref = capi.Pyramid(ctx, B, 752, 480, 5); cur = capi.Pyramid(ctx, B, 752, 480, 5)
ref.upload(torch.from_numpy(pk["ref_imgs"]).to(dev)); cur.upload(torch.from_numpy(pk["cur_imgs"]).to(dev)); ref.build(); cur.build()
cam = capi.Camera.from_dict(uniq[0]["cam"])
d = {k: torch.from_numpy(np.ascontiguousarray(pk[k])).to(dev) for k in ("T_imu_world_ref", "T_imu_world_cur", "n_features", "px", "f", "depth", "eligible")}
d_res = torch.zeros(B * capi.ALIGN_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
opt = capi.sparse_align_options(**({"estimate_illumination_gain": 1, "estimate_illumination_offset": 1} if os.environ.get("EXP_ILLUM") else {}),
                                **({"robustification": 1, "weight_scale": 10.0} if os.environ.get("EXP_ROBUST") else {}))
def run():
    capi.sparse_align(ctx, [ref], [cur], [cam], pk["T_cam_imu"], d["T_imu_world_ref"], d["T_imu_world_cur"], d["n_features"], d["px"], d["f"], d["depth"], d["eligible"], opt, results=d_res)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(10): run()
e1.record(stream); torch.cuda.synchronize()
res = d_res.cpu().numpy().view(capi.ALIGN_RESULT_DTYPE)
Hm = res["H"].reshape(-1, 8, 8)
tm = np.concatenate([res["H"][:, 48:64], Hm[:, :6, 6:8].reshape(-1, 12)], 1)
if tm[:, 0].max() > 0:  # profiling build (make -C svo_pro_universal_b200/csrc timing): per-phase clock64() counters
    names = ["total", "setup", "ref_patches", "residual_pass", "wait_barrier1", "h_rebuild", "serial_update", "wait_barrier2", "iters", "h_rebuilds"]
    m = tm.mean(0)
    print("phase cycles per pair (mean over pairs): " + ", ".join(f"{n}={v:.0f}" for n, v in zip(names, m)))
    it = max(m[8], 1.0)
    print(f"per iteration: residual {m[3]/it:.0f}, barrier1 {m[4]/it:.0f}, H {m[5]/it:.0f}, serial {m[6]/it:.0f}, barrier2 {m[7]/it:.0f}; per level: patches {m[2]/4:.0f}")
    print("residual pass per warp / iteration: " + " ".join(f"w{w}={m[10+w]/it:.0f}" for w in range(6)))
    print("serial phase / iteration: " + ", ".join(f"{n}={m[16+k]/it:.0f}" for k, n in enumerate(["totals", "dx", "broadcast", "update", "store+refresh"])))
print(f"lib={os.path.basename(os.environ.get('SVO_CUDA_LIB','libsvo_cuda.so'))} pad={os.environ.get('SVO_ALIGN_PAD_SMEM','0')} B={B} align_ms={e0.elapsed_time(e1)/10:.4f} iters_mean={res['iters'][:, :4].sum(1).mean():.2f} iters0={res['iters'][0][:4].tolist()}")
tag = os.path.basename(os.environ.get('SVO_CUDA_LIB', 'libsvo_cuda.so')).replace('.so', '')
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez(os.path.join(ROOT, "gpurun_out", f"align_res_{tag}_B{B}.npz"), T=res["T_icur_iref"], iters=res["iters"], n=res["n_tracked"], chi2=res["chi2"])
