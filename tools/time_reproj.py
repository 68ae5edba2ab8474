"""Times svo_cuda_reproject_match on 1024 current frames sharing one map (CUDA events); SVO_CUDA_LIB selects the build under test."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from svo_pro_universal_b200 import capi, synth
dev = torch.device("cuda:0")
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
ctx = capi.Context(0); ctx.set_stream(stream.cuda_stream)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
sc = synth.make_reproject_scene(21, n_cur=8)
K, F = len(sc["kf_imgs"]), int(os.environ.get("REPROJ_FRAMES", "1024"))
cam = capi.Camera.from_dict(sc["cam"])
ref = capi.Pyramid(ctx, K, 752, 480, 5); cur = capi.Pyramid(ctx, 8, 752, 480, 5)
ref.upload(np.stack(sc["kf_imgs"])); cur.upload(np.stack(sc["cur_imgs"])); ref.build(); cur.build()
tb = dict(sc["tables"])
tb["feat"] = capi.make_features(tb["feat"]["px"], tb["feat"]["f"], tb["feat"]["grad"], tb["feat"]["type"], tb["feat"]["level"])
d_tb = {k: (t(v.view(np.uint8) if v.dtype.fields else v) if isinstance(v, np.ndarray) else v) for k, v in tb.items()}
ef = np.ascontiguousarray(sc["entry_feat"], np.int32); E1 = len(ef)
d_idx = t((np.arange(F) % 8).astype(np.int32)); d_T = t(np.ascontiguousarray(sc["cur_Ts"][np.arange(F) % 8], np.float64))
d_nin, d_eb, d_ef = t(np.zeros(F, np.int32)), t((np.arange(F + 1) * E1).astype(np.int32)), t(np.tile(ef, F))
occ0 = torch.zeros((F, 416), dtype=torch.uint8, device=dev); d_occ = occ0.clone()
d_res = torch.zeros(F * E1 * capi.REPROJ_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
d_st = torch.zeros(F * capi.REPROJ_STATS_DTYPE.itemsize, dtype=torch.uint8, device=dev)
for max_n in (120, 1000):
    ropt = capi.reprojector_options(max_n_features=max_n)
    def run():
        d_occ.copy_(occ0)
        capi.reproject_match(ctx, ref, cur, cam, cam, d_tb, d_T, d_nin, d_eb, d_ef, d_occ, ropt, cur_frame_idx=d_idx, results=d_res, stats=d_st)
    for _ in range(3): run()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    torch.cuda.synchronize()
    for a, b in ev:
        a.record(stream); run(); b.record(stream)
    torch.cuda.synchronize()
    ms = float(np.median([a.elapsed_time(b) for a, b in ev]))
    st = d_st.cpu().numpy().view(capi.REPROJ_STATS_DTYPE)
    print(f"lib={os.path.basename(os.environ.get('SVO_CUDA_LIB', 'default'))} max_n={max_n} F={F} ms={ms:.3f} frames/s={F/(ms*1e-3):.0f} trials={st['n_trials'].mean():.1f} matches={st['n_matches'].mean():.1f}")
