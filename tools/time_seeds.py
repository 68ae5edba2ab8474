"""Times svo_cuda_update_seeds on 50k seeds x 64 ordered observations (CUDA events); SVO_CUDA_LIB selects the build under test."""
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
S, O = 50000, int(os.environ.get("SEED_OBS", "64"))
NSEQ_U, NOBS_U = 4, 16
seqs = [synth.make_seed_sequence(400 + s, n_seeds=400, n_obs=NOBS_U) for s in range(NSEQ_U)]
per = min(len(q["px"]) for q in seqs)
NSEQ = S // per
cam = capi.Camera.from_dict(seqs[0]["cam"])
ref = capi.Pyramid(ctx, NSEQ_U, 752, 480, 5); cur = capi.Pyramid(ctx, NSEQ_U * NOBS_U, 752, 480, 5)
ref.upload(np.stack([q["ref_img"] for q in seqs])); cur.upload(np.stack([im for q in seqs for im in q["cur_imgs"]]))
ref.build(); cur.build()
sid = np.arange(NSEQ) % NSEQ_U
catq = lambda k: np.concatenate([seqs[i][k][:per] for i in sid])
ftq = capi.make_features(catq("px"), catq("f"), catq("grad"), catq("type").astype(np.int32), catq("level"))
Sq = len(ftq)
ref_idx = np.repeat(sid, per).astype(np.int32)
obs = np.arange(O) % NOBS_U
obs_frame = (ref_idx[None, :] * NOBS_U + obs[:, None]).astype(np.int32)
Tq = np.concatenate([q["T_cur_ref"] for q in seqs])
types0 = t(catq("type").astype(np.uint8)); stq0 = t(catq("state"))
d_ftq = torch.from_numpy(ftq.view(np.uint8)).to(dev)
d_types, d_st = types0.clone(), stq0.clone()
d_mu2, d_ref_idx, d_obs, d_Tq = t(np.full(Sq, seqs[0]["mu_range"])), t(ref_idx), t(obs_frame), t(Tq)
mopt, dopt = capi.matcher_options(), capi.depth_filter_options()
def seeds():
    d_types.copy_(types0); d_st.copy_(stq0)
    return capi.update_seeds(ctx, ref, cur, cam, cam, d_ftq, d_types, d_st, d_mu2, d_obs, d_obs, d_Tq, mopt, dopt, ref_frame_idx=d_ref_idx, want_match_results=False)
seeds()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
torch.cuda.synchronize()
for a, b in ev:
    a.record(stream); n, _ = seeds(); b.record(stream)
torch.cuda.synchronize()
ms = float(np.median([a.elapsed_time(b) for a, b in ev]))
print(f"lib={os.path.basename(os.environ.get('SVO_CUDA_LIB', 'default'))} seeds={Sq} obs={O} ms={ms:.3f} obs/s={Sq*O/(ms*1e-3)/1e6:.1f}M success={float(n.item())/(Sq*O):.4f} state_sum={float(d_st.sum().item()):.9g}")
