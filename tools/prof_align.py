"""Tiny driver for ncu captures: one batch of B pairs through pyramid build + sparse alignment, a few launches."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
from svo_pro_universal_b200 import capi, synth, batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = capi.Context(0)
uniq = [synth.make_align_pair(5000 + s) for s in range(8)]
pk = batch.tile_batch(batch.pack_align_batch(uniq, max_features=180), B)
ref = capi.Pyramid(ctx, B, 752, 480, 5); cur = capi.Pyramid(ctx, B, 752, 480, 5)
ref.upload(pk["ref_imgs"]); cur.upload(pk["cur_imgs"]); ref.build()
cam = capi.Camera.from_dict(uniq[0]["cam"])
for _ in range(reps):
    cur.build()
    res = capi.sparse_align(ctx, [ref], [cur], [cam], pk["T_cam_imu"], pk["T_imu_world_ref"], pk["T_imu_world_cur"], pk["n_features"],
                            pk["px"], pk["f"], pk["depth"], pk["eligible"], capi.sparse_align_options())
print("iters", res["iters"][:4, :4].tolist(), "n", res["n_tracked"][:4])
