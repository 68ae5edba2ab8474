"""A/B timer for the keyframe stage (chain with stereo triangulation, 2048 pairs) and bench path f3-like dense lists; SVO_CUDA_LIB selects the build."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from svo_pro_universal_b200 import capi, frontend  # noqa: E402

dev = torch.device("cuda:0")
ctx = capi.Context(0)
scenes = [frontend.make_stereo_scene(81 + s) for s in range(2)]
fb = frontend.StereoFrontendBatch(ctx, scenes, 2048, dev, stereo_triangulation=True)
for _ in range(2):
    fb.step()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
ts = []
for _ in range(5):
    ev[0].record(fb.stream)
    fb.step(lambda i: ev[i + 1].record(fb.stream))
    torch.cuda.synchronize()
    ts.append(ev[6].elapsed_time(ev[7]))
print(os.environ.get("SVO_CUDA_LIB", "default"), "stereo stage %.3f ms per 2048 pairs" % float(np.median(ts)),
      "triangulated", float(fb.results()["stereo_stats"]["n_succeeded"].mean()))
