"""Minimal driver for ncu captures of the detector kernels (f2 + a): pyramid, FAST, edgelet score / decode, FastGrad on 1184 frames."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svo_pro_universal_b200 import capi, synth  # noqa: E402

B = int(os.environ.get("PROF_FRAMES", 1184))
ctx = capi.Context(0)
pyr = capi.Pyramid(ctx, B, 752, 480, 5)
uimg = np.stack([synth.make_image(200 + s) for s in range(16)])
pyr.upload(uimg[np.arange(B) % 16])
pyr.build()
for _ in range(3):
    e = capi.edgelet_detect(ctx, pyr, 100, 8, 30)
for _ in range(3):
    f, e2 = capi.fastgrad_detect(ctx, pyr, capi.detector_options(), 100)
print("edgelets/frame", float((e["score"] > 100).sum()) / B, "after FAST", float((e2["score"] > 100).sum()) / B,
      "corners/frame", float((f["score"] > 10).sum()) / B)
