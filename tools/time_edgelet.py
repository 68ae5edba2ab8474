"""A/B timer for the detector kernels (CUDA events, 1024 frames): SVO_CUDA_LIB selects the library build."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from svo_pro_universal_b200 import capi, synth  # noqa: E402

dev = torch.device("cuda:0")
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx = capi.Context(0)
ctx.set_stream(stream.cuda_stream)
B = 1024
uniq = np.stack([synth.make_image(200 + s) for s in range(16)])
pyr = capi.Pyramid(ctx, B, 752, 480, 5)
pyr.upload(torch.from_numpy(uniq[np.arange(B) % 16]).to(dev))
pyr.build()
corners = torch.zeros(B * 416 * 20, dtype=torch.uint8, device=dev)
edgelets = torch.zeros(B * 416 * 20, dtype=torch.uint8, device=dev)
opt = capi.detector_options()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    torch.cuda.synchronize()
    for a, b in ev:
        a.record(stream); fn(); b.record(stream)
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]))


print(os.environ.get("SVO_CUDA_LIB", "default"),
      "fast %.4f ms" % timed(lambda: capi.fast_detect(ctx, pyr, opt, corners_out=corners)),
      "edgelets %.4f ms" % timed(lambda: capi.edgelet_detect(ctx, pyr, 100, 8, 30, corners_out=edgelets)),
      "fastgrad %.4f ms" % timed(lambda: capi.fastgrad_detect(ctx, pyr, opt, 100, corners_out=corners, edgelets_out=edgelets)))
