"""Quick CUDA-event timing of the detector path (config 2: 1024 frames) for iteration; the full per-path bench is bench_kernels.py."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svo_pro_universal_b200 import capi, synth  # noqa: E402

dev = torch.device("cuda:0"); stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
ctx = capi.Context(0); ctx.set_stream(stream.cuda_stream)
B = 1024
uniq = np.stack([synth.make_image(200 + s) for s in range(16)])
imgs = torch.from_numpy(uniq[np.arange(B) % 16]).to(dev)
pyr = capi.Pyramid(ctx, B, 752, 480, 5); pyr.upload(imgs); pyr.build()
opt = capi.detector_options()
corners = torch.zeros(B * 416 * 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    torch.cuda.synchronize()
    for a, b in ev:
        a.record(stream); fn(); b.record(stream)
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]))


print("pyramid ms", timed(lambda: pyr.build()))
print("detect ms", timed(lambda: capi.fast_detect(ctx, pyr, opt, corners_out=corners)))
print("pyr+detect ms", timed(lambda: capi.fast_detect(ctx, pyr, opt, corners_out=corners, fused_pyramid=True)))
