"""BASELINE.json configs[4]: the full front-end batch (pyramid + 2-camera sparse align + Reprojector feature alignment + depth
filter + FastGrad detector) on --pairs synthetic stereo frame pairs, sharded over the ranks of a torchrun launch by contiguous
blocks of pairs (strong scaling: the total is fixed), no collective inside the path; one JSON line on rank 0.
    python tools/bench_frontend.py --pairs 8192
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_frontend.py --pairs 8192
Timing: CUDA events on the launching stream around --steps passes, max over ranks. The CPU line is the oracle port of the
same chain, single thread, on the unique scenes."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def cpu_chain(scenes, seconds=6.0):
    """Oracle port of the same per-pair chain (single thread): stereo pairs per second (the chain itself lives in bench.py)."""
    import bench
    prep = bench.cpu_chain_prepare(scenes)
    t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < seconds:
        n += bench.cpu_chain_once(prep[n % len(prep)])
    return n / (time.perf_counter() - t0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=8192, help="stereo frame pairs in total (sharded over the ranks)")
    ap.add_argument("--unique", type=int, default=4)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--stereo", action="store_true", help="add the keyframe stage: StereoTriangulation of the features just detected")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from svo_pro_universal_b200 import capi, frontend, shard
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ctx = capi.Context(local)
    lo, hi = shard.partition(args.pairs, world, rank)
    scenes = [frontend.make_stereo_scene(81 + s) for s in range(args.unique)]
    fb = frontend.StereoFrontendBatch(ctx, scenes, hi - lo, dev, stereo_triangulation=args.stereo)
    stage_names = fb.STAGES + ((fb.STEREO_STAGE,) if args.stereo else ())
    stream = fb.stream
    torch.cuda.set_stream(stream)
    for _ in range(args.warmup):
        fb.step()
    sync = lambda: (torch.cuda.synchronize(), dist.barrier() if world > 1 else None, torch.cuda.synchronize())
    sync()
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        fb.step()
    e1.record(stream)
    sync()
    ms = e0.elapsed_time(e1) / args.steps
    # stage breakdown of one more pass (events between the stages)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(stage_names) + 1)]
    evs[0].record(stream)
    fb.step(lambda i: evs[i + 1].record(stream))
    torch.cuda.synchronize()
    stages = {n: evs[i].elapsed_time(evs[i + 1]) for i, n in enumerate(stage_names)}
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms = float(tt.item())
    out = fb.results()
    assert (out["align"]["n_tracked"] > 250).all() and out["reproj_stats"]["n_matches"].mean() > 60
    if world > 1:  # the only collective: gather the per-pair poses to rank 0
        g = shard.gather_to_rank0(out["align"]["T_icur_iref"].copy(), args.pairs)
        assert rank != 0 or g.shape == (args.pairs, 7)
    if rank == 0:
        cpu = cpu_chain(scenes)
        print(json.dumps({"path": "configs[4]: full front-end batch (pyramid + stereo sparse align + Reprojector + pose optimizer + depth filter + FastGrad detector)",
                          "config": f"{args.pairs} synthetic stereo frame pairs ({args.unique} unique scenes tiled; every frame resident in HBM: "
                                    f"{4 * (hi - lo) * 483360 / 1e9:.1f} GB of pyramids per GPU), 180 + 150 features, 120 seeds per pair",
                          "n_gpus": world, "stereo_pairs_per_s": args.pairs / (ms * 1e-3), "ms_per_step": ms, "scaling": "strong",
                          "stage_ms_rank0": stages,
                          "gpu_launches_per_step": (ctx.launches - l0) // args.steps,
                          "mean_matches_per_frame": float(out["reproj_stats"]["n_matches"].mean()),
                          "mean_triangulated_per_pair": float(out["stereo_stats"]["n_succeeded"].mean()) if args.stereo else None,
                          "seed_success_frac": out["n_seed_ok"] / max(1, fb.S),
                          "cpu_baseline": {"stereo_pairs_per_s_single_thread": cpu, "kind": "port (oracle chain)", "cores": 1}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
