#!/usr/bin/env python
"""Headline benchmark: aligned frame-pairs/s (752x480, 4-level SparseImgAlign, ~180 features) and p50 single-pair latency,
plus one measured line per BASELINE.json config (`paths`), each checked against the CPU oracle on a random sample of the
full-size batch after its timed region.

    python bench.py --gpus N --steps K --warmup W                   # our CUDA path (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's own CPU path on the host cores

Headline: one STEP = one pass of the hot path over one batch of B synthetic frame pairs per GPU:
    build the pyramid of every NEW (cur) frame from its level-0 image  ->  SparseImgAlign::run for every pair
(the ref frame's pyramid already exists: it was the previous step's new frame — src/svo/src/frame_handler_base.cpp:184-186
builds it at frame creation, :634 aligns against the last frame). The initial pose guess of every pair carries its own small
random error, so the Gauss-Newton iteration counts differ from pair to pair as they do in a real batch.
`value`  : whole-job pairs/s with every input resident in HBM (CUDA events on the launching stream, max over ranks).
`e2e`    : the same metric through the C ABI with HOST buffers: per step the new frames' level-0 images and the feature
           arrays go host(pinned)->device and the per-pair results come back, all inside the timed region.
`paths`  : BASELINE.json configs[1..4] (FAST detection of 1024 frames, matcher on 512 k features, 50 k seeds x 64
           observations, the 8192-pair stereo front-end chain), each with value / kernel_ms / roofline / e2e / cpu_baseline
           (N = 1 only) / parity_sampled.
Only the cpu_baseline legs, the reference arm and the sampled parity checks (outside every timed region) execute oracle/.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, N_LEVELS, N_FEATURES = 752, 480, 5, 180
# SURVEY.md §8d algorithmic bytes per unit
ALGO_BYTES_PER_PAIR = 247044       # (b) ref L1-L4 + cur L1-L4 (2 x 119,850 B) + 180 x 40 B features + 2 x 72 B state
ALGO_BYTES_PER_FRAME = 487466      # (a) read L0 + write L1-L4 + 416 corners out, pyramid and detection fused
ALGO_BYTES_FAST_ONLY = 360960 + 92160 + 23040 + 416 * 20   # detection alone: read the (pitched) levels 0-2 once + corners
ALGO_BYTES_PER_FEATURE = 333       # (c) patch with border + cur footprint + feature in + result out
ALGO_BYTES_PER_UPDATE = 80         # (d) pure filter update, state in + out
METRIC = "aligned frame-pairs/sec (752x480, 4-level SparseImgAlign, ~180 features)"


def ncu_traffic(kernel, units):
    """`roofline.traffic`: dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of `kernel` from this round's `ncu --set full`
    capture (profiles/ncu_traffic.json, written by tools/ncu_traffic.py), scaled from the captured launch's units to `units`; None if
    the kernel was not captured."""
    try:
        k = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["kernels"][kernel]
    except (OSError, KeyError, ValueError):
        return None
    return k["dram_bytes_per_unit"] * units


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        return {}


def ncu_pipes(kernel):
    """Pipe / issue utilisation of `kernel` in this round's `ncu --set full` capture (profiles/ncu_traffic.json): what the profiler says
    bounds a kernel that is not memory bound. Shares measured under the profiler on the capture's workload, not live values."""
    try:
        j = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return dict(j["kernels"][kernel]["pipes"], capture=j["capture"] + "/" + j["kernels"][kernel]["file"])
    except (OSError, KeyError, ValueError):
        return None


def roofline(kernel, bytes_per_unit, units, kernel_ms, peaks, note, traffic=None, ncu=None):
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = bytes_per_unit * units / (kernel_ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "ncu": ncu,
            "kernel": kernel, "kernel_ms": kernel_ms, "algorithmic_bytes_per_unit": bytes_per_unit, "units_per_launch": units,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)", "note": note}


def make_unique_pairs(n_unique, seed0):
    from svo_pro_universal_b200 import synth
    return [synth.make_align_pair(seed0 + s, n_features=N_FEATURES) for s in range(n_unique)]


def perturbed_initial_poses(T_imu_world_ref, seed):
    """Initial guess of the new frame's pose per pair: the ref pose (identity motion, frame_handler_base.cpp:346-358 without a
    motion prior) composed with a small random error (sigma 0.15 deg, 3 mm), different for every pair of the batch."""
    from svo_pro_universal_b200 import synth
    rng = np.random.default_rng(seed)
    out = np.empty_like(T_imu_world_ref)
    for i, T in enumerate(T_imu_world_ref):
        out[i] = synth.se3_mul(synth.se3_exp_small(rng.normal(size=3) * np.deg2rad(0.15), rng.normal(size=3) * 0.003), T)
    return out


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions. They last a few tens of milliseconds, far below
    nvidia-smi's 200 ms period, so the samples come from NVML (the library nvidia-smi itself reads) polled in-process every
    ~1 ms; the B200_PROFILING.md nvidia-smi line is the fallback."""

    _REASONS = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
                ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"))

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.max_mhz = index, [], set(), None
        self.proc, self.rows, self._stop, self._thread, self.source = None, [], threading.Event(), None, None
        self._on = threading.Event()

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def start(self):
        self._on.set()
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            masks = [(n, getattr(pynvml, a)) for n, a in self._REASONS]

            def poll():
                while not self._stop.is_set():
                    if self._on.is_set():
                        try:
                            self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                            r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                            for n, m in masks:
                                if r & m:
                                    self.reasons.add(n)
                        except pynvml.NVMLError:
                            pass
                    time.sleep(0.001)

            self._thread = threading.Thread(target=poll, daemon=True)
            self._thread.start()
            self.source = "nvml, ~1 ms period, sampled during the timed regions of every leg"
            return
        except Exception:  # noqa: BLE001 - any NVML problem falls back to the nvidia-smi recipe
            self._thread = None
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self._physical_index()}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            self.source = "nvidia-smi -lms 200"
        except OSError:
            self.proc = None

    def pause(self):   # CPU-only stretches (data generation, oracle checks) are not GPU load: do not sample them
        self._on.clear()

    def resume(self):
        self._on.set()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=1.0)
        if self.proc:
            self.proc.terminate()
            mx = []
            for r in self.rows:
                try:
                    self.sm.append(float(r[1])); mx.append(float(r[2]))
                except (ValueError, IndexError):
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            if mx:
                self.max_mhz = float(max(mx))
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples (NVML and nvidia-smi unavailable)"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": self.source}


# --------------------------------------------------------------------------------------------------------------------
# CPU legs (the reference's own compiled sources where they travelled to this box, else the oracle port)

def threaded_throughput(fn, n_threads, seconds):
    """Units per second of fn(worker_index, call_index) -> units, called back to back from n_threads Python threads (the ctypes
    calls release the GIL, so the C code runs in parallel) for about `seconds`."""
    from concurrent.futures import ThreadPoolExecutor
    stop = time.perf_counter() + seconds

    def worker(w):
        n, k = 0, 0
        while time.perf_counter() < stop:
            n += fn(w, k); k += 1
        return n

    t0 = time.perf_counter()
    with ThreadPoolExecutor(n_threads) as ex:
        total = sum(ex.map(worker, range(n_threads)))
    return total / (time.perf_counter() - t0)


def cpu_chain_prepare(scenes):
    """Per unique stereo scene: the oracle pyramids of the two reference frames (they exist before the step, as on the GPU)."""
    from oracle import orc
    return [(sc, {c: orc.create_img_pyramid(sc["imgs"][f"r{c}"], 5) for c in range(2)}) for sc in scenes]


def cpu_chain_once(prepared_scene):
    """The front-end chain of ONE stereo pair with the oracle port (cpu_baseline of BASELINE configs[4]): pyramids of the two new
    frames, 2-camera SparseImgAlign, Reprojector per camera, DepthFilter::updateSeeds, FastGrad detector. Returns 1 (a pair)."""
    from oracle import orc
    from svo_pro_universal_b200 import synth
    sc, pyr_r = prepared_scene
    cam, keep = sc["cam"], []
    ident = np.array([1.0, 0, 0, 0, 0, 0, 0])
    ang = float(np.arctan(1 / (2 * cam["fx"])) + np.arctan(1 / (2 * cam["fy"])))
    pyr_c = {c: orc.create_img_pyramid(sc["imgs"][f"c{c}"], 5) for c in range(2)}
    rfs = [orc.make_frame(pyr_r[c], cam, sc["T_cam_imu"][c], sc["T_imu_world_ref"], sc["px"][c], sc["f"][c], sc["depth"][c], keep=keep) for c in range(2)]
    cfs = [orc.make_frame(pyr_c[c], cam, sc["T_cam_imu"][c], sc["T_imu_world_ref"], keep=keep) for c in range(2)]
    r = orc.sparse_align(rfs, cfs, orc.default_align_options(estimate_illumination_gain=1, estimate_illumination_offset=1))
    for c in range(2):
        m = len(sc["px"][c])
        kf = orc.make_frame(pyr_r[c], cam, ident, sc["T_f_w_ref"][c], keep=keep)
        cf = orc.make_frame(pyr_c[c], cam, ident, np.array(r.T_f_w[c][:]), keep=keep)
        st = np.tile([1.0, 1e-6, 10.0, 10.0], (m, 1)); st[:, 0] = 1.0 / sc["depth"][c]
        R, tt = synth.se3_to_Rt(synth.se3_inv(sc["T_f_w_ref"][0]))
        feat = np.zeros(m, [("px", "<f8", 2), ("f", "<f8", 3), ("grad", "<f8", 2), ("type", "<i4"), ("level", "<i4")])
        feat["px"], feat["f"], feat["grad"], feat["type"] = sc["px"][c], sc["f"][c], [1.0, 0.0], 7 if c == 0 else 4
        tb = dict(n_kfs=1, n_feat=m, n_points=m if c == 0 else 0, kf_seed_mu_range=np.array([1 / 1.5]), kf_feat_begin=np.array([0, m], np.int32),
                  feat=feat, feat_score=np.linspace(60.0, 11.0, m), feat_seed_state=st,
                  feat_point=(np.arange(m) if c == 0 else np.full(m, -1)).astype(np.int32), feat_kf=np.zeros(m, np.int32),
                  pt_pos=((sc["f"][0] * sc["depth"][0][:, None]) @ R.T + tt) if c == 0 else np.zeros((1, 3)),
                  pt_n_failed=np.zeros(max(m, 1), np.int32), pt_n_succeeded=np.zeros(max(m, 1), np.int32),
                  pt_obs_begin=(np.arange(m + 1) if c == 0 else np.zeros(1)).astype(np.int32),
                  obs_feat=(np.arange(m) if c == 0 else np.zeros(1)).astype(np.int32))
        orc.reproject_match([kf], tb, cf, np.arange(m, dtype=np.int32), 0, np.zeros(416, np.uint8), orc.ReprojOptions(30, 120, 1, 0, 0, 200.0, ang))
    ns = len(sc["seed_px"])
    oft = orc.make_features(sc["seed_px"], sc["seed_f"], np.tile([1.0, 0.0], (ns, 1)), np.full(ns, 1, np.int32), np.zeros(ns, np.int32))
    orc.update_seeds(rfs[0], [cfs[0]], sc["T_cur_ref_gt"].reshape(1, 7), oft, np.full(ns, 1, np.uint8), sc["seed_state"].copy(),
                     sc["seed_mu_range"], orc.default_matcher_options())
    orc.detect_features(orc.DETECTOR_FAST_GRAD, pyr_c[0])
    return 1


def orc_frames(orc, pairs, keep):
    """Oracle ref frames (prebuilt pyramid + features) and level-0 images of a list of synthetic pairs."""
    refs, l0 = [], []
    for d in pairs:
        rp = orc.create_img_pyramid(d["ref_img"], N_LEVELS)
        refs.append(orc.make_frame(rp, d["cam"], d["T_cam_imu"], d["T_imu_world_ref"], d["px"], d["f"], d["depth"], d["eligible"], keep=keep))
        l0.append(d["cur_img"])
    return refs, l0


def cpu_step_fn(n_pairs_per_step, n_threads, seed0=1000, n_unique=16, kind="auto"):
    """Returns (fn, sample description, kind); fn() runs one bounded CPU step (pyramid + SparseImgAlign::run per pair, every pair
    with its own perturbed initial pose) on all threads. kind "reference" = the reference's own sources compiled into
    oracle/_ref/libfrontend_ref.so (vk::halfSample + svo::SparseImgAlign::run; Eigen/OpenCV/glog resolved to the stand-ins of
    oracle/shim), "port" = the oracle restatement; "auto" picks the compiled reference when it travelled to this box."""
    from oracle import orc
    uniq = make_unique_pairs(n_unique, seed0)
    keep = []
    refs, l0 = orc_frames(orc, uniq, keep)
    idx = [i % n_unique for i in range(n_pairs_per_step)]
    T0 = perturbed_initial_poses(np.stack([uniq[i]["T_imu_world_ref"] for i in idx]), seed0 + 7)
    R, L = [refs[i] for i in idx], [l0[i] for i in idx]
    Cc = [orc.make_frame([uniq[i]["cur_img"]], uniq[i]["cam"], uniq[i]["T_cam_imu"], T0[k], keep=keep) for k, i in enumerate(idx)]
    opt = orc.default_align_options()
    if kind == "auto":
        kind = "reference" if orc.ref_frontend_lib() is not None else "port"

    if kind == "reference":
        def fn():
            return orc.ref_pyramid_align_batch(L, R, Cc, opt, N_LEVELS, n_threads)
        what = ("the reference's own SparseImgAlign + halfSample sources compiled -O3 against a scalar (non-vectorised) Eigen stand-in "
                "(oracle/_ref/libfrontend_ref.so)")
    else:
        def fn():
            return orc.pyramid_align_batch(L, R, Cc, opt, N_LEVELS, n_threads)
        what = "oracle port of the reference CPU path (-O3)"

    fn._keep = (keep, uniq)
    return fn, f"{n_pairs_per_step} pairs/step ({n_unique} unique synthetic pairs tiled, per-pair perturbed initial pose), {what}, {n_threads} threads", kind


def cpu_baseline(n_threads, seconds):
    """Throughput (all host threads) and single-thread p50 latency of the CPU path over about `seconds` of work per build. Both
    builds are timed — the reference's own compiled sources and the oracle port — and the FASTER one is the baseline (`value`),
    so that a speed-up quoted against it is not inflated by the scalar Eigen stand-in of the compiled reference."""
    from oracle import orc
    per_step = max(64, 16 * n_threads)
    out = {}
    for kind in ("reference", "port"):
        if kind == "reference" and orc.ref_frontend_lib() is None:
            continue
        fn, sample, _ = cpu_step_fn(per_step, n_threads, kind=kind)
        fn()
        t0 = time.perf_counter(); reps = 0
        while time.perf_counter() - t0 < seconds:
            fn(); reps += 1
        value = per_step * reps / (time.perf_counter() - t0)
        lat_fn, _, _ = cpu_step_fn(1, 1, kind=kind)
        cl = []
        for _ in range(40):
            t = time.perf_counter(); lat_fn(); cl.append(time.perf_counter() - t)
        out[kind] = {"value": value, "sample": sample + ", ~%d s" % seconds, "latency_ms_p50_single_thread": 1e3 * float(np.median(cl))}
    best = max(out, key=lambda k: out[k]["value"])
    cb = {"value": out[best]["value"], "unit": "pairs/s", "cores": n_threads, "kind": best, "sample": out[best]["sample"],
          "latency_ms_p50_single_thread": min(o["latency_ms_p50_single_thread"] for o in out.values()),
          "note": "value = the faster of the two CPU builds below (speed-ups are quoted against the faster one)",
          "builds": {k: {"value": v["value"], "latency_ms_p50_single_thread": v["latency_ms_p50_single_thread"], "sample": v["sample"]}
                     for k, v in out.items()}}
    return cb


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # other ranks exit 0 without work
    n_threads = os.cpu_count() or 1
    per_step = max(64, 16 * n_threads)
    fn, sample, kind = cpu_step_fn(per_step, n_threads)
    for _ in range(max(1, args.warmup)):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    lat_fn, _, _ = cpu_step_fn(1, 1, kind=kind)
    lats = []
    for _ in range(50):
        t = time.perf_counter(); lat_fn(); lats.append(time.perf_counter() - t)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "SparseImgAlign batch: pyramid of the new frame + run() per pair, 752x480, levels 4->1, 180 features, 4x4 patches "
                               "(BASELINE configs[0] batched)",
                   "pairs_per_step": per_step, "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": n_threads, "kind": kind, "sample": sample,
                         "latency_ms_p50_single_thread": 1e3 * float(np.median(lats))},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


# --------------------------------------------------------------------------------------------------------------------
# GPU legs

class Rig:
    """What every GPU leg shares: device, context, stream, rank bookkeeping and the timing helpers (CUDA events on the launching
    stream, a barrier + synchronize on both sides, max over ranks)."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from svo_pro_universal_b200 import capi
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.ctx = capi.Context(self.local_rank)
        self.stream = torch.cuda.Stream(device=self.dev)  # a real (non-legacy) stream: the C ABI treats a NULL handle as "use your own"
        torch.cuda.set_stream(self.stream)
        self.ctx.set_stream(self.stream.cuda_stream)      # our kernels and torch's events share one stream
        self.peaks = load_peaks()
        self.sampler = ClockSampler(self.local_rank)
        self.flush_buf = None

    def t(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)

    def pin(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).pin_memory()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def flush_l2(self):
        """Write a buffer larger than the 126 MB L2 (not inside any timed event pair)."""
        if self.flush_buf is None:
            self.flush_buf = self.torch.empty(256 << 20, dtype=self.torch.uint8, device=self.dev)
        self.flush_buf.fill_(1)

    def timed(self, fn, steps, warmup, flush=False):
        """ms per step of fn(), max over ranks. flush=False: one event pair around `steps` back-to-back calls (the inputs are larger
        than L2). flush=True: one event pair per step with an L2 flush between the steps, the per-step times are averaged."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        if not flush:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            for _ in range(steps):
                fn()
            e1.record(self.stream)
            self.barrier()
            ms = e0.elapsed_time(e1) / steps
        else:
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            for a, b in ev:
                self.flush_l2()
                a.record(self.stream); fn(); b.record(self.stream)
            self.barrier()
            ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
        return self.max_over_ranks(ms)

    def timed_e2e(self, fn, steps, warmup):
        """ms per step of a host-buffer call sequence: the larger of the device time (events) and the host wall clock, max over ranks."""
        torch = self.torch
        for _ in range(max(1, min(warmup, 3))):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(self.stream)
        for _ in range(steps):
            fn()
        e1.record(self.stream)
        self.barrier()
        wall = (time.perf_counter() - t0) * 1e3
        return self.max_over_ranks(max(e0.elapsed_time(e1), wall)) / steps


def pose_diff(Ta, Tb):
    dt = float(np.abs(np.asarray(Ta)[4:] - np.asarray(Tb)[4:]).max())
    dq = 2 * float(np.arccos(min(1.0, abs(float(np.dot(np.asarray(Ta)[:4], np.asarray(Tb)[:4]))))))
    return dq, dt


def leg_headline(rig, sample_n=64):
    """BASELINE configs[0] batched: pyramid of the new frame + SparseImgAlign::run, B pairs per GPU."""
    import torch
    from svo_pro_universal_b200 import capi, batch, shard
    args, ctx, dev, stream = rig.args, rig.ctx, rig.dev, rig.stream
    B = args.batch  # pairs per GPU per step (weak scaling: per-GPU work is fixed)
    uniq = make_unique_pairs(args.unique, 5000 + 1000 * rig.rank)
    pk = batch.tile_batch(batch.pack_align_batch(uniq, max_features=N_FEATURES), B)
    pk["T_imu_world_cur"] = perturbed_initial_poses(pk["T_imu_world_ref"], 99 + rig.rank)
    cam = capi.Camera.from_dict(uniq[0]["cam"])
    gopt = capi.sparse_align_options()

    ref = capi.Pyramid(ctx, B, W, H, N_LEVELS)
    cur = capi.Pyramid(ctx, B, W, H, N_LEVELS)
    # page-locked host copies (the e2e leg reads these every step); the images optionally in write-combined pages (--pinned wc)
    hb_cur = None
    if args.pinned == "wc":
        hb_cur = capi.HostBuffer(ctx, pk["cur_imgs"].shape, np.uint8, write_combined=True)
        hb_cur.array[...] = pk["cur_imgs"]
        h_cur = torch.from_numpy(hb_cur.array)
    else:
        h_cur = rig.pin(pk["cur_imgs"])
    h = {k: rig.pin(pk[k]) for k in ("T_imu_world_ref", "T_imu_world_cur", "n_features", "px", "f", "depth", "eligible")}
    h_res = torch.zeros(B * capi.ALIGN_RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    d = {k: v.to(dev) for k, v in h.items()}   # device-resident copies (the `value` leg)
    d_cur0 = h_cur.to(dev)
    d_res = torch.zeros(B * capi.ALIGN_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    ref.upload(torch.from_numpy(pk["ref_imgs"]).to(dev))
    ref.build()
    torch.cuda.synchronize()
    cur.upload(d_cur0)                  # the new frames' level-0 images are resident in the pyramid batch before timing
    del d_cur0

    def step_device(ev=None):
        cur.build()                     # levels 1..4 of every new frame
        if ev:
            ev[0].record(stream)
        capi.sparse_align(ctx, [ref], [cur], [cam], pk["T_cam_imu"], d["T_imu_world_ref"], d["T_imu_world_cur"], d["n_features"],
                          d["px"], d["f"], d["depth"], d["eligible"], gopt, results=d_res)
        if ev:
            ev[1].record(stream)

    def step_e2e():
        cur.upload(h_cur, sync=False)   # H2D of the new frames' level-0 images from page-locked memory
        cur.build()
        capi.sparse_align(ctx, [ref], [cur], [cam], pk["T_cam_imu"], h["T_imu_world_ref"], h["T_imu_world_cur"], h["n_features"],
                          h["px"], h["f"], h["depth"], h["eligible"], gopt, results=h_res)  # stages H2D, copies results D2H, syncs

    # ---- value: device-resident ----
    for _ in range(args.warmup):
        step_device()
    rig.sampler.resume()
    rig.barrier()
    launches0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0.record(stream)
    for s in range(args.steps):
        step_device(kev[s])
    e1.record(stream)
    rig.barrier()
    launches = ctx.launches - launches0
    ms_total = rig.max_over_ranks(e0.elapsed_time(e1))
    align_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    value = rig.world * B * args.steps / (ms_total * 1e-3)

    # ---- e2e: host buffers through the C ABI ----
    e2e_ms = rig.timed_e2e(step_e2e, args.steps, args.warmup)
    rig.sampler.pause()
    e2e_value = rig.world * B / (e2e_ms * 1e-3)
    h2d = int(h_cur.numel() + sum(v.numel() * v.element_size() for v in h.values()))
    d2h = int(h_res.numel())

    # the timed device path produced converged poses (guards against timing a no-op), and the e2e leg the same ones
    res = d_res.cpu().numpy().view(capi.ALIGN_RESULT_DTYPE)
    res_h = h_res.numpy().view(capi.ALIGN_RESULT_DTYPE)
    assert (res["n_tracked"] > 100).all() and np.isfinite(res["T_icur_iref"]).all()
    assert np.array_equal(res["T_icur_iref"], res_h["T_icur_iref"]), "host-buffer and device-resident calls disagree"

    if rig.world > 1:  # the only collective of the job: a final gather of the per-pair results
        gathered = shard.gather_to_rank0(res, rig.world * B)
        assert rig.rank != 0 or gathered.shape[0] == rig.world * B

    if rig.rank != 0:
        return None

    # ---- sampled parity: random pairs of the full-size batch against the oracle (outside the timed region) ----
    from oracle import orc
    rng = np.random.default_rng(4242)
    pick = rng.choice(B, size=min(sample_n, B), replace=False)
    keep, max_dq, max_dt = [], 0.0, 0.0
    pyr_cache = {}
    for i in pick:
        u = int(i) % len(uniq)
        dd = uniq[u]
        if u not in pyr_cache:
            pyr_cache[u] = (orc.create_img_pyramid(dd["ref_img"], N_LEVELS), orc.create_img_pyramid(dd["cur_img"], N_LEVELS))
        rf = orc.make_frame(pyr_cache[u][0], dd["cam"], dd["T_cam_imu"], dd["T_imu_world_ref"], dd["px"], dd["f"], dd["depth"], dd["eligible"], keep=keep)
        cf = orc.make_frame(pyr_cache[u][1], dd["cam"], dd["T_cam_imu"], pk["T_imu_world_cur"][i], keep=keep)
        o = orc.sparse_align([rf], [cf], orc.default_align_options())
        dq, dt = pose_diff(np.array(o.T_icur_iref[:]), res["T_icur_iref"][i])
        max_dq, max_dt = max(max_dq, dq), max(max_dt, dt)
        assert dq < 1e-4 and dt < 1e-4, f"pair {i}: pose differs from the oracle ({dq} rad, {dt} m)"
        assert int(res["n_tracked"][i]) == o.n_tracked and list(res["iters"][i][:4]) == list(o.iters[:4]), f"pair {i}: counts differ"
    iters = res["iters"][:, :4].sum(1)
    parity = {"status": "ok", "units_checked": int(len(pick)), "of": B, "max_rot_diff_rad": max_dq, "max_trans_diff_m": max_dt,
              "tolerance": "1e-4 rad / 1e-4 m, n_tracked and per-level iteration counts equal",
              "iterations_per_pair": {"min": int(iters.min()), "mean": float(iters.mean()), "max": int(iters.max())}}

    # ---- single-pair latency (B = 1 through the host-buffer path) ----
    ref1 = capi.Pyramid(ctx, 1, W, H, N_LEVELS); cur1 = capi.Pyramid(ctx, 1, W, H, N_LEVELS)
    ref1.upload(torch.from_numpy(pk["ref_imgs"][:1]).to(dev)); ref1.build()
    h1 = {k: rig.pin(pk[k][:1]) for k in h}
    h1_img = rig.pin(pk["cur_imgs"][:1]); h1_res = torch.zeros(capi.ALIGN_RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    lat, lat_align = [], []
    for i in range(120):
        t = time.perf_counter()
        cur1.upload(h1_img); cur1.build()
        ctx.synchronize()
        t_mid = time.perf_counter()
        capi.sparse_align(ctx, [ref1], [cur1], [cam], pk["T_cam_imu"], h1["T_imu_world_ref"], h1["T_imu_world_cur"], h1["n_features"],
                          h1["px"], h1["f"], h1["depth"], h1["eligible"], gopt, results=h1_res)
        t_end = time.perf_counter()
        if i >= 20:
            lat.append(t_end - t); lat_align.append(t_end - t_mid)
    # launch-latency breakdown of the single-frame call: device time of every stage (CUDA events on the launching stream)
    # next to the host wall clock of the whole call; the gap is launch + staging + synchronisation overhead
    d1 = {k: v[:1].to(dev) for k, v in h1.items()}
    d1_res = torch.zeros(capi.ALIGN_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    stage_us = {"h2d_image": [], "pyramid_kernel": [], "align_kernel": [], "d2h_result": []}
    for i in range(60):
        evs[0].record(stream)
        cur1.upload(h1_img)
        evs[1].record(stream)
        cur1.build()
        evs[2].record(stream)
        capi.sparse_align(ctx, [ref1], [cur1], [cam], pk["T_cam_imu"], d1["T_imu_world_ref"], d1["T_imu_world_cur"], d1["n_features"],
                          d1["px"], d1["f"], d1["depth"], d1["eligible"], gopt, results=d1_res)
        evs[3].record(stream)
        h1_res.copy_(d1_res, non_blocking=True)
        evs[4].record(stream)
        ctx.synchronize(); torch.cuda.synchronize()
        if i >= 10:
            for k, (a, b) in zip(stage_us, zip(evs[:-1], evs[1:])):
                stage_us[k].append(1e3 * a.elapsed_time(b))
    breakdown = {k: float(np.median(v)) for k, v in stage_us.items()}

    return {
        "value": value, "ms_per_step": ms_total / args.steps, "gpu_launches": int(launches), "B": B,
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "h2d_GBs_per_rank": h2d / (e2e_ms * 1e-3) / 1e9,
                "host_pages": "write-combined page-locked (svo_cuda_host_alloc)" if args.pinned == "wc" else "page-locked (torch pin_memory)",
                "note": "new frames' level-0 images + feature arrays H2D from page-locked memory, results D2H, every step"},
        "roofline": roofline("sparse_align_kernel<ILL=0, ROBUST=0, DJ=0, SLOTS=180>", ALGO_BYTES_PER_PAIR, B, align_ms, rig.peaks,
                             "algorithmic bytes 247,044 B/pair; the kernel is FP64-issue / latency bound, not HBM bound (DESIGN.md 4b)",
                             traffic=ncu_traffic("sparse_align_kernel", B), ncu=ncu_pipes("sparse_align_kernel")),
        "parity_sampled": parity,
        "latency": {"p50_ms_pair_e2e": 1e3 * float(np.median(lat)), "p95_ms_pair_e2e": 1e3 * float(np.percentile(lat, 95)),
                    "p50_ms_align_call": 1e3 * float(np.median(lat_align)),
                    "device_us_p50": breakdown,
                    "note": "B=1 through the host-buffer C ABI: image H2D + pyramid + align + result D2H; device_us_p50 = CUDA-event "
                            "time of each stage of one single-frame call (the rest of p50_ms_pair_e2e is launch, staging and sync overhead)"},
        "l2": "inputs larger than L2: %.1f GB of frame data per step per GPU vs 126 MB L2" % ((h_cur.numel() + B * 2 * 119850) / 1e9),
    }


def leg_fast(rig):
    """BASELINE configs[1]: FAST pyramid detection + grid NMS on 1024 synthetic 752x480 frames per GPU (bit-exact corners)."""
    import torch
    from svo_pro_universal_b200 import capi, synth
    args, ctx, dev = rig.args, rig.ctx, rig.dev
    B, NU = args.fast_frames, 16
    uniq = np.stack([synth.make_image(200 + s) for s in range(NU)])
    fidx = np.arange(B) % NU
    h_imgs = rig.pin(uniq[fidx])
    pyr = capi.Pyramid(ctx, B, W, H, N_LEVELS)
    pyr.upload(h_imgs.to(dev))
    opt = capi.detector_options()
    n_cells = capi.grid_cells(W, H, opt.cell_size)[0]
    d_corners = torch.zeros(B * n_cells * capi.CORNER_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    h_corners = torch.zeros(B * n_cells * capi.CORNER_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    rig.sampler.resume()
    l0 = ctx.launches
    ms_all = rig.timed(lambda: capi.fast_detect(ctx, pyr, opt, corners_out=d_corners, fused_pyramid=True), args.steps, args.warmup)
    launches = (ctx.launches - l0) // (args.steps + args.warmup)
    ms_det = rig.timed(lambda: capi.fast_detect(ctx, pyr, opt, corners_out=d_corners), args.steps, 1)

    def step_e2e():
        pyr.upload(h_imgs)
        capi.fast_detect(ctx, pyr, opt, corners_out=h_corners, fused_pyramid=True)   # corners land in host memory, the call synchronises
    e2e_ms = rig.timed_e2e(step_e2e, args.steps, args.warmup)
    rig.sampler.pause()
    got = d_corners.cpu().numpy().view(capi.CORNER_DTYPE).reshape(B, n_cells)
    assert np.array_equal(got, h_corners.numpy().view(capi.CORNER_DTYPE).reshape(B, n_cells))
    if rig.rank != 0:
        return None
    from oracle import orc
    rng = np.random.default_rng(11)
    pick = rng.choice(B, size=min(64, B), replace=False)
    exp = {}
    n_corners = 0
    for i in pick:
        u = int(fidx[i])
        if u not in exp:
            exp[u] = orc.fast_detector(uniq[u])
        for k in ("x", "y", "level", "score"):
            assert np.array_equal(got[i][k], exp[u][k]), f"frame {i}: FAST corner field {k} differs from the oracle"
        n_corners += int((got[i]["score"] > opt.threshold).sum())
    out = {"config": f"{B} synthetic 752x480 frames per GPU ({NU} unique images tiled, every frame its own memory), FAST-10, thr 10, border 8, cell 30, levels 0-2",
           "metric": "frames/s (pyramid + FAST + 3x3 non-max + grid arg-max)", "value": rig.world * B / (ms_all * 1e-3), "unit": "frames/s",
           "scaling": "weak", "ms_per_step": ms_all, "kernel_ms": ms_det, "gpu_launches": int(launches),
           "roofline": roofline("fast_level_kernel<10> (+ key init / decode)", ALGO_BYTES_FAST_ONLY, B, ms_det, rig.peaks,
                                "detection alone re-reads levels 0-2; the kernel is issue bound (DESIGN.md 4a); fused pyramid + detection "
                                "moves %.0f GB/s of the 487,466 B/frame minimum" % (ALGO_BYTES_PER_FRAME * B / (ms_all * 1e-3) / 1e9),
                                traffic=ncu_traffic("fast_level_kernel", B), ncu=ncu_pipes("fast_level_kernel")),
           "e2e": {"value": rig.world * B / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(h_imgs.numel()),
                   "d2h_bytes_per_step": int(h_corners.numel())},
           "l2": "inputs larger than L2 (%.0f MB of level-0 images per step)" % (h_imgs.numel() / 1e6),
           "parity_sampled": {"status": "ok", "units_checked": int(len(pick)), "of": B, "tolerance": "bit-exact x, y, level, score of all 416 cells",
                              "corners_per_frame": n_corners / len(pick)}}
    if rig.world == 1:
        nt = os.cpu_count() or 1
        use_ref = orc.ref_detect_lib() is not None
        pyrs = [orc.ref_create_img_pyramid(uniq[u], 3) if use_ref else orc.create_img_pyramid(uniq[u], 3) for u in range(NU)]

        def one(w, k):
            u = (w + k) % NU
            if use_ref:
                orc.fast_detector_pyr(orc.ref_create_img_pyramid(uniq[u], 3), which="ref")
            else:
                orc.fast_detector(uniq[u], n_levels=3)
            return 1
        del pyrs
        v = threaded_throughput(one, nt, 4.0)
        out["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": nt, "kind": "reference" if use_ref else "port",
                               "sample": "pyramid (3 levels) + fastDetector per frame over the 16 unique images for ~4 s, one frame per thread"
                                         + (", the reference's own fast_neon + feature_detection_utils sources (oracle/_ref/libdetect_ref.so)" if use_ref else "")}
    return out


def leg_match(rig):
    """BASELINE configs[2]: align2D 8x8 patch refinement (+ align1D for edgelets) through findMatchDirect, and the epipolar
    search through findEpipolarMatchDirect, for 2000 features per frame over 256 frame pairs per GPU."""
    import torch
    from svo_pro_universal_b200 import capi, synth
    args, ctx, dev = rig.args, rig.ctx, rig.dev
    NP, NF, NU = args.match_pairs, 2000, 8
    sets = [synth.make_match_set(300 + s, n_features=NF) for s in range(NU)]
    pid = np.arange(NP) % NU
    ref = capi.Pyramid(ctx, NP, W, H, N_LEVELS); cur = capi.Pyramid(ctx, NP, W, H, N_LEVELS)
    ref.upload(rig.t(np.stack([m["ref_img"] for m in sets]))[torch.from_numpy(pid).to(dev)].contiguous())
    cur.upload(rig.t(np.stack([m["cur_img"] for m in sets]))[torch.from_numpy(pid).to(dev)].contiguous())
    ref.build(); cur.build()
    cam = capi.Camera.from_dict(sets[0]["cam"])
    cat = lambda k: np.concatenate([sets[i][k] for i in pid])
    ft = capi.make_features(cat("px"), cat("f"), cat("grad"), cat("type"), cat("level"))
    M = len(ft)
    begin = np.concatenate([[0], np.cumsum([len(sets[i]["px"]) for i in pid])])
    fidx = np.repeat(np.arange(NP), [len(sets[i]["px"]) for i in pid]).astype(np.int32)   # feature -> frame pair (every pair has its own frames)
    T = np.stack([sets[i]["T_cur_ref"] for i in pid])
    rng = np.random.default_rng(1)
    depth, guess = cat("depth"), cat("px_guess")
    inv = 1.0 / depth
    est = inv * rng.uniform(0.7, 1.4, M)
    spread = rng.uniform(0.1, 0.8, M) * inv
    dinv = np.stack([est, est + spread, np.maximum(est - spread, 1e-8)], 1)
    hft = rig.pin(ft.view(np.uint8))
    h = {k: rig.pin(v) for k, v in dict(idx=fidx, T=T, depth=depth, guess=guess, dinv=dinv).items()}
    d_ft = hft.to(dev)
    d = {k: v.to(dev) for k, v in h.items()}
    d_out0 = torch.zeros(M * capi.MATCH_OUT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_out1 = torch.zeros_like(d_out0)
    h_out0 = torch.zeros(M * capi.MATCH_OUT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    h_out1 = torch.zeros_like(h_out0).pin_memory()
    mopt = capi.matcher_options()

    def direct(ftrs, a, out):
        capi.find_match_direct(ctx, ref, cur, cam, cam, a["T"], ftrs, a["depth"], a["guess"], mopt, ref_frame_idx=a["idx"], cur_frame_idx=a["idx"],
                               T_idx=a["idx"], out=out)

    def epipolar(ftrs, a, out):
        capi.find_epipolar_match_direct(ctx, ref, cur, cam, cam, a["T"], ftrs, a["dinv"], mopt, ref_frame_idx=a["idx"], cur_frame_idx=a["idx"],
                                        T_idx=a["idx"], out=out)
    rig.sampler.resume()
    l0 = ctx.launches
    ms_direct = rig.timed(lambda: direct(d_ft, d, d_out0), args.steps, args.warmup, flush=True)
    ms_epi = rig.timed(lambda: epipolar(d_ft, d, d_out1), args.steps, args.warmup, flush=True)
    launches = (ctx.launches - l0) // (args.steps + args.warmup)

    def step_e2e():
        direct(hft, h, h_out0)
        epipolar(hft, h, h_out1)
    e2e_ms = rig.timed_e2e(step_e2e, max(2, args.steps // 2), 1)
    rig.sampler.pause()
    r0 = d_out0.cpu().numpy().view(capi.MATCH_OUT_DTYPE)
    r1 = d_out1.cpu().numpy().view(capi.MATCH_OUT_DTYPE)
    assert np.array_equal(r0["result"], h_out0.numpy().view(capi.MATCH_OUT_DTYPE)["result"])
    assert np.array_equal(r1["result"], h_out1.numpy().view(capi.MATCH_OUT_DTYPE)["result"])
    if rig.rank != 0:
        return None
    from oracle import orc
    rng2 = np.random.default_rng(12)
    pick_pairs = rng2.choice(NP, size=min(8, NP), replace=False)
    n_chk, max_px, keep = 0, 0.0, []
    oopt = orc.default_matcher_options()
    for p in pick_pairs:
        s_ = sets[pid[p]]
        rf = orc.make_frame(orc.create_img_pyramid(s_["ref_img"], 5), s_["cam"], keep=keep)
        cf = orc.make_frame(orc.create_img_pyramid(s_["cur_img"], 5), s_["cam"], keep=keep)
        sel = np.sort(rng2.choice(len(s_["px"]), size=48, replace=False))
        g = begin[p] + sel
        oft = orc.make_features(s_["px"][sel], s_["f"][sel], s_["grad"][sel], s_["type"][sel], s_["level"][sel])
        e0 = orc.find_match_direct_batch(rf, cf, s_["T_cur_ref"], oft, depth[g], guess[g], oopt)
        e1 = orc.find_epipolar_match_direct_batch(rf, cf, s_["T_cur_ref"], oft, dinv[g], oopt)
        for got, exp, what in ((r0[g], e0, "findMatchDirect"), (r1[g], e1, "findEpipolarMatchDirect")):
            assert np.array_equal(got["result"], exp["result"]) and np.array_equal(got["search_level"], exp["search_level"]), f"{what}: pair {p} result codes differ"
            ok = exp["result"] == 0
            if ok.any():
                dpx = float(np.abs(got["px_cur"][ok] - exp["px_cur"][ok]).max())
                max_px = max(max_px, dpx)
                assert dpx < 1e-3, f"{what}: pair {p} sub-pixel position differs by {dpx} px"
        ok = e1["result"] == 0
        if ok.any():
            assert np.allclose(r1[g]["depth"][ok], e1["depth"][ok], rtol=1e-6), f"pair {p}: triangulated depth differs"
        n_chk += len(sel)
    ms_step = ms_direct + ms_epi
    out = {"config": f"{M} features per GPU = {NP} frame pairs x ~{NF} ({NU} unique pairs tiled, every pair its own frames in HBM), 25 % edgelets (align1D), "
                     "inverse-depth spread 10-80 % for the epipolar search (unit sphere, <= 100 steps)",
           "metric": "features/s through findMatchDirect (warp + align2D / align1D) AND findEpipolarMatchDirect (warp + ZMSSD scan + sub-pixel + triangulation)",
           "value": rig.world * M / (ms_step * 1e-3), "unit": "features/s", "scaling": "weak", "ms_per_step": ms_step,
           "kernel_ms": {"find_match_direct": ms_direct, "find_epipolar_match_direct": ms_epi}, "gpu_launches": int(launches),
           "success_frac": {"find_match_direct": float((r0["result"] == 0).mean()), "find_epipolar_match_direct": float((r1["result"] == 0).mean())},
           "mean_epi_length_px": float(r1["epi_length_pyramid"].mean()),
           "roofline": roofline("match_kernel<0> + epipolar match kernels", 2 * ALGO_BYTES_PER_FEATURE, M, ms_step, rig.peaks,
                                "333 B/feature compulsory per call; both calls are bound by issue slots (ordered float sums, per-feature control flow), not memory (DESIGN.md 4c)",
                                traffic=(lambda a, b: None if a is None or b is None else a + b)(ncu_traffic("match_kernel<0>", M), ncu_traffic("match_kernel<1>", M)),
                                ncu=ncu_pipes("match_kernel<0>")),
           "e2e": {"value": rig.world * M / (e2e_ms * 1e-3), "unit": "features/s",
                   "h2d_bytes_per_step": int(2 * (hft.numel() + sum(v.numel() * v.element_size() for v in h.values()))),
                   "d2h_bytes_per_step": int(h_out0.numel() + h_out1.numel()),
                   "note": "feature arrays in, svo_match_out records back, both calls; the frames are resident (they were uploaded when created)"},
           "l2": "L2 flushed between timed steps (256 MB fill)",
           "parity_sampled": {"status": "ok", "units_checked": int(n_chk), "of": M, "max_px_diff": max_px,
                              "tolerance": "result codes and search levels equal, sub-pixel position 1e-3 px, depth 1e-6 relative"}}
    if rig.world == 1:
        nt = os.cpu_count() or 1
        s0 = sets[0]
        rf = orc.make_frame(orc.create_img_pyramid(s0["ref_img"], 5), s0["cam"], keep=keep)
        cf = orc.make_frame(orc.create_img_pyramid(s0["cur_img"], 5), s0["cam"], keep=keep)
        n0 = len(s0["px"])
        oft = orc.make_features(s0["px"], s0["f"], s0["grad"], s0["type"], s0["level"])
        t0 = time.perf_counter(); reps = 0
        while time.perf_counter() - t0 < 5.0:
            orc.find_match_direct_batch(rf, cf, s0["T_cur_ref"], oft, depth[:n0], guess[:n0], oopt, n_threads=nt)
            orc.find_epipolar_match_direct_batch(rf, cf, s0["T_cur_ref"], oft, dinv[:n0], oopt, n_threads=nt)
            reps += 1
        out["cpu_baseline"] = {"value": n0 * reps / (time.perf_counter() - t0), "unit": "features/s", "cores": nt, "kind": "port",
                               "sample": f"both calls on the {n0} features of one frame pair, repeated for ~5 s, {nt} threads (oracle port, pinned to the "
                                         "reference's compiled matcher.cpp / feature_alignment.cpp by tests/)"}
    return out


def leg_seeds(rig):
    """BASELINE configs[3]: DepthFilter seed updates, 50 k seeds x 64 ordered observations with epipolar matching per GPU; and the
    pure Gaussian x Beta filter update on the same shape."""
    import torch
    from svo_pro_universal_b200 import capi, synth
    args, ctx, dev = rig.args, rig.ctx, rig.dev
    S, O = args.seeds, args.seed_obs
    NSEQ_U, NOBS_U = 4, 16
    seqs = [synth.make_seed_sequence(400 + s, n_seeds=400, n_obs=NOBS_U) for s in range(NSEQ_U)]
    per = min(len(q["px"]) for q in seqs)
    NSEQ = max(1, S // per)                 # keyframes, every one with its own frame and its own observation frames in HBM
    sid = np.arange(NSEQ) % NSEQ_U
    ref = capi.Pyramid(ctx, NSEQ, W, H, N_LEVELS); cur = capi.Pyramid(ctx, NSEQ * NOBS_U, W, H, N_LEVELS)
    ref.upload(rig.t(np.stack([q["ref_img"] for q in seqs]))[torch.from_numpy(sid).to(dev)].contiguous())
    cur_u = rig.t(np.stack([im for q in seqs for im in q["cur_imgs"]]))                       # [NSEQ_U * NOBS_U]
    cur_sel = (sid[:, None] * NOBS_U + np.arange(NOBS_U)[None, :]).reshape(-1)
    for c0 in range(0, len(cur_sel), 512):
        cur.upload(cur_u[torch.from_numpy(cur_sel[c0:c0 + 512]).to(dev)].contiguous(), first=c0)
    del cur_u
    ref.build(); cur.build()
    cam = capi.Camera.from_dict(seqs[0]["cam"])
    catq = lambda k: np.concatenate([seqs[i][k][:per] for i in sid])
    ftq = capi.make_features(catq("px"), catq("f"), catq("grad"), catq("type").astype(np.int32), catq("level"))
    Sq = len(ftq)
    kf_of_seed = np.repeat(np.arange(NSEQ), per).astype(np.int32)
    obs = np.arange(O) % NOBS_U
    obs_frame = (kf_of_seed[None, :] * NOBS_U + obs[:, None]).astype(np.int32)                # frame of observation o of seed s
    obs_T = (sid[kf_of_seed][None, :] * NOBS_U + obs[:, None]).astype(np.int32)              # its transformation (unique sequences)
    Tq = np.concatenate([q["T_cur_ref"] for q in seqs])
    types0, st0 = catq("type").astype(np.uint8), catq("state")
    mu = np.full(Sq, seqs[0]["mu_range"])
    h = {k: rig.pin(v) for k, v in dict(ft=ftq.view(np.uint8), mu=mu, kf=kf_of_seed, obs=obs_frame, obsT=obs_T, T=Tq).items()}
    h_types, h_st = rig.pin(types0), rig.pin(st0)
    d = {k: v.to(dev) for k, v in h.items()}
    d_types0, d_st0 = rig.t(types0), rig.t(st0)
    d_types, d_st = d_types0.clone(), d_st0.clone()
    mopt, dopt = capi.matcher_options(), capi.depth_filter_options()
    res = {}

    def step_device():
        d_types.copy_(d_types0); d_st.copy_(d_st0)
        res["n"], _ = capi.update_seeds(ctx, ref, cur, cam, cam, d["ft"], d_types, d_st, d["mu"], d["obs"], d["obsT"], d["T"], mopt, dopt,
                                        ref_frame_idx=d["kf"], want_match_results=False)
    rig.sampler.resume()
    l0 = ctx.launches
    ms = rig.timed(step_device, max(2, args.steps // 2), 1, flush=True)
    launches = (ctx.launches - l0) // (max(2, args.steps // 2) + 1)
    g_types, g_st, n_succ = d_types.cpu().numpy(), d_st.cpu().numpy(), int(res["n"].item())
    w_types, w_st = h_types.numpy().copy(), h_st.numpy().copy()

    def step_e2e():
        w_types[:] = types0; w_st[:] = st0
        res["nh"], _ = capi.update_seeds(ctx, ref, cur, cam, cam, h["ft"].numpy().view(capi.FEATURE_DTYPE), w_types, w_st, h["mu"].numpy(), h["obs"].numpy(),
                                         h["obsT"].numpy(), h["T"].numpy(), mopt, dopt, ref_frame_idx=h["kf"].numpy(), want_match_results=False)
    e2e_ms = rig.timed_e2e(step_e2e, 2, 1)
    assert np.array_equal(w_types, g_types) and np.array_equal(w_st, g_st) and int(res["nh"][0]) == n_succ
    # pure filter update, one launch for the 64 ordered updates
    rng = np.random.default_rng(2)
    fs0 = np.tile(np.array([0.25, (1 / 1.5) ** 2 / 36.0, 10.0, 10.0]), (Sq, 1))
    fz = 0.25 + rng.normal(size=(O, Sq)) * 0.01
    ft2 = np.ascontiguousarray(np.broadcast_to((1e-4 / (np.arange(O) + 1))[:, None], (O, Sq)))
    d_fs0, d_fz, d_ft2, d_fmu = rig.t(fs0), rig.t(fz), rig.t(ft2), rig.t(np.full(Sq, 1 / 1.5))
    d_fs = d_fs0.clone()

    def step_filter():
        d_fs.copy_(d_fs0)
        capi.update_filter_seq(ctx, d_fz, d_ft2, d_fmu, d_fs)
    ms_f = rig.timed(step_filter, args.steps, args.warmup, flush=True)
    ms_copy = rig.timed(lambda: d_fs.copy_(d_fs0), args.steps, 1, flush=True)
    rig.sampler.pause()
    step_filter()   # the reset-copy timing above left the initial state in d_fs
    g_fs = d_fs.cpu().numpy()
    if rig.rank != 0:
        return None
    from oracle import orc
    rng2 = np.random.default_rng(13)
    keep, n_chk = [], 0
    oopt = orc.default_matcher_options()
    for k in rng2.choice(NSEQ, size=min(4, NSEQ), replace=False):
        q = seqs[sid[k]]
        rf = orc.make_frame(orc.create_img_pyramid(q["ref_img"], 5), q["cam"], keep=keep)
        cfs = [orc.make_frame(orc.create_img_pyramid(q["cur_imgs"][o], 5), q["cam"], keep=keep) for o in obs]
        sel = np.sort(rng2.choice(per, size=16, replace=False))
        oft = orc.make_features(q["px"][sel], q["f"][sel], q["grad"][sel], q["type"][sel].astype(np.int32), q["level"][sel])
        ty, st = np.ascontiguousarray(q["type"][sel].astype(np.uint8)), np.ascontiguousarray(q["state"][sel])
        orc.update_seeds(rf, cfs, q["T_cur_ref"][obs], oft, ty, st, q["mu_range"], oopt)
        g = k * per + sel
        assert np.array_equal(g_types[g], ty), f"keyframe {k}: seed types differ from the oracle after {O} observations"
        assert np.allclose(g_st[g, :2], st[:, :2], rtol=1e-4, atol=0) and np.allclose(g_st[g, 2:], st[:, 2:], rtol=1e-4), f"keyframe {k}: seed states differ"
        n_chk += len(sel)
    pick = rng2.choice(Sq, size=64, replace=False)
    fexp = np.ascontiguousarray(fs0[pick])
    for o in range(O):
        zo, to, mo = np.ascontiguousarray(fz[o, pick]), np.ascontiguousarray(ft2[o, pick]), np.full(len(pick), 1 / 1.5)
        orc.lib().orc_update_filter_vogiatzis_batch(len(pick), zo.ctypes.data_as(orc.f64p), to.ctypes.data_as(orc.f64p), mo.ctypes.data_as(orc.f64p),
                                                    fexp.ctypes.data_as(orc.f64p), None, 1)
    assert np.allclose(g_fs[pick], fexp, rtol=1e-4, atol=0), "fused filter updates differ from 64 oracle updates"
    conv = np.isin(g_types, (synth.K_CORNER_SEED_CONV, synth.K_EDGELET_SEED_CONV))
    out = {"config": f"{Sq} seeds x {O} ordered observations per GPU ({NSEQ} keyframes x {per} seeds, {NOBS_U} observation frames per keyframe revisited in order, "
                     f"{NSEQ + NSEQ * NOBS_U} frames in HBM; {NSEQ_U} unique sequences tiled); ONE svo_cuda_update_seeds call: a step + a match kernel per observation wave and seed group (4 concurrent groups on the context's side streams; launches counted in gpu_launches)",
           "metric": "seed-observations/s through depth_filter_utils::updateSeed (visibility gate + epipolar match + tau + Vogiatzis update + convergence)",
           "value": rig.world * Sq * O / (ms * 1e-3), "unit": "seed-observations/s", "scaling": "weak", "ms_per_step": ms, "kernel_ms": ms,
           "gpu_launches": int(launches), "success_frac": n_succ / (Sq * O), "converged_frac": float(conv.mean()),
           "roofline": roofline("seed_step_kernel + seed_match_kernel (per observation wave)", ALGO_BYTES_PER_FEATURE + ALGO_BYTES_PER_UPDATE, Sq * O, ms, rig.peaks,
                                "333 B (match) + 80 B (state) per seed-observation compulsory; the match kernel is issue bound like (c)"),
           "filter_only": {"value": rig.world * Sq * O / ((ms_f - ms_copy) * 1e-3), "unit": "updates/s", "kernel_ms": ms_f - ms_copy,
                           "note": f"svo_cuda_update_filter_seq: {O} ordered Vogiatzis updates per seed in one launch, state in registers (timed with the state reset copy, "
                                   f"{ms_copy:.4f} ms, subtracted)",
                           "roofline": roofline("filter_seq_kernel<false>", 16 + 64.0 / O, Sq * O, max(ms_f - ms_copy, 1e-6), rig.peaks,
                                                "16 B streamed per update + 64 B of state per seed; bound by the FP64 pipe (~150 FP64 instructions per update at 64 lanes / clk / SM)")},
           "e2e": {"value": rig.world * Sq * O / (e2e_ms * 1e-3), "unit": "seed-observations/s",
                   "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in h.values()) + types0.nbytes + st0.nbytes),
                   "d2h_bytes_per_step": int(types0.nbytes + st0.nbytes + 4),
                   "note": "seed tables and observation lists in, seed types / states back; the frames are resident"},
           "l2": "L2 flushed between timed steps (256 MB fill)",
           "parity_sampled": {"status": "ok", "units_checked": int(n_chk), "of": Sq, "filter_units_checked": 64,
                              "tolerance": "seed types equal, mean / variance / a / b 1e-4 relative after all observations; fused filter 1e-4 relative after 64 updates (the variance update cancels ~7 digits on converged seeds)"}}
    if rig.world == 1:
        nt = os.cpu_count() or 1
        q0 = seqs[0]
        rf = orc.make_frame(orc.create_img_pyramid(q0["ref_img"], 5), q0["cam"], keep=keep)
        cfs = [orc.make_frame(orc.create_img_pyramid(im, 5), q0["cam"], keep=keep) for im in q0["cur_imgs"]]
        oft = orc.make_features(q0["px"][:per], q0["f"][:per], q0["grad"][:per], q0["type"][:per].astype(np.int32), q0["level"][:per])
        t0 = time.perf_counter(); reps = 0
        while time.perf_counter() - t0 < 5.0:
            ty = q0["type"][:per].copy(); stt = q0["state"][:per].copy()
            orc.update_seeds(rf, cfs, q0["T_cur_ref"], oft, ty, stt, q0["mu_range"], oopt, n_threads=nt); reps += 1
        out["cpu_baseline"] = {"value": per * NOBS_U * reps / (time.perf_counter() - t0), "unit": "seed-observations/s", "cores": nt, "kind": "port",
                               "sample": f"{per} seeds x {NOBS_U} observations of one keyframe, repeated for ~5 s, {nt} threads (oracle port, pinned to the reference's "
                                         "compiled depth_filter.cpp / matcher.cpp by tests/)"}
    return out


def leg_frontend(rig):
    """BASELINE configs[4]: the full front-end batch (pyramid + 2-camera sparse align + Reprojector + PoseOptimizer + DepthFilter
    update + FastGrad detector) on 8192 stereo frame pairs in total, sharded over the ranks by contiguous blocks (strong scaling)."""
    import torch
    from svo_pro_universal_b200 import capi, frontend, shard, synth
    args, ctx, dev = rig.args, rig.ctx, rig.dev
    lo, hi = shard.partition(args.frontend_pairs, rig.world, rig.rank)
    Bl = hi - lo
    scenes = [frontend.make_stereo_scene(81 + s) for s in range(4)]
    fb = frontend.StereoFrontendBatch(ctx, scenes, Bl, dev)
    torch.cuda.set_stream(fb.stream)
    old_stream, rig.stream = rig.stream, fb.stream
    steps = max(2, args.steps // 2)
    rig.sampler.resume()
    l0 = ctx.launches
    ms = rig.timed(fb.step, steps, args.warmup)
    launches = (ctx.launches - l0) // (steps + args.warmup)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(fb.STAGES) + 1)]
    evs[0].record(fb.stream)
    fb.step(lambda i: evs[i + 1].record(fb.stream))
    torch.cuda.synchronize()
    stages = {n: evs[i].elapsed_time(evs[i + 1]) for i, n in enumerate(fb.STAGES)}
    # e2e: the two new frames of every pair come from pinned host memory, the per-pair poses and the new frame's features go back
    h_cur = torch.empty((2 * Bl, H, W), dtype=torch.uint8).pin_memory()
    for c in range(2):
        for s_ in range(len(scenes)):
            h_cur[c * Bl:(c + 1) * Bl][torch.from_numpy(np.flatnonzero(fb.sid == s_))] = torch.from_numpy(scenes[s_]["imgs"][f"c{c}"])
    h_align = torch.zeros(Bl * capi.ALIGN_RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    h_po = torch.zeros(Bl * capi.POSE_OPT_RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    h_corners = torch.zeros(fb.d_corners.numel(), dtype=torch.uint8).pin_memory()
    h_edgelets = torch.zeros(fb.d_corners.numel(), dtype=torch.uint8).pin_memory()

    def step_e2e():
        with torch.cuda.stream(fb.stream):
            fb.cur.upload(h_cur)
            fb._step(lambda i: None)
            h_align.copy_(fb.d_align, non_blocking=True); h_po.copy_(fb.d_po, non_blocking=True)
            h_corners.copy_(fb.d_corners, non_blocking=True); h_edgelets.copy_(fb.d_edgelets, non_blocking=True)
        fb.stream.synchronize()
    e2e_ms = rig.timed_e2e(step_e2e, 2, 1)
    rig.sampler.pause()
    out_ = fb.results()
    assert (out_["align"]["n_tracked"] > 250).all() and out_["reproj_stats"]["n_matches"].mean() > 60
    seed_lv = fb.seed_ftrs.cpu().numpy().view(capi.FEATURE_DTYPE)["level"]
    n_seed = len(scenes[0]["seed_px"])
    if rig.world > 1:  # the only collective: gather the per-pair poses to rank 0
        g = shard.gather_to_rank0(out_["align"]["T_icur_iref"].copy(), args.frontend_pairs)
        assert rig.rank != 0 or g.shape == (args.frontend_pairs, 7)
    fb.release()
    rig.stream = old_stream
    torch.cuda.set_stream(old_stream)
    ctx.set_stream(old_stream.cuda_stream)
    if rig.rank != 0:
        return None
    from oracle import orc
    rng2 = np.random.default_rng(14)
    pick = rng2.choice(Bl, size=min(12, Bl), replace=False)
    seed_begin = np.concatenate([[0], np.cumsum([len(scenes[s_]["seed_px"]) for s_ in fb.sid])])
    max_dq = max_dt = 0.0
    for i in pick:
        sc = scenes[fb.sid[i]]
        keep, cam = [], sc["cam"]
        pyr = {k: orc.create_img_pyramid(v, 5) for k, v in sc["imgs"].items()}
        rfs = [orc.make_frame(pyr[f"r{c}"], cam, sc["T_cam_imu"][c], sc["T_imu_world_ref"], sc["px"][c], sc["f"][c], sc["depth"][c], keep=keep) for c in range(2)]
        cfs = [orc.make_frame(pyr[f"c{c}"], cam, sc["T_cam_imu"][c], sc["T_imu_world_ref"], keep=keep) for c in range(2)]
        r = orc.sparse_align(rfs, cfs, orc.default_align_options(estimate_illumination_gain=1, estimate_illumination_offset=1))
        g = out_["align"][i]
        dq, dt = pose_diff(g["T_icur_iref"], np.array(r.T_icur_iref[:]))
        max_dq, max_dt = max(max_dq, dq), max(max_dt, dt)
        assert dq < 1e-4 and dt < 1e-4 and g["n_tracked"] == r.n_tracked and list(g["iters"][:4]) == list(r.iters[:4]), f"pair {i}: stereo alignment differs"
        co = orc.fast_detector(sc["imgs"]["c0"])
        for k in ("x", "y", "level", "score"):
            assert np.array_equal(out_["corners"][i][k], co[k]), f"pair {i}: FAST corner field {k} differs"
        eo = orc.edgelet_detector_v2(pyr["c0"], 100, 8, 30, (co["score"] > 10).astype(np.uint8))
        for k in ("x", "y", "level", "score", "angle"):
            assert np.array_equal(out_["edgelets"][i][k], eo[k]), f"pair {i}: edgelet field {k} differs"
        a, b = seed_begin[i], seed_begin[i + 1]
        n = b - a
        oft = orc.make_features(sc["seed_px"], sc["seed_f"], np.tile([1.0, 0.0], (n, 1)), np.full(n, synth.K_CORNER_SEED, np.int32), seed_lv[a:b])
        ty, st = np.full(n, synth.K_CORNER_SEED, np.uint8), sc["seed_state"].copy()
        orc.update_seeds(orc.make_frame(pyr["r0"], cam, keep=keep), [orc.make_frame(pyr["c0"], cam, keep=keep)], sc["T_cur_ref_gt"].reshape(1, 7), oft, ty, st,
                         sc["seed_mu_range"], orc.default_matcher_options())
        assert np.array_equal(out_["seed_types"][a:b], ty) and np.allclose(out_["seed_state"][a:b], st, rtol=1e-4), f"pair {i}: seed update differs"
    out = {"config": f"{args.frontend_pairs} synthetic stereo frame pairs in total, {Bl} on this GPU (4 unique scenes tiled; every frame resident in HBM: "
                     f"{4 * Bl * 483360 / 1e9:.1f} GB of pyramids per GPU), 180 + 150 features, {n_seed} seeds per pair",
           "metric": "stereo frame pairs/s through the front-end chain (pyramid + 2-camera SparseImgAlign + Reprojector + PoseOptimizer + DepthFilter update + FastGrad detector)",
           "value": args.frontend_pairs / (ms * 1e-3), "unit": "stereo pairs/s", "scaling": "strong", "ms_per_step": ms, "kernel_ms": stages,
           "gpu_launches": int(launches), "mean_matches_per_frame": float(out_["reproj_stats"]["n_matches"].mean()),
           "seed_success_frac": out_["n_seed_ok"] / max(1, fb.S),
           "roofline": roofline("front-end chain (all stages)", 2 * ALGO_BYTES_PER_FRAME + 2 * ALGO_BYTES_PER_PAIR, Bl, ms, rig.peaks,
                                "compulsory bytes per stereo pair: two new frames through the pyramid / detector + the 2-camera alignment levels; "
                                "every stage is issue or latency bound (stage times in kernel_ms)"),
           "e2e": {"value": args.frontend_pairs / (e2e_ms * 1e-3), "unit": "stereo pairs/s", "h2d_bytes_per_step": int(h_cur.numel()),
                   "d2h_bytes_per_step": int(h_align.numel() + h_po.numel() + h_corners.numel() + h_edgelets.numel()),
                   "note": "the two new level-0 images of every pair H2D from pinned memory, alignment + pose-optimiser results and the new left frame's corners / edgelets D2H"},
           "l2": "inputs larger than L2 (GBs of frames per step)",
           "parity_sampled": {"status": "ok", "units_checked": int(len(pick)), "of": Bl, "max_rot_diff_rad": max_dq, "max_trans_diff_m": max_dt,
                              "tolerance": "stereo alignment pose 1e-4 rad / 1e-4 m + equal iteration counts, FAST corners and edgelets bit-exact, seed types equal / states 1e-4 relative "
                                           "(Reprojector and PoseOptimizer stages are checked entry by entry in tests/test_gpu_frontend_chain.py)"}}
    if rig.world == 1:
        nt = os.cpu_count() or 1
        prep = cpu_chain_prepare(scenes)
        cpu_chain_once(prep[0])  # page the oracle in
        v = threaded_throughput(lambda w, k: cpu_chain_once(prep[(w + k) % len(prep)]), nt, 6.0)
        out["cpu_baseline"] = {"value": v, "unit": "stereo pairs/s", "cores": nt, "kind": "port",
                               "sample": f"the same chain per pair with the oracle port over the 4 unique scenes for ~6 s, one pair per thread, {nt} threads"}
    return out


def run_ours(args):
    rig = Rig(args)
    if rig.rank == 0:
        rig.sampler.start()
        rig.sampler.pause()
    head = leg_headline(rig)
    paths = {}
    want = [p for p in args.paths.split(",") if p]
    legs = {"fast_1024": leg_fast, "match_512k": leg_match, "seeds_50k_x64": leg_seeds, "frontend_8192": leg_frontend}
    for name in want:
        r = legs[name](rig)
        rig.torch.cuda.empty_cache()
        if rig.rank == 0:
            paths[name] = r
    if rig.rank != 0:
        if rig.world > 1:
            rig.dist.destroy_process_group()
        return
    clocks = rig.sampler.stop()
    cpu_bl = cpu_baseline(os.cpu_count() or 1, 8.0) if rig.world == 1 else None
    out = {
        "metric": METRIC, "value": head["value"], "unit": "pairs/s", "n_gpus": rig.world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "SparseImgAlign batch: pyramid of the new frame + run() per pair, 752x480, levels 4->1, 180 features, 4x4 patches "
                               "(BASELINE configs[0] batched; per-pair perturbed initial pose)",
                   "pairs_per_gpu_per_step": head["B"], "unique_pairs": args.unique,
                   "parallelism": f"frame-pair sharding x{rig.world}, no collective in the hot path", "l2": head["l2"]},
        "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "clocks": clocks, "roofline": head["roofline"],
        "cpu_baseline": cpu_bl, "parity_sampled": head["parity_sampled"], "latency": head["latency"], "paths": paths,
    }
    print(json.dumps(out))
    if rig.world > 1:
        rig.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="frame pairs per GPU per step (headline)")
    ap.add_argument("--unique", type=int, default=32, help="unique synthetic pairs generated per rank (tiled to --batch)")
    ap.add_argument("--pinned", default="torch", choices=["torch", "wc"], help="host pages of the e2e image uploads: torch pin_memory or write-combined")
    ap.add_argument("--paths", default="fast_1024,match_512k,seeds_50k_x64,frontend_8192",
                    help="comma-separated BASELINE config legs to run after the headline ('' = none)")
    ap.add_argument("--fast-frames", type=int, default=1024)
    ap.add_argument("--match-pairs", type=int, default=256)
    ap.add_argument("--seeds", type=int, default=50000)
    ap.add_argument("--seed-obs", type=int, default=64)
    ap.add_argument("--frontend-pairs", type=int, default=8192)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
