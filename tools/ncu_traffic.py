"""DRAM traffic per launch of the kernels of one capture -> profiles/ncu_traffic.json (read by bench.py for `roofline.traffic`).
    python tools/ncu_traffic.py <tag>        # reads gpurun_out/<tag>_ncu_*.ncu-rep (bash tools/gpu_round.sh <tag> ncu)
Units per captured launch are those of tools/prof_paths.py (the workload every capture runs)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ncu_summary  # noqa: E402

# capture file stem -> (key used by bench.py, units processed by the captured launch, unit)
UNITS = {
    "sparse_align_kernel": ("sparse_align_kernel", 1184, "pairs"),
    "pyr_down": ("pyr_down_fused_kernel", 1184, "frames"),
    "fast_level": ("fast_level_kernel", 1184, "frames"),
    "match_kernel__int_0": ("match_kernel<0>", 64 * 2000, "features"),
    "match_kernel__int_1": ("match_kernel<1>", 64 * 2000, "features"),
    "filter_seq_kernel": ("filter_seq_kernel", 50000 * 16, "updates"),
    "seed_match_kernel": ("seed_match_kernel", 12800, "work items (one observation wave)"),
    "seed_step_kernel": ("seed_step_kernel", 12800, "seeds (one observation wave)"),
    "reproj_match": ("reproj_match_kernel", 296, "frames"),
    "pose_optimize_kernel": ("pose_optimize_kernel", 4736, "bundles"),
    "edgelet_score": ("edgelet_score_kernel", 1184, "frames"),
}


def to_bytes(v, u):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


PIPE_KEYS = {
    "issue_slots_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "fp64_pipe_pct": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "lsu_pipe_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smem_wavefronts_pct": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "achieved_occupancy_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "active_threads_per_instruction": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "warp_instructions": "smsp__inst_executed.sum",
    "registers_per_thread": "launch__registers_per_thread",
}


def pipes(d):
    """What the capture says bounds the kernel: pipe / issue utilisation of the captured launch (ncu --set full, under the profiler:
    shares, not times)."""
    out = {}
    for k, m in PIPE_KEYS.items():
        if m in d:
            try:
                out[k] = float(d[m][0].replace(",", ""))
            except ValueError:
                pass
    return out


def main(tag):
    out = {"capture": tag, "command": f"bash tools/gpu_round.sh {tag} ncu  (ncu --set full --clock-control none, tools/prof_paths.py)", "kernels": {}}
    for stem, (key, units, unit) in UNITS.items():
        rep = os.path.join(ROOT, "gpurun_out", f"{tag}_ncu_{stem}.ncu-rep")
        if not os.path.exists(rep):
            continue
        d = ncu_summary.load(rep)[0]
        rd, wr = to_bytes(*d["dram__bytes_read.sum"]), to_bytes(*d["dram__bytes_write.sum"])
        out["kernels"][key] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "units_per_launch": units, "unit": unit,
                               "dram_bytes_per_unit": (rd + wr) / units, "kernel_name": d.get("Kernel Name", ("?", ""))[0][:120],
                               "duration_us": float(d["gpu__time_duration.sum"][0].replace(",", "")) * {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(d["gpu__time_duration.sum"][1], 1),
                               "file": f"{tag}_ncu_{stem}.ncu-rep", "pipes": pipes(d)}
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
