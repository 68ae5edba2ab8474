"""Assembles profiles/<tag>_current.md from the files one `bash tools/gpu_round.sh <tag> tests smoke bench paths launches ncu`
call merged into gpurun_out/ (tests log, bench lines, per-path lines, ncu launch lists, ncu --set full summaries, source hot spots).
    python tools/make_profile.py r01e "title line" [notes.md] > profiles/r01e_current.md"""
import glob
import io
import json
import os
import subprocess
import sys
from contextlib import redirect_stdout

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ncu_lines  # noqa: E402
import ncu_summary  # noqa: E402
import summarize_launches  # noqa: E402

tag, title = sys.argv[1], sys.argv[2]
G = os.path.join(ROOT, "gpurun_out")


def rd(name):
    p = os.path.join(G, f"{tag}_{name}")
    return open(p).read().strip() if os.path.exists(p) else ""


def cap(fn, *a):
    buf = io.StringIO()
    with redirect_stdout(buf):
        fn(*a)
    return buf.getvalue().strip()


print(f"# {title}\n")
print(f"Command: `bash tools/gpu_round.sh {tag} tests smoke bench paths launches ncu` (the script holds the exact ncu lines). GPU: `{rd('gpu.txt').splitlines()[-1] if rd('gpu.txt') else '?'}`\n")
if len(sys.argv) > 3:
    print(open(sys.argv[3]).read().strip() + "\n")
print("## GPU parity tests\n```\n" + "\n".join(rd("tests.log").splitlines()[-3:]) + "\n" + rd("smoke.log").splitlines()[-1] + "\n```\n")
bench = rd("bench.json")
print("## bench.py (N=1, --steps 10 --warmup 3)\n```json\n" + bench + "\n```\n")
print("## bench.py --impl reference\n```json\n" + rd("bench_ref.json") + "\n```\n")
try:
    b = json.loads(bench.splitlines()[-1])
    lat = b["latency"]
    print("## Single-frame call (B = 1): launch-latency breakdown\n\n| stage (device time, CUDA events, p50) | us |\n|---|---|")
    tot = 0.0
    for k, v in lat["device_us_p50"].items():
        print(f"| {k} | {v:.1f} |"); tot += v
    print(f"| **sum of device stages** | {tot:.1f} |")
    print(f"| host wall clock of the whole call, p50 (image H2D + pyramid + align + result D2H through the host-buffer C ABI) | {1e3 * lat['p50_ms_pair_e2e']:.1f} |")
    print(f"| launch + staging (one packed H2D / D2H copy through the context's arena) + synchronisation overhead | {1e3 * lat['p50_ms_pair_e2e'] - tot:.1f} |")
    cb = b["cpu_baseline"]
    print(f"| reference CPU path ({cb['kind']}), single thread, p50 | {1e3 * cb['latency_ms_p50_single_thread']:.0f} |\n")
except Exception as e:  # noqa: BLE001
    print(f"(no latency table: {e})\n")
print("## tools/bench_kernels.py (per-path, CUDA events)\n```json\n" + rd("paths.jsonl") + "\n```\n")
for name, what in (("launches_bench.csv", "bench.py --steps 2 --warmup 3 --batch 1184"), ("launches_paths.csv", "tools/prof_paths.py"),
                   ("launches_frontend.csv", "tools/bench_frontend.py --pairs 8192 --steps 1 --warmup 1 (the configs[4] chain, 8192 stereo pairs)")):
    p = os.path.join(G, f"{tag}_{name}")
    if os.path.exists(p):
        print(f"## ncu launch list of `{what}` (gpu__time_duration.sum, cold-cache, serialised)\n" + cap(summarize_launches.main, p) + "\n")
reps = sorted(glob.glob(os.path.join(G, f"{tag}_ncu_*.ncu-rep")))
print("## ncu --set full, one launch per kernel (tools/prof_paths.py)")
print(cap(ncu_summary.main, reps) + "\n")
for rep in reps:
    k = os.path.basename(rep)[len(tag) + 5:-8]
    print(f"## Source-level hot spots: {k} (tools/ncu_lines.py)\n")
    print(cap(ncu_lines.main, rep, 16) + "\n")
