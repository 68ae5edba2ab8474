"""Parity at BASELINE.json's FULL sizes: the five bench workloads (4096 alignment pairs, 1024 FAST frames, 512 k matcher features,
50 k seeds x 64 observations, the 8192-pair front-end chain) run through bench.py's legs, each of which checks a random sample of its
units against the CPU oracle after the timed region (poses 1e-4 rad / 1e-4 m + equal iteration counts, corners bit-exact, sub-pixel
positions 1e-3 px, seed states 1e-4 relative) and fails loudly on a mismatch. A batch-index or grid-dimension bug that the small parity
cases cannot see (thousands of units, multi-chunk launches) shows up here."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_workloads_pass_their_sampled_oracle_checks():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3"], capture_output=True, text=True, cwd=ROOT,
                       timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    checks = {"headline": line["parity_sampled"]}
    for name, leg in line["paths"].items():
        checks[name] = leg["parity_sampled"]
    assert set(checks) == {"headline", "fast_1024", "match_512k", "seeds_50k_x64", "frontend_8192"}
    for name, c in checks.items():
        assert c["status"] == "ok" and c["units_checked"] >= 12 and c["of"] >= 1024, (name, c)
    assert line["config"]["pairs_per_gpu_per_step"] == 4096 and line["gpu_launches"] > 0
    assert line["roofline"]["frac"] > 0 and line["e2e"]["h2d_bytes_per_step"] > 0
