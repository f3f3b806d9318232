"""bench.py contract pieces that need no GPU: the reference arm's JSON line and the workload table."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(300)
def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    """`bench.py --impl reference` times the reference's CPU scoring idiom on a bounded sample and prints ONE JSON
    line: same metric/config as the GPU arm, impl = reference, a cpu_baseline describing the run, e2e = the line's
    own value with zero transfer bytes.  (Also run with OMP_NUM_THREADS=1, as torchrun exports it: the arm must
    still take all host cores.)"""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-seconds", "0.5"], capture_output=True, text=True, env=env, timeout=280)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["unit"] == "claims/s" and rec["higher_is_better"] is True
    assert rec["metric"].startswith("claims/sec top-10 over 25M") and "workload" in rec["config"]
    assert rec["value"] > 0 and rec["n_gpus"] == 1 and rec["steps"] == 1
    cb = rec["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == rec["value"] and "rows" in cb["sample"]
    assert cb["cores"] == len(os.sched_getaffinity(0)) or cb["cores"] == os.cpu_count()
    assert rec["e2e"] == {"value": rec["value"], "unit": "claims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_workload_table_names_the_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.WORKLOADS["fever_sentences_25M"] == (10000, 25_000_000, 768, 10)      # the metric's config
    assert bench.WORKLOADS["fever_pages_5.4M"] == (10000, 5_400_000, 768, 10)          # configs[1]
    assert bench.WORKLOADS["large_batch_65k_x_5.4M_top100"] == (65536, 5_400_000, 768, 100)  # configs[4]
