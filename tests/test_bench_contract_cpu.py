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


def test_post_timing_parity_checker_accepts_the_truth_and_catches_corruptions():
    """bench.py verifies the results of its own timed steps (the `parity` object of the JSON line).  The checker
    itself, on the CPU with world = 1: exact results pass; a wrong id, a score off by 5 %, an unsorted list and a
    planted claim that does not come back first are each caught."""
    import torch
    sys.path.insert(0, ROOT)
    import bench
    from oracle import dense_topk
    g = torch.Generator().manual_seed(7)
    nc, nq, dim, k, n_pl = 4000, 200, 64, 10, 16
    corpus = torch.nn.functional.normalize(torch.randn(nc, dim, generator=g), dim=1).bfloat16()
    queries = torch.nn.functional.normalize(torch.randn(nq, dim, generator=g), dim=1).bfloat16()
    planted_rows = [((2 * p + 1) * nc) // (2 * n_pl) for p in range(n_pl)]
    planted_claims = list(range(nq - n_pl, nq))
    noise = torch.randn(n_pl, dim, generator=g)
    queries[nq - n_pl:] = torch.nn.functional.normalize(corpus[planted_rows].float() + 0.05 * noise, dim=1).bfloat16()
    s, i = dense_topk.search(queries, corpus, k)
    dev = torch.device("cpu")

    def check(gs, gi):
        return bench.parity_check(torch, None, 1, dev, corpus, 0, queries, k, gs, gi, planted_claims, planted_rows)

    good = check(s, i)
    assert good["ok"] and good["checked"] == 64 and good["planted"] == 16 and good["ids_mismatched"] == 0
    assert good["ids_compared"] > 300 and good["max_rel_err"] < 1e-5
    bad_i = i.clone()
    bad_i[0, 0] = (bad_i[0, 0] + 1) % nc                       # claim 0 is always in the sample
    assert not check(s, bad_i)["ok"]
    bad_s = s.clone()
    bad_s[0, 3] *= 1.05
    r = check(bad_s, i)
    assert not r["ok"] and r["max_rel_err"] > 2e-2
    unsorted = s.clone()
    unsorted[5, [0, 1]] = unsorted[5, [1, 0]]
    assert not check(unsorted, i)["sorted_desc"]
    moved = i.clone()
    moved[planted_claims[3], [0, 1]] = moved[planted_claims[3], [1, 0]]
    r = check(s, moved)
    assert not r["planted_top1_ok"] and not r["ok"]
