"""Robustness of the kernels that wait on grid-wide flags, NaN / exact-zero edge cases, and the multi-GPU exchange
under torchrun.  GPU only.  Anything that could leave a trapped CUDA context behind runs in a subprocess."""
import os
import socket
import subprocess
import sys

import pytest
import torch

import drs_b200
from oracle import dense_topk

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda:0"

_TWO_STREAMS = r"""
import sys, torch
sys.path.insert(0, %r)
import drs_b200
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(7)
unit = lambda x: torch.nn.functional.normalize(x, dim=1)
corpus = unit(torch.randn(1_500_000, 256, generator=g, device=dev)).bfloat16()
qa = unit(torch.randn(600, 256, generator=g, device=dev)).bfloat16()      # 3 claim tiles: a multi-round scan (round barrier on)
qb = unit(torch.randn(700, 256, generator=g, device=dev)).bfloat16()
ra, rb = drs_b200.search(qa, corpus, 10), drs_b200.search(qb, corpus, 10)
torch.cuda.synchronize()
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
ok = True
for it in range(12):
    with torch.cuda.stream(sa):
        xa = drs_b200.search(qa, corpus, 10)
    with torch.cuda.stream(sb):
        xb = drs_b200.search(qb, corpus, 10)
    torch.cuda.synchronize()
    ok = ok and torch.equal(xa[1], ra[1]) and torch.equal(xb[1], rb[1]) and torch.equal(xa[0], ra[0]) and torch.equal(xb[0], rb[0])
print("two_streams_ok", ok, "hang", drs_b200._lib.hang_report()["flag"], "fallbacks", drs_b200.get_option("debug.coop_fallbacks"))
sys.exit(0 if ok else 1)
"""


def test_two_barrier_scans_on_two_streams_do_not_deadlock():
    """Two multi-round scans issued on two streams at once.  Each spins on its own grid-wide round counter; launched
    plainly, the two grids can each get half of the SMs and wait for the other half forever (until the watchdog traps
    and poisons the context).  Launched cooperatively the driver places one whole grid at a time.  Results must equal
    the serial searches, with no hang report."""
    out = subprocess.run([sys.executable, "-c", _TWO_STREAMS % ROOT], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "two_streams_ok True hang 0" in out.stdout, out.stdout[-2000:]


def test_exact_zero_scores_keep_the_lower_index_rule_with_threshold_seeds():
    """Rows scoring exactly 0.0 (zero-padded corpus rows, orthogonal rows) must still break ties to the lower index
    when later units start from a seeded floor: the floor below +0.0 must not decode to -0.0 (which `>` cannot tell
    from +0.0).  Claims have non-negative scores only against 3 rows; everything else is an exact 0 tie."""
    dim, nc, nq, k = 64, 400_000, 600, 10
    corpus = torch.zeros(nc, dim, device=DEV)
    corpus[:, 0] = 0.0
    hot = torch.tensor([5, 200_000, 399_999], device=DEV)
    corpus[hot, 1] = torch.tensor([0.5, 0.25, 0.125], device=DEV)
    corpus[:, 2] = 1.0                                     # orthogonal to the claims: score contribution 0
    q = torch.zeros(nq, dim, device=DEV)
    q[:, 1] = 1.0
    for dtype in (torch.bfloat16, torch.float32):
        s, i = drs_b200.search(q.to(dtype), corpus.to(dtype), k)
        want_i = torch.tensor([5, 200_000, 399_999, 0, 1, 2, 3, 4, 6, 7], device=DEV)
        want_s = torch.tensor([0.5, 0.25, 0.125] + [0.0] * 7, device=DEV)
        assert torch.equal(i, want_i.expand(nq, k)), (dtype, i[0].tolist())
        assert torch.equal(s, want_s.expand(nq, k))


def test_nan_scores_are_never_selected():
    """A NaN row ranks below every number (numpy argsort order, tfidf_doc_ranker.py:70-71) in search and re-rank."""
    g = torch.Generator(device=DEV).manual_seed(3)
    c = torch.nn.functional.normalize(torch.randn(5000, 64, generator=g, device=DEV), dim=1)
    q = torch.nn.functional.normalize(torch.randn(40, 64, generator=g, device=DEV), dim=1)
    c[17] = float("nan")
    c[4000, 3] = float("nan")
    clean = torch.ones(5000, dtype=torch.bool, device=DEV)
    clean[[17, 4000]] = False
    keep = torch.nonzero(clean).squeeze(1)
    for dtype in (torch.float32, torch.bfloat16):
        s, i = drs_b200.search(q.to(dtype), c.to(dtype), 10)
        rs, ri = drs_b200.search(q.to(dtype), c.to(dtype)[keep], 10)
        assert torch.equal(i, keep[ri]) and torch.equal(s, rs) and not torch.isnan(s).any()
    cand = torch.arange(0, 100, device=DEV).expand(40, 100).contiguous()
    s, i = drs_b200.rerank(q, c, cand, 15)
    assert not (i == 17).any() and not torch.isnan(s).any()
    allnan = torch.full((3, 64), float("nan"), device=DEV)
    s, i = drs_b200.search(allnan, c, 5)
    assert (i == -1).all() and torch.isinf(s).all()


def test_sharded_index_row_ids_round_trip():
    """get_doc_index / get_doc_id speak GLOBAL rows on a shard (id_base > 0), the ids `search` returns."""
    g = torch.Generator(device=DEV).manual_seed(5)
    emb = torch.nn.functional.normalize(torch.randn(300, 64, generator=g, device=DEV), dim=1)
    ids = [f"doc_{r}" for r in range(1000, 1300)]
    shard = drs_b200.DenseIndex(emb, ids, device=DEV, id_base=1000)
    s, i = shard.search(emb[42:43], 1)
    assert int(i[0, 0]) == 1042 and shard.get_doc_id(1042) == "doc_1042" and shard.get_doc_index("doc_1042") == 1042
    assert shard.closest_docs(emb[7], 1)[0] == ["doc_1007"]
    with pytest.raises(IndexError):
        shard.get_doc_id(5)


def test_randomised_stress_of_multi_round_scans_and_losses():
    """tools/gpu_stress.py with a fixed seed: 60 search shapes whose claims' units run over several rounds (threshold
    seeds, round barrier, adaptive passes, planted duplicates across units) and 24 InfoNCE shapes (symmetric H, transposed
    reads, queues), each against an exact fp32 / fp64 computation on the GPU."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_stress.py"), "60", "24", "11"], capture_output=True,
                         text=True, timeout=800)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "search: 60 / 60 ok" in out.stdout and "loss: 24 / 24 ok" in out.stdout


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_search_under_torchrun_equals_single_gpu():
    """tests/gpu_sharded_check.py on 2 real GPUs (fused p2p exchange, query-sliced k = 100 exchange, NCCL path, batch
    size changes, buffer regrowth, CUDA-graph replay), each result compared bit for bit with the single-GPU search."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs on the box (the committed log profiles/r02_sharded_check_*.log is the multi-GPU record)")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "gpu_sharded_check.py"), "--quick"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=840)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "all_ranks=False" not in out.stdout and "equal_to_single_gpu=True" in out.stdout
