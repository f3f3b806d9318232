#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, one rank per GPU, NCCL backend for the rendezvous): the sharded search --
per-rank scan, then ONE exchange step: the fused select + NVLink peer-memory exchange + merge kernel (k <= 16), the
query-sliced peer-memory exchange (k > 16), or NCCL all-gather + merge -- must equal the single-GPU search of the
whole corpus bit for bit, and the CPU oracle on a claim sample.  `--quick` skips the latency section.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/gpu_sharded_check.py
Launched by tests/test_robustness_gpu.py::test_sharded_search_under_torchrun_equals_single_gpu when the box has
>= 2 GPUs; a log of an 8-GPU run is kept under profiles/."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import drs_b200  # noqa: E402
from oracle import dense_topk  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
QUICK = "--quick" in sys.argv
ok = True
# (the k = 100 and k = 33 cases take the query-sliced exchange; 9 claims over 8 ranks leaves ragged / empty slices)
for nq, nc, dim, k in [(300, 100003, 128, 10), (1000, 1000000, 768, 10), (64, 7, 64, 5), (1000, 300000, 128, 100),
                       (9, 50000, 64, 33), (5000, 200000, 64, 256)]:
    g = torch.Generator(device=dev).manual_seed(1337)            # same data on every rank
    corpus = torch.nn.functional.normalize(torch.randn(nc, dim, generator=g, device=dev), dim=1).bfloat16()
    if nc > 100:
        corpus[nc - 1] = corpus[3]                               # a tie across the first and last shard
    queries = torch.nn.functional.normalize(torch.randn(nq, dim, generator=g, device=dev), dim=1).bfloat16()
    queries[0] = corpus[min(3, nc - 1)]
    lo, hi = drs_b200.shard_bounds(nc, rank, world)
    index = drs_b200.ShardedDenseIndex(corpus[lo:hi].clone(), nc, device=dev)                      # auto: fused p2p
    index_nccl = drs_b200.ShardedDenseIndex(corpus[lo:hi].clone(), nc, device=dev, exchange="nccl")
    fs, fi = drs_b200.search(queries, corpus, k)
    same = True
    for rep in range(3):                                                                            # epochs / both buffer parities
        s, i = index.search(queries, k)
        same = same and torch.equal(i, fi) and torch.equal(s, fs)
    s2, i2 = index_nccl.search(queries, k)
    same = same and torch.equal(i2, fi) and torch.equal(s2, fs)
    if nq > 64:                                                                                     # a smaller batch on the same buffers
        s3, i3 = index.search(queries[:17], k)
        same = same and torch.equal(i3, fi[:17]) and torch.equal(s3, fs[:17])
    if nq == 1000 and k == 100:                                                                     # buffer regrowth between calls
        big = torch.cat([queries] * 12)
        s4, i4 = index.search(big, k)
        same = same and torch.equal(i4[-nq:], fi) and torch.equal(s4[:nq], fs)
        s5, i5 = index.search(queries, k)
        same = same and torch.equal(i5, fi) and torch.equal(s5, fs)
    rv, ri = dense_topk.search(queries[:32].cpu(), corpus.cpu(), k)
    oracle_ok = torch.equal(i[:32].cpu()[:, 0], ri[:, 0]) and torch.allclose(s[:32].cpu(), rv, rtol=2e-2, atol=1e-4)
    tie_ok = nc <= 100 or i[0, :2].tolist() == [3, nc - 1]
    flag = torch.tensor([int(same and oracle_ok and tie_ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"sharded x{world} [{index.exchange}] nq={nq} nc={nc} dim={dim} k={k}: equal_to_single_gpu={same} oracle={oracle_ok} "
              f"tie={tie_ok} all_ranks={bool(flag.item())}", flush=True)
    ok = ok and bool(flag.item())

# the fused exchange keeps its epoch on the device: a captured CUDA graph of the sharded search replays correctly
nc, dim, k = 200_000, 128, 10
g = torch.Generator(device=dev).manual_seed(99)
corpus = torch.nn.functional.normalize(torch.randn(nc, dim, generator=g, device=dev), dim=1).bfloat16()
lo, hi = drs_b200.shard_bounds(nc, rank, world)
idx = drs_b200.ShardedDenseIndex(corpus[lo:hi].clone(), nc, device=dev)
static_q = corpus[:16].clone()
for _ in range(2):
    idx.search(static_q, k)
torch.cuda.synchronize()
dist.barrier()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    gs, gi = idx.search(static_q, k)
graph_ok = idx.exchange == "p2p"
for rep in range(4):
    static_q.copy_(corpus[1000 * rep + 7: 1000 * rep + 23])
    graph.replay()
    torch.cuda.synchronize()
    graph_ok = graph_ok and gi[:, 0].tolist() == list(range(1000 * rep + 7, 1000 * rep + 23))
es, ei = idx.search(static_q, k)                                   # eager calls interleave with replays
graph_ok = graph_ok and torch.equal(ei, gi) and torch.equal(es, gs)
flag = torch.tensor([int(graph_ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"sharded x{world} [{idx.exchange}] CUDA graph replay of the fused exchange: {bool(flag.item())}", flush=True)
ok = ok and bool(flag.item())

if QUICK:
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)

# top-100 exchange at a BASELINE configs[4]-like shape: 65 536 claims, sliced peer-memory kernel vs NCCL all-gather + merge
nc, dim, k = 100_000 * world, 128, 100
g = torch.Generator(device=dev).manual_seed(1337 + rank)
lo, hi = drs_b200.shard_bounds(nc, rank, world)
shard = torch.nn.functional.normalize(torch.randn(hi - lo, dim, generator=g, device=dev), dim=1).bfloat16()
q = torch.nn.functional.normalize(torch.randn(65536, dim, generator=torch.Generator(device=dev).manual_seed(5), device=dev), dim=1).bfloat16()
res = {}
for mode in ("p2p", "nccl"):
    idx = drs_b200.ShardedDenseIndex(shard, nc, device=dev, exchange=mode)
    for _ in range(2):
        out = idx.search(q, k)
    timing = {}
    dist.barrier()
    torch.cuda.synchronize()
    for _ in range(5):
        out = idx.search(q, k, timing=timing)
    torch.cuda.synchronize()
    loc = sum(a.elapsed_time(b) for a, b in timing["local"]) / 5
    ex = sum(a.elapsed_time(b) for a, b in timing["exchange"]) / 5
    t = torch.tensor([loc, ex], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[mode] = (t.tolist(), out)
eq = torch.equal(res["p2p"][1][0], res["nccl"][1][0]) and torch.equal(res["p2p"][1][1], res["nccl"][1][1])
ok = ok and eq
if rank == 0:
    print(f"top-100 exchange x{world}, 65536 claims: sliced p2p {res['p2p'][0][1]:.3f} ms vs nccl all-gather + merge {res['nccl'][0][1]:.3f} ms "
          f"(local scan + select {res['p2p'][0][0]:.3f} ms), same={eq}", flush=True)

# latency of the exchange in the small-batch regime: fused p2p kernel vs select + 2 all-gathers + merge
nc, dim, k = 4_000_000, 768, 10
g = torch.Generator(device=dev).manual_seed(1337 + rank)
lo, hi = drs_b200.shard_bounds(nc, rank, world)
shard = torch.nn.functional.normalize(torch.randn(hi - lo, dim, generator=g, device=dev), dim=1).bfloat16()
for nq in (16, 128, 10000):
    q = torch.nn.functional.normalize(torch.randn(nq, dim, generator=torch.Generator(device=dev).manual_seed(5), device=dev), dim=1).bfloat16()
    res = {}
    for mode in ("p2p", "nccl"):
        idx = drs_b200.ShardedDenseIndex(shard, nc, device=dev, exchange=mode)
        for _ in range(3):
            out = idx.search(q, k)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            out = idx.search(q, k)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[mode] = (t.item(), out)
    eq = torch.equal(res["p2p"][1][0], res["nccl"][1][0]) and torch.equal(res["p2p"][1][1], res["nccl"][1][1])
    ok = ok and eq
    if rank == 0:
        print(f"exchange latency x{world} nq={nq} rows/rank={hi - lo}: p2p {res['p2p'][0]:.3f} ms/search, nccl {res['nccl'][0]:.3f} ms/search, same={eq}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
