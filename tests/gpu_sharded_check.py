#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, one rank per GPU, NCCL): the sharded search
(per-rank fused kernel -> all-gather over NVLink -> on-GPU merge) must equal the single-GPU
search of the whole corpus bit for bit, and the CPU oracle on a claim sample.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/gpu_sharded_check.py"""
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
ok = True
for nq, nc, dim, k in [(300, 100003, 128, 10), (1000, 1000000, 768, 10), (64, 7, 64, 5)]:
    g = torch.Generator(device=dev).manual_seed(1337)            # same data on every rank
    corpus = torch.nn.functional.normalize(torch.randn(nc, dim, generator=g, device=dev), dim=1).bfloat16()
    if nc > 100:
        corpus[nc - 1] = corpus[3]                               # a tie across the first and last shard
    queries = torch.nn.functional.normalize(torch.randn(nq, dim, generator=g, device=dev), dim=1).bfloat16()
    queries[0] = corpus[min(3, nc - 1)]
    lo, hi = drs_b200.shard_bounds(nc, rank, world)
    index = drs_b200.ShardedDenseIndex(corpus[lo:hi].clone(), nc, device=dev)
    s, i = index.search(queries, k)
    fs, fi = drs_b200.search(queries, corpus, k)
    same = torch.equal(i, fi) and torch.equal(s, fs)
    rv, ri = dense_topk.search(queries[:32].cpu(), corpus.cpu(), k)
    oracle_ok = torch.equal(i[:32].cpu()[:, 0], ri[:, 0]) and torch.allclose(s[:32].cpu(), rv, rtol=2e-2, atol=1e-4)
    tie_ok = nc <= 100 or i[0, :2].tolist() == [3, nc - 1]
    flag = torch.tensor([int(same and oracle_ok and tie_ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"sharded x{world} nq={nq} nc={nc} dim={dim} k={k}: equal_to_single_gpu={same} oracle={oracle_ok} "
              f"tie={tie_ok} all_ranks={bool(flag.item())}", flush=True)
    ok = ok and bool(flag.item())
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
