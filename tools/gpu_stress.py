#!/usr/bin/env python
"""Randomised stress (test tooling): many search shapes where claims' units run over several rounds (threshold seeds,
round barrier, A super-blocks, adaptive passes) and many InfoNCE shapes (symmetric H, transposed reads), each compared
with an exact reference computed on the GPU in fp32 / fp64.  Prints one line per failure and a summary.
    python tools/gpu_stress.py [search cases=120] [loss cases=40] [seed=0]      (DRS_OPTIONS=name=value,... sets engine
options, e.g. tune.symmetric_lse=2 puts every whole-tile batch on the symmetric forward)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import drs_b200 as drs  # noqa: E402

dev = torch.device("cuda:0")
n_search = int(sys.argv[1]) if len(sys.argv) > 1 else 120
n_loss = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rng = np.random.RandomState(int(sys.argv[3]) if len(sys.argv) > 3 else 0)
unit = lambda x: torch.nn.functional.normalize(x, dim=1)
for opt in os.environ.get("DRS_OPTIONS", "").split(","):
    if "=" in opt:
        drs.set_option(opt.split("=")[0], int(opt.split("=")[1]))


def exact_topk(q, c, k):
    """fp32 matmul of the same values, (score desc, id asc) order, plus the (k+1)-th score for gap tests"""
    best_s = torch.empty(q.shape[0], 0, device=dev)
    best_i = torch.empty(q.shape[0], 0, dtype=torch.int64, device=dev)
    qf = q.float()
    for r0 in range(0, c.shape[0], 1 << 18):
        sc = qf @ c[r0:r0 + (1 << 18)].float().T
        kk = min(k + 9, sc.shape[1])
        s, i = torch.topk(sc, kk, dim=1)
        s, i = torch.cat([best_s, s], 1), torch.cat([best_i, i + r0], 1)
        o = torch.argsort(i, dim=1, stable=True)
        s, i = torch.gather(s, 1, o), torch.gather(i, 1, o)
        o = torch.argsort(s, dim=1, descending=True, stable=True)
        best_s, best_i = torch.gather(s, 1, o)[:, :k + 9], torch.gather(i, 1, o)[:, :k + 9]
    return best_s, best_i


fails = 0
for case in range(n_search):
    dtype = torch.float32 if case % 4 == 3 else torch.bfloat16
    nq = int(rng.choice([130, 300, 600, 1100, 2500, 5000]))
    nc = int(rng.choice([30_000, 120_000, 400_000, 1_000_003]))
    dim = int(rng.choice([64, 128, 256, 768]))
    k = int(rng.choice([1, 5, 10, 16, 33, 100]))
    g = torch.Generator(device=dev).manual_seed(1000 + case)
    c = unit(torch.randn(nc, dim, generator=g, device=dev))
    j = torch.randint(0, nc, (nq,), generator=g, device=dev)
    q = unit(c[j] + float(rng.choice([0.05, 0.3, 1.0])) * torch.randn(nq, dim, generator=g, device=dev))
    if case % 3 == 0:                       # exact duplicates spread over the corpus: ties across units and rounds
        src = torch.randint(0, nc, (64,), generator=g, device=dev)
        dst = torch.randint(0, nc, (64,), generator=g, device=dev)
        c[dst] = c[src]
    q, c = q.to(dtype), c.to(dtype)
    s, i = drs.search(q, c, k)
    rs, ri = exact_topk(q, c, k)
    tol, gap = (1e-5, 5e-6) if dtype == torch.float32 else (2e-2, 1e-4)
    kk = s.shape[1]
    ok_s = bool(((s - rs[:, :kk]).abs() <= tol * rs[:, :kk].abs() + tol * 0.1).all())
    d_prev = torch.cat([torch.full((nq, 1), 1e30, device=dev), rs[:, :kk - 1] - rs[:, 1:kk]], 1) if kk > 1 else torch.full((nq, 1), 1e30, device=dev)
    d_next = rs[:, :kk] - rs[:, 1:kk + 1]
    strict = (d_prev > gap) & (d_next > gap)
    ok_i = bool(torch.equal(i[strict], ri[:, :kk][strict]))
    # ties: equal scores must come in ascending id order
    same = s[:, :-1] == s[:, 1:]
    ok_t = bool((i[:, :-1][same] < i[:, 1:][same]).all()) if kk > 1 else True
    ok_d = bool((s[:, :-1] >= s[:, 1:]).all()) if kk > 1 else True
    if not (ok_s and ok_i and ok_t and ok_d):
        fails += 1
        print(f"SEARCH FAIL case {case}: dtype={dtype} nq={nq} nc={nc} dim={dim} k={k} scores={ok_s} ids={ok_i} ties={ok_t} sorted={ok_d}", flush=True)
print(f"search: {n_search - fails} / {n_search} ok", flush=True)

lfails = 0
for case in range(n_loss):
    n = int(rng.choice([4, 64, 128, 256, 384, 512, 1024, 1536, 2048, 4096]))   # (4096 x 768: the symmetric forward by default)
    dim = int(rng.choice([64, 128, 256, 768]))      # (narrower rows: bf16 rounding alone moves gradient rows by ~3 %, 'auto' takes fp32)
    klen = int(rng.choice([0, 0, 64, 512, 2048, 4104, 12544]))     # long queues: dq = Hq x queue runs split-K
    temp = float(rng.choice([0.05, 0.07, 0.2, 0.02]))   # (0.02: logits beyond the bound of the symmetric forward)
    g = torch.Generator(device=dev).manual_seed(5000 + case)
    q = unit(torch.randn(n, dim, generator=g, device=dev)).requires_grad_(True)
    kk = unit(torch.randn(n, dim, generator=g, device=dev) * 0.5 + q.detach()).requires_grad_(True)
    queue = torch.nn.functional.normalize(torch.randn(dim, klen, generator=g, device=dev), dim=0) if klen else None
    prec = "fp32" if rng.rand() < 0.25 and n <= 1024 else "bf16"
    loss = drs.NCELoss({"temperature": temp, "precision": prec})(q, kk, queue)
    loss.backward()
    # closed form in float64 on the GPU (contrastive_loss.py:56-93)
    # (T = 0.02 on the bf16 path: the rounding of the embeddings alone moves single gradient rows by more than the 3 % bar
    #  -- logits scale with 1/T -- so those cases are compared on the bf16-rounded embeddings the kernels multiply)
    rounded = prec == "bf16" and temp < 0.05
    qd = (q.detach().bfloat16() if rounded else q.detach()).double().requires_grad_(True)
    kd = (kk.detach().bfloat16() if rounded else kk.detach()).double().requires_grad_(True)
    if rounded and queue is not None:
        queue = queue.bfloat16().float()
    f = torch.cat([qd, kd])
    sm = (f @ f.T) / temp
    sm = sm.masked_fill(torch.eye(2 * n, dtype=torch.bool, device=dev), float("-inf"))
    logits = sm if queue is None else torch.cat([sm, torch.cat([qd @ queue.double(), qd @ queue.double()]) / temp], 1)
    pos = torch.arange(2 * n, device=dev).roll(n)
    ref = (torch.logsumexp(logits, 1) - sm[torch.arange(2 * n, device=dev), pos]).sum() / 2
    ref.backward()
    ltol, gtol = (2e-2, 3e-2) if prec == "bf16" else (1e-5, 2e-4)
    ok_l = abs(loss.item() - ref.item()) <= ltol * abs(ref.item()) + 1e-4
    def rows_ok(a, b):
        rn = b.norm(dim=1)
        floor = rn.max().clamp_min(1e-30) * 1e-6
        return bool(((a.double() - b).norm(dim=1) / rn.clamp_min(floor)).max() <= gtol)
    ok_g = rows_ok(q.grad, qd.grad) and rows_ok(kk.grad, kd.grad)
    if not (ok_l and ok_g):
        lfails += 1
        print(f"LOSS FAIL case {case}: {prec} n={n} dim={dim} queue={klen} T={temp} loss={ok_l} ({loss.item()} vs {ref.item()}) grads={ok_g}", flush=True)
print(f"loss: {n_loss - lfails} / {n_loss} ok", flush=True)
sys.exit(1 if fails or lfails else 0)
