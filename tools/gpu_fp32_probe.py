#!/usr/bin/env python
"""Probe (test tooling): accuracy and speed of the fp32 search on the tensor cores (3 x TF32) vs the FFMA kernel,
against a float64 reference on the GPU."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import drs_b200 as drs  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1337)
unit = lambda x: torch.nn.functional.normalize(x, dim=1)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for nq, nc, dim, k in [(1000, 100_000, 768, 5), (1000, 100_000, 128, 5), (4096, 1_000_000, 768, 10)]:
    c = unit(torch.randn(nc, dim, generator=g, device=dev))
    j = torch.randint(0, nc, (nq,), generator=g, device=dev)
    q = unit(c[j] + 0.1 * torch.randn(nq, dim, generator=g, device=dev))
    q[: nq // 2] = unit(torch.randn(nq // 2, dim, generator=g, device=dev))
    ref = torch.empty(nq, k, dtype=torch.float64, device=dev)
    refi = torch.empty(nq, k, dtype=torch.int64, device=dev)
    for a in range(0, nq, 256):
        sc = q[a:a + 256].double() @ c.double().T
        v, i = torch.topk(sc, k, dim=1)
        ref[a:a + 256], refi[a:a + 256] = v, i
    for mode in (0, 1):
        drs.set_option("search.fp32_mode", mode)
        s, i = drs.search(q, c, k)
        ms = timed(lambda: drs.search(q, c, k))
        err = (s.double() - ref).abs()
        rel = (err / ref.abs().clamp_min(1e-3)).max().item()
        print(f"{nq} x {nc} x {dim} top-{k} fp32_mode={mode}: {ms:.3f} ms ({2.0 * nq * nc * dim / ms / 1e9:.1f} TFLOP/s effective), "
              f"max abs err {err.max().item():.3e}, max rel err {rel:.3e}, ids equal {torch.equal(i, refi)}", flush=True)
    drs.set_option("search.fp32_mode", 0)

# faiss-shaped k-means assignment: 200k points x 8192 centroids x 128, fp32 (src/contrastor/utils.py:64-67)
x = torch.randn(200_000, 128, generator=g, device=dev)
cen = torch.randn(8192, 128, generator=g, device=dev)
for mode in (0, 1):
    drs.set_option("search.fp32_mode", mode)
    d, i = drs.flat_l2_search(x, cen, 1)
    ms = timed(lambda: drs.flat_l2_search(x, cen, 1), 5)
    dref = ((x[:4096].double()[:, None, :] - cen.double()[None, i[:4096, 0], :].squeeze(0)) ** 2).sum(-1) if False else None
    full = torch.cdist(x[:2048].double(), cen.double()) ** 2
    rv, ri = full.min(dim=1)
    print(f"flat_l2 200000 x 8192 x 128 fp32_mode={mode}: {ms:.3f} ms, ids equal {torch.equal(i[:2048, 0], ri)}, "
          f"max rel dist err {((d[:2048, 0].double() - rv).abs() / rv).max().item():.3e}", flush=True)
drs.set_option("search.fp32_mode", 0)
