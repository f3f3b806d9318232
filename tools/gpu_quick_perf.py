#!/usr/bin/env python
"""Perf probe (test tooling): the four tensor-bound shapes in ~10 s -- 10k x 5.4M top-10, 65 536 x 675k top-100 (the
per-GPU share of BASELINE configs[4]), NCELoss fwd+bwd at 4096 x 768, and 1k x 100k fp32 top-5.
    python tools/gpu_quick_perf.py [what=all|search|cfg5|infonce|fp32]   (DRS_OPTIONS=name=value,... sets engine options)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import drs_b200 as drs  # noqa: E402

PEAK = 1389.5e12
dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "all"
for name in os.environ.get("DRS_OPTIONS", "").split(","):
    if "=" in name:
        drs.set_option(name.split("=")[0], int(name.split("=")[1]))
unit = lambda x: torch.nn.functional.normalize(x, dim=1)


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


g = torch.Generator(device=dev).manual_seed(1337)
if what in ("all", "search", "cfg5"):
    nc = 5_400_000
    c = torch.empty(nc, 768, dtype=torch.bfloat16, device=dev)
    for r0 in range(0, nc, 1 << 20):
        r1 = min(nc, r0 + (1 << 20))
        c[r0:r1] = unit(torch.randn(r1 - r0, 768, generator=g, device=dev))
    if what in ("all", "search"):
        q = unit(torch.randn(10000, 768, generator=g, device=dev)).bfloat16()
        ms = timed(lambda: drs.search(q, c, 10), 8)
        fl = 2.0 * 10000 * nc * 768
        print(f"10000 x {nc} top-10: {ms:.3f} ms, {fl / ms / 1e9:.1f} TFLOP/s, mma_frac {fl / ms / 1e-3 / PEAK:.4f}", flush=True)
    q = unit(torch.randn(65536, 768, generator=g, device=dev)).bfloat16()
    cc = c[:675_000]
    ms = timed(lambda: drs.search(q, cc, 100), 5)
    fl = 2.0 * 65536 * 675_000 * 768
    print(f"65536 x 675000 top-100: {ms:.3f} ms, {fl / ms / 1e9:.1f} TFLOP/s, mma_frac {fl / ms / 1e-3 / PEAK:.4f}, "
          f"open claims per pass {sys.modules[drs.__name__ + '.retrieval'].open_claims_per_pass()}", flush=True)
    ms = timed(lambda: drs.search(q, cc, 10), 5)
    print(f"65536 x 675000 top-10: {ms:.3f} ms, {fl / ms / 1e9:.1f} TFLOP/s, mma_frac {fl / ms / 1e-3 / PEAK:.4f}", flush=True)
    del c, cc, q
if what in ("all", "infonce"):
    n, dim = 4096, 768
    q = unit(torch.randn(n, dim, generator=g, device=dev)).requires_grad_(True)
    k = unit(torch.randn(n, dim, generator=g, device=dev) * 0.5 + q.detach()).requires_grad_(True)
    crit = drs.NCELoss({"temperature": 0.05})

    def step():
        q.grad = None
        k.grad = None
        crit(q, k, None).backward()

    best = min(timed(step, 20) for _ in range(3))
    fl = 3 * 2.0 * (2 * n) ** 2 * dim
    print(f"NCELoss fwd+bwd {n} x {dim}: {best:.4f} ms, {fl / best / 1e9:.1f} TFLOP/s algorithmic, mma_frac {fl / best / 1e-3 / PEAK:.4f}", flush=True)
if what in ("all", "fp32"):
    c = unit(torch.randn(100_000, 768, generator=g, device=dev))
    q = unit(torch.randn(1000, 768, generator=g, device=dev))
    for mode in (0, 1):
        drs.set_option("search.fp32_mode", mode)
        ms = timed(lambda: drs.search(q, c, 5), 10)
        print(f"1000 x 100000 x 768 fp32 top-5, fp32_mode={mode}: {ms:.3f} ms, {2.0 * 1000 * 100000 * 768 / ms / 1e9:.1f} TFLOP/s effective", flush=True)
    drs.set_option("search.fp32_mode", 0)
