#!/usr/bin/env python
"""Profiling target (test tooling): ONE configuration of the search per process, for ncu.
    python tests/gpu_profile_target.py hbm   [B=16]  [rows=8000000]   small-batch, HBM-bound scan (CG=1)
    python tests/gpu_profile_target.py cfg5  [k=100]                  65 536 claims x 675 000 rows (configs[4] per-GPU share)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import drs_b200 as drs  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "hbm"
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1337)
    if mode == "hbm":
        b = int(sys.argv[2]) if len(sys.argv) > 2 else 16
        nc = int(sys.argv[3]) if len(sys.argv) > 3 else 8_000_000
        nq, k = b, 10
    else:
        nq, nc, k = 65536, 675000, int(sys.argv[2]) if len(sys.argv) > 2 else 100
    c = torch.empty(nc, 768, dtype=torch.bfloat16, device=dev)
    for r0 in range(0, nc, 1 << 20):
        r1 = min(nc, r0 + (1 << 20))
        c[r0:r1] = torch.nn.functional.normalize(torch.randn(r1 - r0, 768, generator=g, device=dev), dim=1)
    q = torch.nn.functional.normalize(torch.randn(nq, 768, generator=g, device=dev), dim=1).bfloat16()
    for name in os.environ.get("DRS_OPTIONS", "").split(","):          # e.g. DRS_OPTIONS=tune.round_barrier=0
        if "=" in name:
            drs.set_option(name.split("=")[0], int(name.split("=")[1]))
    for _ in range(3):
        s, i = drs.search(q, c, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        drs.search(q, c, k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{mode}: {nq} x {nc} top-{k}: {ms:.3f} ms/search, {nc * 768 * 2 / ms / 1e6:.1f} GB/s corpus stream, "
          f"{2.0 * nq * nc * 768 / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
