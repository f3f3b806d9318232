#!/usr/bin/env python
"""Profiling target (test tooling): ONE configuration of the search per process, for ncu.
    python tools/gpu_profile_target.py hbm      [B=16]  [rows=8000000]   small-batch, HBM-bound scan (CG=1)
    python tools/gpu_profile_target.py cfg5     [k=100]                  65 536 claims x 675 000 rows (configs[4] per-GPU share)
    python tools/gpu_profile_target.py headline                          10 000 claims x 25M rows, top-10 (the bench workload)
    python tools/gpu_profile_target.py fp32                              1 000 claims x 100 000 x 768 fp32, top-5 (3 x TF32)
Under ncu the engine leaves the cooperative-launch attribute out by itself (it sees the injection environment).
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
    elif mode == "headline":
        nq, nc, k = 10000, 25_000_000, 10
    elif mode == "fp32":
        nq, nc, k = 1000, 100_000, 5
    else:
        nq, nc, k = 65536, 675000, int(sys.argv[2]) if len(sys.argv) > 2 else 100
    c = torch.empty(nc, 768, dtype=torch.bfloat16, device=dev)
    for r0 in range(0, nc, 1 << 20):
        r1 = min(nc, r0 + (1 << 20))
        c[r0:r1] = torch.nn.functional.normalize(torch.randn(r1 - r0, 768, generator=g, device=dev), dim=1)
    q = torch.nn.functional.normalize(torch.randn(nq, 768, generator=g, device=dev), dim=1).bfloat16()
    if mode == "fp32":
        c, q = c.float(), q.float()
    for name in os.environ.get("DRS_OPTIONS", "").split(","):          # e.g. DRS_OPTIONS=tune.round_barrier=0
        if "=" in name:
            drs.set_option(name.split("=")[0], int(name.split("=")[1]))
    reps = 2 if mode == "headline" else 5
    for _ in range(2 if mode == "headline" else 3):
        s, i = drs.search(q, c, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        drs.search(q, c, k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{mode}: {nq} x {nc} top-{k}: {ms:.3f} ms/search, {nc * 768 * 2 / ms / 1e6:.1f} GB/s corpus stream, "
          f"{2.0 * nq * nc * 768 / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
