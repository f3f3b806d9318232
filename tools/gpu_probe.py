#!/usr/bin/env python
"""Perf probe (test tooling): time the bf16 search kernel with parts of the pipeline disabled
(debug.flags) to see which stage bounds a tile.  python tests/gpu_probe.py [nq nc cg]"""
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import drs_b200 as drs  # noqa: E402


def timed(q, c, k, iters):
    drs.search(q, c, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        drs.search(q, c, k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def clocks_sampler(stop, out):
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active",
                                "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
            out.append(r.stdout.strip().split("\n")[0])
        except Exception:  # noqa: BLE001
            pass
        time.sleep(0.2)


def main():
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    nc = int(sys.argv[2]) if len(sys.argv) > 2 else 4000000
    cgs = [int(sys.argv[3])] if len(sys.argv) > 3 else [2, 1]
    dim, k = 768, 10
    g = torch.Generator(device="cuda").manual_seed(1337)
    c = torch.randn(nc, dim, generator=g, device="cuda", dtype=torch.bfloat16)
    q = torch.randn(nq, dim, generator=g, device="cuda", dtype=torch.bfloat16)
    res = {}
    for cg in cgs:
        drs.set_option("search.cta_group", cg)
        for flags, label in [(0, "full"), (1, "no_functor"), (3, "no_tmem_ld"), (7, "tma_only"), (4, "no_mma_full_epi")]:
            drs.set_option("debug.flags", flags)
            stop, samples = threading.Event(), []
            th = threading.Thread(target=clocks_sampler, args=(stop, samples))
            th.start()
            ms = timed(q, c, k, 5)
            stop.set()
            th.join()
            res[f"cg{cg}_{label}"] = dict(ms=round(ms, 3), tflops=round(2.0 * nq * nc * dim / ms / 1e9, 1),
                                          clocks=samples[-3:])
            print(f"cg{cg} {label:16s} {ms:9.3f} ms  {2.0 * nq * nc * dim / ms / 1e9:8.1f} TFLOP/s  {samples[-2:]}", flush=True)
        drs.set_option("debug.flags", 0)
    # cuBLAS reference point: same flops as a plain bf16 GEMM (writes the nq x chunk score block)
    chunk = 262144
    cc = c[:chunk]
    out = torch.empty(nq, chunk, device="cuda", dtype=torch.bfloat16)
    torch.matmul(q, cc.T, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        torch.matmul(q, cc.T, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"cublas {nq}x{chunk}x{dim}: {ms:.3f} ms {2.0 * nq * chunk * dim / ms / 1e9:.1f} TFLOP/s")
    res["cublas"] = dict(ms=ms, tflops=2.0 * nq * chunk * dim / ms / 1e9)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
