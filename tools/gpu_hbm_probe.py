#!/usr/bin/env python
"""Perf probe (test tooling): the small-query-batch, HBM-bound regime.  Times the bf16 scan for
B = 1..128 claims with parts of the pipeline disabled (debug.flags) and with a zero-padded batch,
to find what keeps small batches below the B = 128 bandwidth.  python tests/gpu_hbm_probe.py [nc]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import drs_b200 as drs  # noqa: E402


def timed(q, c, k, iters=10):
    for _ in range(2):
        drs.search(q, c, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        drs.search(q, c, k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    nc = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
    dim, k = 768, 10
    g = torch.Generator(device="cuda").manual_seed(1337)
    c = torch.nn.functional.normalize(torch.randn(nc, dim, generator=g, device="cuda"), dim=1).bfloat16()
    q = torch.nn.functional.normalize(torch.randn(128, dim, generator=g, device="cuda"), dim=1).bfloat16()
    gb = nc * dim * 2 / 1e9
    res = {}

    def run(label, qq, **opts):
        for name, v in opts.items():
            drs.set_option(name.replace("__", "."), v)
        ms = timed(qq, c, k)
        for name in opts:
            drs.set_option(name.replace("__", "."), 0 if name != "tune__round_barrier" else 1)
        res[label] = dict(ms=round(ms, 3), gbs=round(gb / ms * 1e3, 1))
        print(f"{label:34s} {ms:8.3f} ms  {gb / ms * 1e3:8.1f} GB/s", flush=True)

    for b in (1, 16, 64, 128):
        run(f"B{b}_full", q[:b].contiguous())
    qz = q.clone()
    qz[1:] = 0
    run("B128_zero_padded(1 real row)", qz)
    for b in (1, 128):
        qq = q[:b].contiguous()
        run(f"B{b}_no_functor", qq, debug__flags=1)
        run(f"B{b}_no_tmem_ld", qq, debug__flags=3)
        run(f"B{b}_tma_only", qq, debug__flags=7)
        run(f"B{b}_evict_first", qq, tune__b_hint=1)
        run(f"B{b}_no_round_barrier", qq, tune__round_barrier=0)
    # device copy of the same bytes: the practical HBM ceiling on this box (read + write counted once each)
    dst = torch.empty_like(c)
    dst.copy_(c)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        dst.copy_(c)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res["copy_rw"] = dict(ms=round(ms, 3), gbs=round(2 * gb / ms * 1e3, 1))
    print(f"copy (read+write)                  {ms:8.3f} ms  {2 * gb / ms * 1e3:8.1f} GB/s")
    del dst
    s = torch.empty(1, device="cuda")
    cf = c.view(torch.int16)
    torch.sum(cf[: 1 << 20])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "hbm_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
