#!/usr/bin/env python
"""Perf probe (test tooling): the ridge between the two roofs -- claims per pass x CTA group.
    python tools/gpu_ridge_sweep.py [corpus rows] [csv of claim counts]
(An L2 prefetch of the next corpus tile by the producer -- cp.async.bulk.prefetch.tensor -- was tried here: it made
every point SLOWER, 2.66 -> 4.86 ms at 128 claims x 12M rows: the TMA unit, not HBM latency, is what the extra
requests queue behind.)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import drs_b200 as drs  # noqa: E402


def scan_ms(q, c, k, iters=6):
    prof = []
    for _ in range(2):
        drs.search(q, c, k, profile=prof)
    prof.clear()
    for _ in range(iters):
        drs.search(q, c, k, profile=prof)
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in prof) / len(prof)


def main():
    dev = torch.device("cuda:0")
    nc = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
    nqs = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "128,256,384,512,1024,2048").split(",")]
    g = torch.Generator(device=dev).manual_seed(1)
    c = torch.empty(nc, 768, dtype=torch.bfloat16, device=dev)
    for r0 in range(0, nc, 1 << 20):
        r1 = min(nc, r0 + (1 << 20))
        c[r0:r1] = torch.nn.functional.normalize(torch.randn(r1 - r0, 768, generator=g, device=dev), dim=1)
    qa = torch.nn.functional.normalize(torch.randn(max(nqs), 768, generator=g, device=dev), dim=1).bfloat16()
    t_hbm = nc * 768 * 2 / 6.551e12 * 1e3
    ref = {}
    for nq in nqs:
        q = qa[:nq].contiguous()
        t_mma = 2.0 * nq * nc * 768 / 1389.5e12 * 1e3
        roof = max(t_hbm, t_mma)
        line = f"nq={nq:5d} roof {roof:6.3f} ms ({'hbm' if t_hbm >= t_mma else 'mma'}) |"
        for cg in (1, 2):
            if cg == 1 and nq > 1024:
                continue
            drs.set_option("search.cta_group", cg)
            ms = scan_ms(q, c, 10)
            s, i = drs.search(q, c, 10)
            if nq not in ref:
                ref[nq] = (s.clone(), i.clone())
            same = torch.equal(i, ref[nq][1]) and torch.equal(s, ref[nq][0])
            line += f" cg{cg}: {ms:6.3f} ({roof / ms * 100:5.1f} %){'' if same else ' DIFF'}"
        drs.set_option("search.cta_group", 0)
        print(line, flush=True)


if __name__ == "__main__":
    main()
