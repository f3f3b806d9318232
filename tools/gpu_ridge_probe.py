#!/usr/bin/env python
"""Perf probe (test tooling): the ridge between the two roofs (129..1024 claims per pass) -- CTA pair vs single CTA."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import drs_b200 as drs  # noqa: E402


def scan_ms(q, c, k, iters=10):
    prof = []
    for _ in range(3):
        drs.search(q, c, k, profile=prof)
    prof.clear()
    for _ in range(iters):
        drs.search(q, c, k, profile=prof)
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in prof) / len(prof)


def main():
    dev = torch.device("cuda:0")
    nc = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
    g = torch.Generator(device=dev).manual_seed(1)
    c = torch.empty(nc, 768, dtype=torch.bfloat16, device=dev)
    for r0 in range(0, nc, 1 << 20):
        r1 = min(nc, r0 + (1 << 20))
        c[r0:r1] = torch.nn.functional.normalize(torch.randn(r1 - r0, 768, generator=g, device=dev), dim=1)
    qa = torch.nn.functional.normalize(torch.randn(1024, 768, generator=g, device=dev), dim=1).bfloat16()
    t_hbm = nc * 768 * 2 / 6.551e12 * 1e3
    for nq in (64, 128, 192, 256):
        q = qa[:nq].contiguous()
        t_mma = 2.0 * nq * nc * 768 / 1389.5e12 * 1e3
        line = f"nq={nq:5d} roof {max(t_hbm, t_mma):6.3f} ms |"
        for cg in (1, 2):
            drs.set_option("search.cta_group", cg)
            ms = scan_ms(q, c, 10)
            line += f" cg{cg}: {ms:6.3f} ms ({max(t_hbm, t_mma) / ms * 100:5.1f} %)"
        drs.set_option("search.cta_group", 0)
        print(line, flush=True)


if __name__ == "__main__":
    main()
