"""Perf probe (test tooling): 256 claims x 8M rows (the ridge of the two roofs) with parts of the pipeline disabled."""
import os, sys
sys.path.insert(0, '/root/repo')
import torch, drs_b200 as drs
def scan_ms(q, c, k, iters=10):
    prof = []
    for _ in range(3): drs.search(q, c, k, profile=prof)
    prof.clear()
    for _ in range(iters): drs.search(q, c, k, profile=prof)
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in prof) / len(prof)
dev = torch.device("cuda:0"); nc = 8_000_000
g = torch.Generator(device=dev).manual_seed(1)
c = torch.empty(nc, 768, dtype=torch.bfloat16, device=dev)
for r0 in range(0, nc, 1 << 20):
    r1 = min(nc, r0 + (1 << 20))
    c[r0:r1] = torch.nn.functional.normalize(torch.randn(r1 - r0, 768, generator=g, device=dev), dim=1)
q = torch.nn.functional.normalize(torch.randn(256, 768, generator=g, device=dev), dim=1).bfloat16()
print("hbm roof %.3f ms, mma roof %.3f ms" % (nc*1536/6.551e12*1e3, 2.0*256*nc*768/1389.5e12*1e3))
for cg in (2, 1):
    drs.set_option("search.cta_group", cg)
    line = f"cg{cg}:"
    for flags, label in ((0, "full"), (1, "no_functor"), (3, "no_tmem_ld"), (7, "tma_only"), (4, "no_mma")):
        drs.set_option("debug.flags", flags)
        line += f"  {label} {scan_ms(q, c, 10):6.3f}"
    drs.set_option("debug.flags", 0)
    print(line, flush=True)
