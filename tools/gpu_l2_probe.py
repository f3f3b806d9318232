#!/usr/bin/env python
"""L2-reuse probe (test tooling): search launches per tuning variant.  Run plain for CUDA-event
timings, or under `ncu --metrics dram__bytes_read.sum,... -k regex:gemm_nt_tc` for DRAM bytes.
usage: gpu_l2_probe.py <corpus rows> <variant>[,<variant>...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import drs_b200 as drs  # noqa: E402

VARIANTS = {
    "barrier": {},
    "no_barrier": {"tune.round_barrier": 0},
    "b_evict_first": {"tune.b_hint": 1},
    "b_evict_last": {"tune.b_hint": 2},
    "ctas80": {"search.num_ctas": 80},
    "splits74": {"search.splits": 74},
    "cg1": {"search.cta_group": 1},
}
DEFAULTS = {"tune.b_hint": 0, "tune.stagger_cycles": 0, "search.num_ctas": 0, "search.splits": 0,
            "search.cta_group": 0, "tune.round_barrier": 1}

nq, dim, k = 10000, 768, 10
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
names = sys.argv[2].split(",") if len(sys.argv) > 2 else ["barrier", "no_barrier"]
g = torch.Generator(device="cuda").manual_seed(1337)
c = torch.empty(nc, dim, dtype=torch.bfloat16, device="cuda")
for r0 in range(0, nc, 1 << 20):
    r1 = min(nc, r0 + (1 << 20))
    c[r0:r1] = torch.nn.functional.normalize(torch.randn(r1 - r0, dim, generator=g, device="cuda"), dim=1)
q = torch.nn.functional.normalize(torch.randn(nq, dim, generator=g, device="cuda"), dim=1).bfloat16()
for name in names:
    for o, v in DEFAULTS.items():
        drs.set_option(o, v)
    for o, v in VARIANTS[name].items():
        drs.set_option(o, v)
    drs.search(q, c, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        drs.search(q, c, k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"nc={nc} {name:14s} {ms:9.3f} ms  {2.0 * nq * nc * dim / ms / 1e9:8.1f} TFLOP/s", flush=True)
