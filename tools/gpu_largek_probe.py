"""GPU probe (not a pytest): cost of k = 100 (adaptive passes) next to k = 10 / 32 on the same shapes.
BASELINE configs[4] per-GPU share: 65 536 claims x 675 000 rows (5.4M / 8), top-100."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drs_b200  # noqa: E402
from importlib import import_module  # noqa: E402

retrieval = import_module(drs_b200.__name__ + ".retrieval")


def timed(q, c, k, iters=3):
    drs_b200.search(q, c, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        drs_b200.search(q, c, k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1337)
    out = {}
    for name, nq, nc in (("cfg5_share_65k_x_675k", 65536, 675000), ("10k_x_5.4M", 10000, 5_400_000)):
        c = torch.nn.functional.normalize(torch.randn(nc, 768, generator=g, device=dev), dim=1).bfloat16()
        q = torch.nn.functional.normalize(torch.randn(nq, 768, generator=g, device=dev), dim=1).bfloat16()
        for k in (10, 32, 100):
            ms = timed(q, c, k)
            tf = 2.0 * nq * nc * 768 / (ms * 1e-3) / 1e12
            rec = {"ms": round(ms, 3), "tflops": round(tf, 1), "claims_per_s": round(nq / (ms * 1e-3), 1)}
            if k > 32:
                rec["open_claims_per_pass"] = retrieval.open_claims_per_pass()[:4]
            out[f"{name}_k{k}"] = rec
            print(name, k, rec, flush=True)
        del c, q
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/largek_probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
