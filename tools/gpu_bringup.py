#!/usr/bin/env python
"""GPU bring-up harness (test tooling, not product code).

Runs each kernel variant in its OWN subprocess with a timeout, so a trap or a hang in one
variant cannot take the others down, and prints one line per case.  Usage on the GPU box:

    python tests/gpu_bringup.py            # all cases
    python tests/gpu_bringup.py case_name  # one case, in-process
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _setup():
    import torch
    import drs_b200
    from oracle import dense_topk, infonce
    return torch, drs_b200, dense_topk, infonce


def _unit(x):
    import torch
    return torch.nn.functional.normalize(x, dim=1)


def _search_case(dtype_name, cg, nq, nc, dim, k, planted=True, splits=0):
    torch, drs, dense_topk, _ = _setup()
    dt = {"f32": torch.float32, "bf16": torch.bfloat16}[dtype_name]
    g = torch.Generator(device="cuda").manual_seed(1337)
    c = _unit(torch.randn(nc, dim, generator=g, device="cuda"))
    if planted:
        j = torch.randint(0, nc, (nq,), generator=g, device="cuda")
        q = _unit(c[j] + 0.1 * torch.randn(nq, dim, generator=g, device="cuda"))
    else:
        q = _unit(torch.randn(nq, dim, generator=g, device="cuda"))
    c, q = c.to(dt), q.to(dt)
    drs.set_option("search.cta_group", cg)
    drs.set_option("search.splits", splits)
    t0 = time.time()
    s, i = drs.search(q, c, k)
    torch.cuda.synchronize()
    t1 = time.time()
    # checker: the same values upcast to fp32, fp32 matmul on the GPU (cuBLAS, TF32 off), oracle select on CPU
    torch.backends.cuda.matmul.allow_tf32 = False
    ref_scores = (q.float() @ c.float().T).cpu()
    rv, ri = dense_topk.select_topk_desc(ref_scores, k)
    s, i = s.cpu(), i.cpu()
    tol = 1e-5 if dtype_name == "f32" else 2e-2
    rel = ((s - rv).abs() / rv.abs().clamp_min(1e-3)).max().item()
    # ids must agree wherever the reference gap to the next score exceeds the tolerance
    gap_ok = torch.ones_like(ri, dtype=torch.bool)
    sv, _ = torch.sort(ref_scores, dim=1, descending=True)
    gaps = (sv[:, :k] - sv[:, 1:k + 1]) if nc > k else torch.full((nq, k), 1.0)
    thr = 4e-6 if dtype_name == "f32" else 2e-3
    strict = (gaps > thr)
    strict[:, 1:] &= (gaps[:, :-1] > thr)
    mism = ((i != ri) & strict).sum().item()
    exact = (i == ri).float().mean().item()
    return dict(ok=bool(rel <= tol and mism == 0), rel_err=rel, id_mismatch_strict=mism, id_match_frac=exact,
                first_call_s=round(t1 - t0, 3))


def _infonce_case(precision, cg, n, dim, klen, temp=0.05):
    torch, drs, _, infonce = _setup()
    g = torch.Generator(device="cuda").manual_seed(1337)
    q = _unit(torch.randn(n, dim, generator=g, device="cuda"))
    k = _unit(0.5 * torch.randn(n, dim, generator=g, device="cuda") + q)
    queue = torch.nn.functional.normalize(torch.randn(dim, klen, generator=g, device="cuda"), dim=0) if klen else None
    drs.set_option("infonce.cta_group", cg)
    q.requires_grad_(True)
    k.requires_grad_(True)
    loss = drs.info_nce_loss(q, k, queue, temp, precision)
    (loss * 1.7).backward()
    torch.cuda.synchronize()
    rl, rdq, rdk = infonce.nce_info_loss(q.detach().cpu(), k.detach().cpu(), queue.cpu() if klen else None, temp,
                                         dtype=torch.float64)
    rdq, rdk = rdq * 1.7, rdk * 1.7
    tol = 2e-5 if precision == "fp32" else 2e-2
    el = abs(loss.item() - rl.item()) / abs(rl.item())
    gs = max(rdq.abs().max().item(), rdk.abs().max().item())
    eq = (q.grad.cpu().double() - rdq).abs().max().item() / gs
    ek = (k.grad.cpu().double() - rdk).abs().max().item() / gs
    gtol = 1e-4 if precision == "fp32" else 3e-2
    return dict(ok=bool(el <= tol and eq <= gtol and ek <= gtol), loss=loss.item(), ref=rl.item(), loss_rel=el,
                dq_err=eq, dk_err=ek)


def _bench_case(cg, nq, nc, dim, k, iters=3):
    torch, drs, _, _ = _setup()
    g = torch.Generator(device="cuda").manual_seed(1337)
    c = torch.randn(nc, dim, generator=g, device="cuda", dtype=torch.bfloat16)
    q = torch.randn(nq, dim, generator=g, device="cuda", dtype=torch.bfloat16)
    drs.set_option("search.cta_group", cg)
    drs.search(q, c, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        drs.search(q, c, k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return dict(ok=True, ms=round(ms, 3), tflops=round(2.0 * nq * nc * dim / ms / 1e9, 1),
                gbs=round(nc * dim * 2 / ms / 1e6, 1), claims_per_s=round(nq / ms * 1e3, 1))


CASES = {
    "f32_small": lambda: _search_case("f32", 0, 100, 5000, 64, 5),
    "f32_ragged": lambda: _search_case("f32", 0, 77, 1234, 100, 10, planted=False),
    "f32_cfg1": lambda: _search_case("f32", 0, 1000, 100000, 768, 5),
    "bf16_cg1_tiny": lambda: _search_case("bf16", 1, 128, 256, 64, 5),
    "bf16_cg1_small": lambda: _search_case("bf16", 1, 128, 4096, 64, 5),
    "bf16_cg1_d768": lambda: _search_case("bf16", 1, 300, 20000, 768, 10),
    "bf16_cg1_ragged": lambda: _search_case("bf16", 1, 77, 12345, 200, 10, planted=False),
    "bf16_cg2_tiny": lambda: _search_case("bf16", 2, 256, 256, 64, 5),
    "bf16_cg2_small": lambda: _search_case("bf16", 2, 256, 4096, 64, 5),
    "bf16_cg2_d768": lambda: _search_case("bf16", 2, 300, 20000, 768, 10),
    "bf16_cg2_ragged": lambda: _search_case("bf16", 2, 77, 12345, 200, 10, planted=False),
    "bf16_cg2_big": lambda: _search_case("bf16", 2, 1000, 300000, 768, 10),
    "nce_f32_small": lambda: _infonce_case("fp32", 0, 32, 64, 0),
    "nce_f32_queue": lambda: _infonce_case("fp32", 0, 128, 128, 512),
    "nce_f32_ragged": lambda: _infonce_case("fp32", 0, 50, 100, 37),
    "nce_bf16_cg1": lambda: _infonce_case("bf16", 1, 128, 128, 0),
    "nce_bf16_cg1_queue": lambda: _infonce_case("bf16", 1, 128, 128, 512),
    "nce_bf16_cg2": lambda: _infonce_case("bf16", 2, 512, 768, 0),
    "nce_bf16_cg2_queue": lambda: _infonce_case("bf16", 2, 256, 128, 1024),
    "bench_cg1_10k_1m": lambda: _bench_case(1, 10000, 1000000, 768, 10),
    "bench_cg2_10k_1m": lambda: _bench_case(2, 10000, 1000000, 768, 10),
    "bench_cg2_10k_5m": lambda: _bench_case(2, 10000, 5400000, 768, 10),
    "bench_cg2_10k_25m": lambda: _bench_case(2, 10000, 25000000, 768, 10, iters=2),
    "bench_cg1_128_5m": lambda: _bench_case(1, 128, 5400000, 768, 10),
    "bench_cg1_16_5m": lambda: _bench_case(1, 16, 5400000, 768, 10),
    "bench_cg2_256_5m": lambda: _bench_case(2, 256, 5400000, 768, 10),
}


def main():
    if len(sys.argv) == 2 and sys.argv[1] in CASES:
        try:
            res = CASES[sys.argv[1]]()
        except Exception as e:  # noqa: BLE001
            res = dict(ok=False, error=f"{type(e).__name__}: {e}"[:600])
            try:
                import drs_b200
                res["hang"] = drs_b200._lib.hang_report()
            except Exception:  # noqa: BLE001
                pass
        print("RESULT " + json.dumps(res))
        return
    names = [a for a in sys.argv[1:]] or list(CASES)
    names = [n for n in CASES if any(n.startswith(p) for p in names)] if sys.argv[1:] else names
    summary = {}
    for name in names:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True,
                               timeout=240)
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            res = json.loads(line[-1][7:]) if line else dict(ok=False, error="no result", rc=p.returncode,
                                                             tail=(p.stdout + p.stderr)[-800:])
        except subprocess.TimeoutExpired:
            res = dict(ok=False, error="timeout 240 s")
        res["wall_s"] = round(time.time() - t0, 1)
        summary[name] = res
        print(f"{'PASS' if res.get('ok') else 'FAIL'} {name}: {json.dumps(res)}", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bringup.json"), "w") as f:
        json.dump(summary, f, indent=1)
    print("passed", sum(1 for r in summary.values() if r.get("ok")), "of", len(summary))


if __name__ == "__main__":
    main()
