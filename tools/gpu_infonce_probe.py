#!/usr/bin/env python
"""Perf probe (test tooling): one NCELoss forward+backward at BASELINE configs[3] (4096 x 768) --
run under `ncu --metrics gpu__time_duration.sum` for the launch list, or plain for the step time."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import drs_b200 as drs  # noqa: E402


def main():
    n, dim = 4096, 768
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1337)
    q = torch.nn.functional.normalize(torch.randn(n, dim, generator=g, device=dev), dim=1).requires_grad_(True)
    k = torch.nn.functional.normalize(torch.randn(n, dim, generator=g, device=dev) * 0.5 + q.detach(), dim=1).requires_grad_(True)
    crit = drs.NCELoss({"temperature": 0.05})

    def step():
        q.grad = None
        k.grad = None
        crit(q, k, None).backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    host = (time.perf_counter() - t0) / iters
    torch.cuda.synchronize()
    print(f"NCELoss fwd+bwd {n}x{dim}: {e0.elapsed_time(e1) / iters:.3f} ms/step (host enqueue {host * 1e3:.3f} ms/step)")
    # the same math as separate library calls (cuBLAS GEMM + ATen softmax-CE), for scale
    f = torch.cat([q.detach(), k.detach()]).requires_grad_(True)

    def torch_step():
        f.grad = None
        s = (f @ f.T) / 0.05
        s = s.masked_fill(torch.eye(2 * n, dtype=torch.bool, device=dev), float("-inf"))
        tgt = (torch.arange(2 * n, device=dev) + n) % (2 * n)
        (torch.nn.functional.cross_entropy(s, tgt, reduction="sum") / 2).backward()

    for _ in range(3):
        torch_step()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch_step()
    e1.record()
    torch.cuda.synchronize()
    print(f"torch fp32 matmul + cross_entropy (closed form) {n}x{dim}: {e0.elapsed_time(e1) / iters:.3f} ms/step")


if __name__ == "__main__":
    main()


def reference_config():
    """The reference's own training shapes (config.yaml: batch 128, output 128, queue 12544, T 0.05)."""
    n, dim, klen = 128, 128, 12544
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(7)
    q = torch.nn.functional.normalize(torch.randn(n, dim, generator=g, device=dev), dim=1).requires_grad_(True)
    k = torch.nn.functional.normalize(torch.randn(n, dim, generator=g, device=dev), dim=1).requires_grad_(True)
    queue = torch.nn.functional.normalize(torch.randn(dim, klen, generator=g, device=dev), dim=0)
    crit = drs.NCELoss({"temperature": 0.05})

    def step():
        q.grad = None
        k.grad = None
        crit(q, k, queue).backward()

    def torch_step():
        q.grad = None
        k.grad = None
        f = torch.cat([q, k])
        s = f @ f.T
        s = s.masked_fill(torch.eye(2 * n, dtype=torch.bool, device=dev), float("-inf"))
        lq = (q @ queue).repeat(2, 1)
        logits = torch.cat([s, lq], dim=1) / 0.05
        tgt = (torch.arange(2 * n, device=dev) + n) % (2 * n)
        (torch.nn.functional.cross_entropy(logits, tgt, reduction="sum") / 2).backward()

    for name, fn in (("drs_b200 NCELoss", step), ("torch closed form", torch_step)):
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"reference config (N=128, D=128, queue 12544) {name}: {e0.elapsed_time(e1) / 100 * 1e3:.1f} us/step")


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[2] == "refcfg":
    reference_config()


def graph_replay():
    """NCELoss forward+backward captured once in a CUDA graph (static q, k) and replayed: the step without any
    host-side launch work."""
    n, dim = 4096, 768
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1337)
    q = torch.nn.functional.normalize(torch.randn(n, dim, generator=g, device=dev), dim=1).requires_grad_(True)
    k = torch.nn.functional.normalize(torch.randn(n, dim, generator=g, device=dev) * 0.5 + q.detach(), dim=1).requires_grad_(True)
    crit = drs.NCELoss({"temperature": 0.05})
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            q.grad = None
            k.grad = None
            crit(q, k, None).backward()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    ref_dq = q.grad.clone()
    graph = torch.cuda.CUDAGraph()
    q.grad = None
    k.grad = None
    with torch.cuda.graph(graph):
        loss = crit(q, k, None)
        loss.backward()
    graph.replay()
    torch.cuda.synchronize()
    same = torch.equal(q.grad, ref_dq)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"NCELoss fwd+bwd {n}x{dim} replayed from a CUDA graph: {e0.elapsed_time(e1) / 100:.3f} ms/step, gradients equal to eager: {same}")


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[2] == "graph":
    graph_replay()
