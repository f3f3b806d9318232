#!/usr/bin/env python
"""Perf probe (test tooling): where the HOST time of an NCELoss step goes -- the bare C-ABI calls (enqueue only, fixed
buffers) against the autograd module around them.  Below ~4096 x 768 the step is bound by this, not by the GPU.
    python tools/gpu_host_overhead_probe.py [n] [dim] [queue_len]"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import drs_b200 as drs  # noqa: E402
from drs_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 128
klen = int(sys.argv[3]) if len(sys.argv) > 3 else 12544
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
unit = lambda x: torch.nn.functional.normalize(x, dim=1)
q = unit(torch.randn(n, dim, generator=g, device=dev)).requires_grad_(True)
k = unit(torch.randn(n, dim, generator=g, device=dev)).requires_grad_(True)
queue = torch.nn.functional.normalize(torch.randn(dim, klen, generator=g, device=dev), dim=0) if klen else None
lib = _lib.load()
prec = _lib.DRS_BF16
need = ctypes.c_size_t(0)
_lib.check(lib.drs_infonce_workspace_bytes(n, dim, klen, prec, ctypes.byref(need)))
ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
loss = torch.empty(1, device=dev)
lse = torch.empty(2 * n, device=dev)
gout = torch.ones(1, device=dev)
dq, dk = torch.empty_like(q), torch.empty_like(k)
stream = torch.cuda.current_stream().cuda_stream
qp = queue.data_ptr() if klen else None


def fwd():
    _lib.check(lib.drs_infonce_forward(q.data_ptr(), k.data_ptr(), qp, n, dim, klen, 20.0, prec, loss.data_ptr(), lse.data_ptr(),
                                       ws.data_ptr(), ws.numel(), stream))


def bwd():
    _lib.check(lib.drs_infonce_backward_staged(q.data_ptr(), k.data_ptr(), qp, n, dim, klen, 20.0, prec, lse.data_ptr(),
                                               gout.data_ptr(), dq.data_ptr(), dk.data_ptr(), ws.data_ptr(), ws.numel(), stream))


crit = drs.NCELoss({"temperature": 0.05, "precision": "bf16"})


def module_step():
    q.grad = None
    k.grad = None
    crit(q, k, queue).backward()


def module_fwd_only():
    with torch.no_grad():
        crit(q, k, queue)


def host_time(fn, reps=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return (t1 - t0) / reps * 1e6, (t2 - t0) / reps * 1e6


print(f"n={n} dim={dim} queue={klen}  (enqueue us/call, wall us/call incl. drain)")
for name, fn in [("C ABI forward", fwd), ("C ABI backward (staged)", bwd), ("C ABI forward + backward", lambda: (fwd(), bwd())),
                 ("module forward (no_grad)", module_fwd_only), ("module forward + backward", module_step)]:
    enq, wall = host_time(fn)
    print(f"  {name:32s} {enq:8.1f} {wall:8.1f}")
