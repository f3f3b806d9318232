"""Drop-in ``NCELoss`` / ``InfoNCE`` modules backed by the fused CUDA kernels.

Mirror of src/contrastor/contrastive_loss.py: same class names, constructor argument
(``loss_config`` dict: 'temperature', optional 'cluster'), ``forward`` signatures and reduction
(``NCELoss``: sum over the 2N rows, divided by 2, :92; ``InfoNCE``: mean, :24,:42).  The modules
hold no parameters or buffers, so ``RetrievalModelWrapper.state_dict()`` is unchanged
(checkpoints load with strict=True, src/model.py:93).

``loss_config['precision']`` (extension): 'bf16' (default when the shapes allow it: inputs
rounded to bf16, tcgen05 MMA, fp32 accumulate) or 'fp32' (FFMA path for exact comparison).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

_ws_cache: dict = {}


def _workspace(device, nbytes):
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, "loss")
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def _pick_precision(requested, n, dim, klen):
    bf16_ok = dim % 8 == 0 and n % 4 == 0 and klen % 8 == 0
    if requested in (None, "auto"):
        return _lib.DRS_BF16 if bf16_ok else _lib.DRS_F32
    if requested == "bf16":
        if not bf16_ok:
            raise RuntimeError("bf16 InfoNCE needs dim % 8 == 0, batch % 4 == 0 and queue_len % 8 == 0")
        return _lib.DRS_BF16
    if requested == "fp32":
        return _lib.DRS_F32
    raise ValueError(f"precision must be 'auto', 'bf16' or 'fp32', got {requested!r}")


class _InfoNceFunction(torch.autograd.Function):
    """loss = sum_i [logsumexp_{j != i}(S_ij/T (+) queue_i/T) - S_{i,pos(i)}/T] / 2,
    S = cat(q,k) cat(q,k)^T  -- contrastive_loss.py:56-93 in closed form."""

    @staticmethod
    def forward(ctx, q, k, queue, inv_t, precision):
        if not (q.is_cuda and k.is_cuda):
            raise RuntimeError("drs_b200 NCELoss needs CUDA tensors: there is no CPU path")
        if q.shape != k.shape or q.dim() != 2:
            raise ValueError(f"q and k must both be [N, D], got {tuple(q.shape)} and {tuple(k.shape)}")
        qf = q.detach().contiguous().float()
        kf = k.detach().contiguous().float()
        n, dim = qf.shape
        qu = None
        klen = 0
        if queue is not None:
            if queue.dim() != 2 or queue.shape[0] != dim:
                raise ValueError(f"queue must be [D, K] with D={dim}, got {tuple(queue.shape)}")
            qu = queue.detach().to(device=q.device).contiguous().float()   # :80 queue.clone().detach()
            klen = qu.shape[1]
        prec = _pick_precision(precision, n, dim, klen)
        lib = _lib.load()
        dev = q.device
        with torch.cuda.device(dev):
            need = ctypes.c_size_t(0)
            _lib.check(lib.drs_infonce_workspace_bytes(n, dim, klen, prec, ctypes.byref(need)))
            ws = _workspace(dev, need.value)
            loss = torch.empty(1, dtype=torch.float32, device=dev)
            lse = torch.empty(2 * n, dtype=torch.float32, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.drs_infonce_forward(qf.data_ptr(), kf.data_ptr(), qu.data_ptr() if qu is not None else None,
                                               n, dim, klen, float(inv_t), prec, loss.data_ptr(), lse.data_ptr(),
                                               ws.data_ptr(), ws.numel(), stream))
        ctx.save_for_backward(qf, kf, qu if qu is not None else torch.empty(0, device=dev), lse)
        ctx.has_queue = qu is not None
        ctx.inv_t = float(inv_t)
        ctx.prec = prec
        ctx.in_dtypes = (q.dtype, k.dtype)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        qf, kf, qu, lse = ctx.saved_tensors
        n, dim = qf.shape
        klen = qu.shape[1] if ctx.has_queue else 0
        lib = _lib.load()
        dev = qf.device
        with torch.cuda.device(dev):
            need = ctypes.c_size_t(0)
            _lib.check(lib.drs_infonce_workspace_bytes(n, dim, klen, ctx.prec, ctypes.byref(need)))
            ws = _workspace(dev, need.value)
            g = grad_out.detach().reshape(1).to(device=dev, dtype=torch.float32).contiguous()
            dq = torch.empty_like(qf)
            dk = torch.empty_like(kf)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.drs_infonce_backward(qf.data_ptr(), kf.data_ptr(), qu.data_ptr() if ctx.has_queue else None,
                                                n, dim, klen, ctx.inv_t, ctx.prec, lse.data_ptr(), g.data_ptr(),
                                                dq.data_ptr(), dk.data_ptr(), ws.data_ptr(), ws.numel(), stream))
        return dq.to(ctx.in_dtypes[0]), dk.to(ctx.in_dtypes[1]), None, None, None


def info_nce_loss(q, k, queue=None, temperature=0.05, precision="auto"):
    """Functional form of ``NCELoss._compute_info_loss`` (contrastive_loss.py:56-93)."""
    return _InfoNceFunction.apply(q, k, queue, 1.0 / float(temperature), precision)


class NCELoss(torch.nn.Module):
    """contrastive_loss.py:47-141.  ``forward(q, k, queue, cluster_result=None, index=None)``."""

    def __init__(self, loss_config):
        super().__init__()
        self.T = loss_config['temperature']                       # :50
        self.precision = loss_config.get('precision', 'auto')
        if 'cluster' in loss_config:                              # :52-54
            self.num_cluster = loss_config['cluster']['num_cluster']
            self.num_neg_proto = loss_config['cluster']['num_neg_proto']

    def _compute_info_loss(self, q, k, queue=None):
        return info_nce_loss(q, k, queue, self.T, self.precision)

    def _compute_proto_loss(self, q, cluster_result, index):
        raise NotImplementedError(
            "ProtoNCE (contrastive_loss.py:95-135) is a 'next' row of the hot-path scope (SURVEY.md 8f-2) "
            "and is not built yet")

    def forward(self, q, k, queue, cluster_result=None, index=None):
        loss = self._compute_info_loss(q, k, queue)               # :138
        if cluster_result is not None:                            # :139-140
            loss = loss + self._compute_proto_loss(q, cluster_result, index)
        return loss
