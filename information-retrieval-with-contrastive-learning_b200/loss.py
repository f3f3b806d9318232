"""Drop-in ``NCELoss`` / ``InfoNCE`` modules backed by the fused CUDA kernels.

Mirror of src/contrastor/contrastive_loss.py: same class names, constructor argument
(``loss_config`` dict: 'temperature', optional 'cluster'), ``forward`` signatures and reduction
(``NCELoss``: sum over the 2N rows, divided by 2, :92; ``InfoNCE``: mean, :24,:42).  The modules
hold no parameters or buffers, so ``RetrievalModelWrapper.state_dict()`` is unchanged
(checkpoints load with strict=True, src/model.py:93).

``loss_config['precision']`` (extension): 'auto' (default) = 'bf16' when the shapes allow it and the embeddings are
at least 64 wide, else 'fp32'; 'bf16' = inputs rounded to bf16, tcgen05 MMA, fp32 accumulation and fp32 softmax; 'fp32' = FFMA
path for exact comparison.

NUMERICS DIFFER FROM THE REFERENCE ON THE DEFAULT PATH.  The reference computes the logits with an fp32
matmul (contrastive_loss.py:62); the default here rounds q, k, the queue / prototypes and the
gradient-of-logits matrix to bf16 (BASELINE.json north star: "bf16 in, fp32 accumulate, with an fp32
path for exact comparison").  Measured against the reference's own outputs: loss within 2e-2 relative
(observed ~1e-3), every gradient ROW within 3e-2 relative L2 error and cosine >= 0.999
(tests/test_infonce_gpu.py).  Pass ``loss_config['precision'] = 'fp32'`` for 1e-5 / 1e-4 parity.
"""
from __future__ import annotations

import ctypes
import random

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


def _as_f32(t):
    """detach + fp32 + contiguous; a no-op chain for what a trainer passes (fp32, contiguous) -- below ~4096 x 768 the
    loss step is bound by host time, so the common case avoids dispatcher round trips that change nothing"""
    t = t.detach()
    if t.dtype is not torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def _device_guard(dev):
    """torch.cuda.device(dev) only when dev is not already the current device"""
    return _NO_GUARD if torch.cuda.current_device() == dev.index else torch.cuda.device(dev)


_need_cache: dict = {}


def _workspace_bytes(query, *key):
    """cached *_workspace_bytes(...) answers; engine options that change a layout (tune.k_split, tune.symmetric_lse) bump
    _lib.options_epoch, which is part of the key"""
    k = (query, _lib.options_epoch(), torch.cuda.current_device()) + key
    need = _need_cache.get(k)
    if need is None:
        out = ctypes.c_size_t(0)
        _lib.check(getattr(_lib.load(), query)(*key, ctypes.byref(out)))
        if len(_need_cache) > 256:
            _need_cache.clear()
        need = _need_cache[k] = out.value
    return need


def _pick_precision(requested, n, dim, klen):
    """In-batch form: the 2N x 2N gradient matrix needs 2N % 8 == 0 and the queue a 16-byte row pitch.  The
    row x column forms (MoCo queue, prototypes) pass n = 4, klen = 0: their column count may be ragged (the
    engine pads the pitch and lets the tensor maps zero-fill), only dim % 8 == 0 remains."""
    bf16_ok = dim % 8 == 0 and n % 4 == 0 and klen % 8 == 0
    if requested in (None, "auto"):
        # narrow embeddings stay in fp32: with T = 0.05 the bf16 rounding of a 16-wide dot product moves single gradient
        # rows by ~3 % (tools/gpu_loss_emulation_check.py: the kernel matches a bf16 emulation to 1e-4, the emulation is
        # 2.6e-2 off the fp32 result at dim = 16, 7e-3 at dim = 64) -- and there is nothing to gain at that size
        return _lib.DRS_BF16 if (bf16_ok and dim >= 64) else _lib.DRS_F32
    if requested == "bf16":
        if not bf16_ok:
            raise RuntimeError("bf16 path needs dim % 8 == 0 (and, for the in-batch form, batch % 4 == 0 and queue_len % 8 == 0)")
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
        qf = _as_f32(q)
        kf = _as_f32(k)
        n, dim = qf.shape
        qu = None
        klen = 0
        if queue is not None:
            if queue.dim() != 2 or queue.shape[0] != dim:
                raise ValueError(f"queue must be [D, K] with D={dim}, got {tuple(queue.shape)}")
            qu = _as_f32(queue if queue.device == q.device else queue.to(device=q.device))   # :80 queue.clone().detach()
            klen = qu.shape[1]
        prec = _pick_precision(precision, n, dim, klen)
        lib = _lib.load()
        dev = q.device
        with _device_guard(dev):
            need = _workspace_bytes("drs_infonce_workspace_bytes", n, dim, klen, prec)
            # when a backward will follow, the step owns its workspace: the packed operands and row LSEs the
            # forward leaves there are reused by the backward instead of being staged a second time
            keep = bool(ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
            ws = torch.empty(need, dtype=torch.uint8, device=dev) if keep else _workspace(dev, need)
            loss = torch.empty(1, dtype=torch.float32, device=dev)
            lse = torch.empty(2 * n, dtype=torch.float32, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.drs_infonce_forward(qf.data_ptr(), kf.data_ptr(), qu.data_ptr() if qu is not None else None,
                                               n, dim, klen, float(inv_t), prec, loss.data_ptr(), lse.data_ptr(),
                                               ws.data_ptr(), ws.numel(), stream))
        ctx.save_for_backward(qf, kf, qu if qu is not None else torch.empty(0, device=dev), lse)
        ctx.staged_ws = ws if keep else None
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
        with _device_guard(dev):
            need = _workspace_bytes("drs_infonce_workspace_bytes", n, dim, klen, ctx.prec)
            staged = ctx.staged_ws is not None and ctx.staged_ws.numel() >= need
            ws = ctx.staged_ws if staged else _workspace(dev, need)
            ctx.staged_ws = None            # a second backward through the same graph stages again (H overwrote nothing it needs, but keep it simple)
            g = _as_f32(grad_out if grad_out.device == dev else grad_out.to(device=dev)).reshape(1)
            dq = torch.empty_like(qf)
            dk = torch.empty_like(kf)
            stream = torch.cuda.current_stream(dev).cuda_stream
            backward = lib.drs_infonce_backward_staged if staged else lib.drs_infonce_backward
            _lib.check(backward(qf.data_ptr(), kf.data_ptr(), qu.data_ptr() if ctx.has_queue else None,
                                                n, dim, klen, ctx.inv_t, ctx.prec, lse.data_ptr(), g.data_ptr(),
                                                dq.data_ptr(), dk.data_ptr(), ws.data_ptr(), ws.numel(), stream))
        f32 = torch.float32
        return (dq if ctx.in_dtypes[0] is f32 else dq.to(ctx.in_dtypes[0]),
                dk if ctx.in_dtypes[1] is f32 else dk.to(ctx.in_dtypes[1]), None, None, None)


def info_nce_loss(q, k, queue=None, temperature=0.05, precision="auto"):
    """Functional form of ``NCELoss._compute_info_loss`` (contrastive_loss.py:56-93)."""
    return _InfoNceFunction.apply(q, k, queue, 1.0 / float(temperature), precision)


class _MocoFunction(torch.autograd.Function):
    """``InfoNCE.forward`` (contrastive_loss.py:26-44): logits = [q.k | q @ queue] / T, CE mean."""

    @staticmethod
    def forward(ctx, q, k, queue, inv_t, precision):
        if not (q.is_cuda and k.is_cuda):
            raise RuntimeError("drs_b200 InfoNCE needs CUDA tensors: there is no CPU path")
        if q.shape != k.shape or q.dim() != 2:
            raise ValueError(f"q and k must both be [N, D], got {tuple(q.shape)} and {tuple(k.shape)}")
        qf = q.detach().contiguous().float()
        kf = k.detach().contiguous().float()
        n, dim = qf.shape
        if queue is None or queue.dim() != 2 or queue.shape[0] != dim:
            raise ValueError(f"queue must be [D, K] with D={dim}")
        qu = queue.detach().to(device=q.device).contiguous().float()          # :32 queue.clone().detach()
        klen = qu.shape[1]
        prec = _pick_precision(precision, 4, dim, 0)                          # no batch-size / queue-length constraint here
        lib = _lib.load()
        dev = q.device
        with torch.cuda.device(dev):
            need = ctypes.c_size_t(0)
            _lib.check(lib.drs_moco_workspace_bytes(n, dim, klen, prec, ctypes.byref(need)))
            ws = _workspace(dev, need.value)
            loss = torch.empty(1, dtype=torch.float32, device=dev)
            lse = torch.empty(n, dtype=torch.float32, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.drs_moco_forward(qf.data_ptr(), kf.data_ptr(), qu.data_ptr(), n, dim, klen, float(inv_t), prec,
                                            loss.data_ptr(), lse.data_ptr(), ws.data_ptr(), ws.numel(), stream))
        ctx.save_for_backward(qf, kf, qu, lse)
        ctx.inv_t, ctx.prec, ctx.in_dtypes = float(inv_t), prec, (q.dtype, k.dtype)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        qf, kf, qu, lse = ctx.saved_tensors
        n, dim = qf.shape
        klen = qu.shape[1]
        lib = _lib.load()
        dev = qf.device
        with torch.cuda.device(dev):
            need = ctypes.c_size_t(0)
            _lib.check(lib.drs_moco_workspace_bytes(n, dim, klen, ctx.prec, ctypes.byref(need)))
            ws = _workspace(dev, need.value)
            g = grad_out.detach().reshape(1).to(device=dev, dtype=torch.float32).contiguous()
            dq, dk = torch.empty_like(qf), torch.empty_like(kf)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.drs_moco_backward(qf.data_ptr(), kf.data_ptr(), qu.data_ptr(), n, dim, klen, ctx.inv_t, ctx.prec,
                                             lse.data_ptr(), g.data_ptr(), dq.data_ptr(), dk.data_ptr(), ws.data_ptr(),
                                             ws.numel(), stream))
        return dq.to(ctx.in_dtypes[0]), dk.to(ctx.in_dtypes[1]), None, None, None


class InfoNCE(torch.nn.Module):
    """contrastive_loss.py:20-44 (MoCo form).  ``forward(q, k, queue)``; mean-reduced CE, label 0."""

    def __init__(self, loss_config):
        super().__init__()
        self.T = loss_config['temperature']                       # :23
        self.precision = loss_config.get('precision', 'auto')

    def forward(self, q, k, queue):
        return _MocoFunction.apply(q, k, queue, 1.0 / float(self.T), self.precision)


class _ProtoFunction(torch.autograd.Function):
    """Sum over cluster sets of CE_sum(q @ protos_s^T / temps_s, label i) / num_sets
    (contrastive_loss.py:112-134), prototypes already selected."""

    @staticmethod
    def forward(ctx, q, precision, *sets):
        if not q.is_cuda:
            raise RuntimeError("drs_b200 ProtoNCE needs CUDA tensors: there is no CPU path")
        qf = q.detach().contiguous().float()
        n, dim = qf.shape
        dev = q.device
        protos = [s.detach().to(device=dev).contiguous().float() for s in sets[0::2]]
        inv_temps = [(1.0 / s.detach().to(device=dev).float()).contiguous() for s in sets[1::2]]
        lib = _lib.load()
        total = torch.zeros((), dtype=torch.float32, device=dev)
        lses, precs = [], []
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            for pr, it in zip(protos, inv_temps):
                p = pr.shape[0]
                prec = _pick_precision(precision, 4, dim, 0)        # any prototype count (22, 84, 3200, ...)
                need = ctypes.c_size_t(0)
                _lib.check(lib.drs_proto_workspace_bytes(n, dim, p, prec, ctypes.byref(need)))
                ws = _workspace(dev, need.value)
                loss = torch.empty(1, dtype=torch.float32, device=dev)
                lse = torch.empty(n, dtype=torch.float32, device=dev)
                _lib.check(lib.drs_proto_forward(qf.data_ptr(), pr.data_ptr(), it.data_ptr(), n, dim, p, prec,
                                                 loss.data_ptr(), lse.data_ptr(), ws.data_ptr(), ws.numel(), stream))
                total = total + loss[0]
                lses.append(lse)
                precs.append(prec)
        ctx.save_for_backward(qf, *protos, *inv_temps, *lses)
        ctx.num_sets, ctx.precs, ctx.in_dtype = len(protos), precs, q.dtype
        return total / len(protos)                                # :134

    @staticmethod
    def backward(ctx, grad_out):
        saved = ctx.saved_tensors
        ns = ctx.num_sets
        qf, protos, inv_temps, lses = saved[0], saved[1:1 + ns], saved[1 + ns:1 + 2 * ns], saved[1 + 2 * ns:]
        n, dim = qf.shape
        dev = qf.device
        lib = _lib.load()
        dq = torch.empty_like(qf)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            g = (grad_out.detach().reshape(1).to(device=dev, dtype=torch.float32) / ns).contiguous()
            for s, (pr, it, lse) in enumerate(zip(protos, inv_temps, lses)):
                p = pr.shape[0]
                need = ctypes.c_size_t(0)
                _lib.check(lib.drs_proto_workspace_bytes(n, dim, p, ctx.precs[s], ctypes.byref(need)))
                ws = _workspace(dev, need.value)
                _lib.check(lib.drs_proto_backward(qf.data_ptr(), pr.data_ptr(), it.data_ptr(), n, dim, p, ctx.precs[s],
                                                  lse.data_ptr(), g.data_ptr(), dq.data_ptr(), 1 if s else 0,
                                                  ws.data_ptr(), ws.numel(), stream))
        return (dq.to(ctx.in_dtype), None) + (None,) * (2 * ns)


def proto_nce_loss(q, protos, temps, precision="auto"):
    """Functional ProtoNCE on selected prototypes: ``protos[s]`` is [N + r, D] with row i the positive
    prototype of q[i], ``temps[s]`` the matching densities (contrastive_loss.py:112,:122-123)."""
    flat = []
    for pr, tp in zip(protos, temps):
        flat += [pr, tp]
    return _ProtoFunction.apply(q, precision, *flat)


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

    def select_prototypes(self, cluster_result, index):
        """The host-side selection of contrastive_loss.py:99-112,:122-123: positive prototype of
        each sample, ``num_neg_proto`` sampled negatives, and their densities.
        ``random.sample(set, r)`` (:109) raises on Python >= 3.11, so the population is sorted
        first (a set has no defined order anyway); the draw comes from the global ``random`` stream like
        upstream's ``from random import sample``, so ``random.seed(args.seed)`` (main.py:95) governs it;
        everything else follows the reference,
        including ``range(emb2cluster.max())`` (:105), which leaves the highest cluster id out."""
        protos, temps = [], []
        for emb2cluster, prototypes, density in zip(cluster_result['emb2cluster'], cluster_result['centroids'],
                                                    cluster_result['density']):
            pos_proto_id = emb2cluster[index.tolist()]                                   # :101
            all_proto_id = range(int(emb2cluster.max()))                                 # :105
            neg_proto_id = sorted(set(all_proto_id) - set(pos_proto_id.tolist()))        # :106
            neg_proto_id = random.sample(neg_proto_id, self.num_neg_proto)               # :109 (the global RNG, as upstream)
            ids = torch.cat([pos_proto_id.cpu().long(), torch.LongTensor(neg_proto_id)])
            protos.append(prototypes[ids.to(prototypes.device)])                         # :102,:110,:112
            temps.append(density[ids.to(density.device)])                                # :122-123
        return protos, temps

    def _compute_proto_loss(self, q, cluster_result, index):
        protos, temps = self.select_prototypes(cluster_result, index)
        return proto_nce_loss(q, protos, temps, self.precision)   # :115-134

    def forward(self, q, k, queue, cluster_result=None, index=None):
        loss = self._compute_info_loss(q, k, queue)               # :138
        if cluster_result is not None:                            # :139-140
            loss = loss + self._compute_proto_loss(q, cluster_result, index)
        return loss
