"""Host-side MoCo plumbing next to the loss: the key queue and the momentum encoder update of
``RetrievalModelWrapper`` (src/contrastor/contrastive_module.py:24-30, :42-68).

These are a strided copy and an axpy per parameter -- torch ops on whatever device the tensors live on, no engine
kernel involved; they are here so that a wrapper built around ``drs_b200.NCELoss`` needs nothing else from the
reference's module.  Pinned by tests/golden/queue_maintenance.npz (the reference's own methods, run by make_golden.py)."""
from __future__ import annotations

import torch


def new_queue(dim: int, queue_size: int, device=None, generator=None):
    """contrastive_module.py:26-29: a [dim, queue_size] queue of unit columns and its write pointer."""
    queue = torch.nn.functional.normalize(torch.randn(dim, queue_size, generator=generator), dim=0)
    ptr = torch.zeros(1, dtype=torch.long)
    return (queue.to(device), ptr.to(device)) if device is not None else (queue, ptr)


@torch.no_grad()
def dequeue_and_enqueue(queue: torch.Tensor, queue_ptr: torch.Tensor, keys: torch.Tensor) -> None:
    """contrastive_module.py:55-68, in place: the batch of keys replaces the columns at the pointer, the pointer
    advances modulo the queue size.  A batch size that does not divide the queue leaves both untouched (:59)."""
    batch_size = keys.shape[0]
    queue_size = queue.shape[1]
    if queue_size % batch_size == 0:            # :59
        ptr = int(queue_ptr)                    # :60
        queue[:, ptr:ptr + batch_size] = keys.T  # :63
        queue_ptr[0] = (ptr + batch_size) % queue_size   # :66-68


@torch.no_grad()
def momentum_update(params_q, params_k, momentum: float) -> None:
    """contrastive_module.py:42-52: ``param_k = param_k * m + param_q * (1 - m)`` for every parameter pair, as three
    fused multi-tensor launches (the same fp32 operations in the same order: bit-identical to the reference's loop)."""
    ks = [p.data for p in params_k]
    qs = [p.data for p in params_q]
    if not ks:
        return
    torch._foreach_mul_(ks, momentum)
    torch._foreach_add_(ks, torch._foreach_mul(qs, 1. - momentum))
