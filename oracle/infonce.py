"""CPU oracle: in-batch InfoNCE logits + softmax cross-entropy.  TEST INFRASTRUCTURE.

Restates ``NCELoss._compute_info_loss`` (src/contrastor/contrastive_loss.py:56-93),
``InfoNCE.forward`` (:26-44) and ``NCELoss._compute_proto_loss`` (:95-135) of the
reference in closed form, with analytic gradients.  Pinned against the reference's
own classes run in the build container: tests/golden/make_golden.py imports them
from /root/reference and commits loss + autograd gradients as golden vectors.

Closed form of ``_compute_info_loss`` (what lines :57-92 amount to):

    F = cat(q, k)                       (:61)     2N x D
    S = F @ F.T                         (:62)     2N x 2N
    row i: positive  j = (i + N) mod 2N (:57-58,:71), the diagonal is dropped (:65-68)
           negatives = every other column (:74)
           + queue logits  q[i mod N] . queue[:, c]   (:79-80, note ``.repeat(2, 1)``:
             rows N..2N-1 reuse the QUERY rows' queue logits, not k's)
    logits /= T (:88);  loss = sum_i CE(logits_i, label 0) / 2   (:91-92)
         = sum_i [ logsumexp_{j != i}(S_ij / T  (+) queue_i / T) - S_{i,pos(i)} / T ] / 2
"""
from __future__ import annotations

import torch


def nce_info_loss(q: torch.Tensor, k: torch.Tensor, queue: torch.Tensor | None, temperature: float,
                  dtype=torch.float32):
    """Returns (loss, dq, dk) for upstream gradient 1.  contrastive_loss.py:56-93."""
    q = q.to(dtype)
    k = k.to(dtype)
    n = q.shape[0]
    f = torch.cat([q, k], dim=0)                                   # :61
    s = (f @ f.T) / temperature                                    # :62, :88
    two_n = 2 * n
    eye = torch.eye(two_n, dtype=torch.bool)
    s_masked = s.masked_fill(eye, float("-inf"))                   # :65-68 diagonal dropped
    pos_col = (torch.arange(two_n) + n) % two_n                    # :57-58
    pos = s[torch.arange(two_n), pos_col]                          # :71
    if queue is not None:
        lq = (q @ queue.to(dtype)) / temperature                   # :79 einsum('nc,ck->nk')
        lq2 = lq.repeat(2, 1)                                      # :80
        logits = torch.cat([s_masked, lq2], dim=1)
    else:
        logits = s_masked
    lse = torch.logsumexp(logits, dim=1)
    loss = (lse - pos).sum() / 2                                   # :92 CE(sum)/2

    # analytic gradient: dL/dlogit_ij = (softmax_ij - onehot_pos) / 2
    p = torch.exp(logits - lse[:, None])
    g = p[:, :two_n].clone()
    g[torch.arange(two_n), pos_col] -= 1.0
    g = g / (2 * temperature)
    df = g @ f + g.T @ f                                           # S = F F^T, both operands
    dq, dk = df[:n].clone(), df[n:].clone()
    if queue is not None:
        pq = p[:, two_n:] / (2 * temperature)                      # 2N x K
        dq += (pq[:n] + pq[n:]) @ queue.to(dtype).T                # both halves use q rows
    return loss, dq, dk


def moco_infonce(q: torch.Tensor, k: torch.Tensor, queue: torch.Tensor, temperature: float,
                 dtype=torch.float32):
    """``InfoNCE.forward`` (contrastive_loss.py:26-44): logits = [q.k | q @ queue] / T,
    label 0, CrossEntropyLoss with MEAN reduction.  Returns (loss, dq, dk)."""
    q = q.to(dtype)
    k = k.to(dtype)
    n = q.shape[0]
    l_pos = (q * k).sum(1, keepdim=True)                           # :30
    l_neg = q @ queue.to(dtype)                                    # :32
    logits = torch.cat([l_pos, l_neg], dim=1) / temperature        # :34-37
    lse = torch.logsumexp(logits, dim=1)
    loss = (lse - logits[:, 0]).mean()                             # :42 (mean)
    p = torch.exp(logits - lse[:, None])
    p[:, 0] -= 1.0
    p = p / (temperature * n)
    dq = p[:, :1] * k + p[:, 1:] @ queue.to(dtype).T
    dk = p[:, :1] * q
    return loss, dq, dk


def proto_loss(q: torch.Tensor, protos: list[torch.Tensor], temps: list[torch.Tensor],
               dtype=torch.float32):
    """``NCELoss._compute_proto_loss`` (contrastive_loss.py:95-135) AFTER prototype
    selection: for each cluster set, ``protos[s]`` is ``cat(pos_prototypes, neg_prototypes)``
    (:112, (N+r) x D) and ``temps[s]`` the matching densities (:122-123).  The sampling at
    :105-110 uses ``random.sample(set, r)`` which raises on Python >= 3.11, so the oracle
    takes the selected prototypes as input.
    logits = q @ protos.T / temps (:115,:124); label of row i is i (:118-119);
    loss = sum_sets CE_sum / num_sets (:129-134).  Returns (loss, dq)."""
    q = q.to(dtype)
    n = q.shape[0]
    loss = torch.zeros((), dtype=dtype)
    dq = torch.zeros_like(q)
    for pr, tp in zip(protos, temps):
        pr = pr.to(dtype)
        logits = (q @ pr.T) / tp.to(dtype)[None, :]
        lse = torch.logsumexp(logits, dim=1)
        loss = loss + (lse - logits[torch.arange(n), torch.arange(n)]).sum()
        p = torch.exp(logits - lse[:, None])
        p[torch.arange(n), torch.arange(n)] -= 1.0
        dq += (p / tp.to(dtype)[None, :]) @ pr
    return loss / len(protos), dq / len(protos)
