"""Parity of the fused InfoNCE forward/backward against the reference's own NCELoss outputs
(golden vectors) and the oracle.  GPU only."""
import glob
import os

import numpy as np
import pytest
import torch

import drs_b200
from oracle import infonce

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda:0"


def _run(q, k, queue, temp, precision, cg=0, upstream=1.0):
    q = torch.as_tensor(q).to(DEV).requires_grad_(True)
    k = torch.as_tensor(k).to(DEV).requires_grad_(True)
    queue = torch.as_tensor(queue).to(DEV) if queue is not None else None
    crit = drs_b200.NCELoss({"temperature": temp, "precision": precision})
    drs_b200.set_option("infonce.cta_group", cg)
    try:
        loss = crit(q, k, queue)
        (loss * upstream).backward()
    finally:
        drs_b200.set_option("infonce.cta_group", 0)
    return loss.detach().cpu(), q.grad.cpu(), k.grad.cpu()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "infonce_*.npz"))))
def test_fp32_path_matches_reference_golden(path):
    """fp32 FFMA path vs loss/grads produced by the reference's NCELoss + autograd:
    1e-5 relative on the loss (north-star fp32 tolerance), 1e-4 of the gradient scale."""
    z = np.load(path)
    queue = z["queue"] if "queue" in z.files else None
    loss, dq, dk = _run(z["q"], z["k"], queue, float(z["temperature"]), "fp32")
    assert loss.shape == () and abs(loss.item() - float(z["loss"])) <= 1e-5 * abs(float(z["loss"])) + 1e-5
    scale = max(np.abs(z["dq"]).max(), np.abs(z["dk"]).max())
    np.testing.assert_allclose(dq.numpy(), z["dq"], rtol=0, atol=1e-4 * scale)
    np.testing.assert_allclose(dk.numpy(), z["dk"], rtol=0, atol=1e-4 * scale)


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("name", ["n32_d64", "n32_d64_q96", "n128_d128_q512", "n96_d768", "n8_d16_q32"])
def test_bf16_tcgen05_path_matches_reference_golden(name, cg):
    """bf16 tcgen05 path: 2e-2 relative (north-star bf16 tolerance)."""
    z = np.load(os.path.join(GOLDEN, f"infonce_{name}.npz"))
    queue = z["queue"] if "queue" in z.files else None
    loss, dq, dk = _run(z["q"], z["k"], queue, float(z["temperature"]), "bf16", cg=cg)
    assert abs(loss.item() - float(z["loss"])) <= 2e-2 * abs(float(z["loss"]))
    scale = max(np.abs(z["dq"]).max(), np.abs(z["dk"]).max())
    np.testing.assert_allclose(dq.numpy(), z["dq"], rtol=0, atol=3e-2 * scale)
    np.testing.assert_allclose(dk.numpy(), z["dk"], rtol=0, atol=3e-2 * scale)


@pytest.mark.parametrize("precision,n,dim,klen", [("fp32", 50, 100, 37), ("fp32", 300, 128, 0), ("bf16", 512, 768, 0),
                                                  ("bf16", 256, 128, 12544), ("fp32", 1, 8, 0)])
def test_against_oracle_with_upstream_gradient(precision, n, dim, klen):
    """ragged sizes, the reference's queue length (config.yaml:16), and a non-unit upstream
    gradient (train() divides the loss before backward, src/train.py:145-147)."""
    g = torch.Generator().manual_seed(1337)
    q = torch.nn.functional.normalize(torch.randn(n, dim, generator=g), dim=1)
    k = torch.nn.functional.normalize(torch.randn(n, dim, generator=g) * 0.5 + q, dim=1)
    queue = torch.nn.functional.normalize(torch.randn(dim, klen, generator=g), dim=0) if klen else None
    loss, dq, dk = _run(q, k, queue, 0.05, precision, upstream=0.5)
    rl, rdq, rdk = infonce.nce_info_loss(q, k, queue, 0.05, dtype=torch.float64)
    tol, gtol = (1e-5, 1e-4) if precision == "fp32" else (2e-2, 3e-2)
    assert abs(loss.item() - rl.item()) <= tol * abs(rl.item()) + 1e-5
    scale = max(rdq.abs().max().item(), rdk.abs().max().item(), 1e-12) * 0.5
    # absolute floor: with N == 1 the only logit is the positive, loss and gradients are exactly 0
    assert (dq.double() - 0.5 * rdq).abs().max().item() <= gtol * scale + 1e-5
    assert (dk.double() - 0.5 * rdk).abs().max().item() <= gtol * scale + 1e-5


def test_config4_full_size_forward_backward():
    """BASELINE config 3 (batch 4096 x 768): loss vs the oracle in fp64 on the CPU (the reference
    needs ~3 s there), gradient checked on a row sample through linearity of the loss in 1/2 scale."""
    n, dim = 4096, 768
    g = torch.Generator().manual_seed(1337)
    q = torch.nn.functional.normalize(torch.randn(n, dim, generator=g), dim=1)
    k = torch.nn.functional.normalize(torch.randn(n, dim, generator=g) * 0.5 + q, dim=1)
    loss, dq, dk = _run(q, k, None, 0.05, "bf16")
    rl, rdq, rdk = infonce.nce_info_loss(q, k, None, 0.05, dtype=torch.float32)
    assert abs(loss.item() - rl.item()) <= 2e-2 * abs(rl.item())
    scale = rdq.abs().max().item()
    assert (dq - rdq).abs().max().item() <= 3e-2 * scale
    assert (dk - rdk).abs().max().item() <= 3e-2 * scale


def test_module_is_stateless_and_k_no_grad_ok():
    crit = drs_b200.NCELoss({"temperature": 0.05})
    assert len(crit.state_dict()) == 0
    g = torch.Generator().manual_seed(3)
    q = torch.nn.functional.normalize(torch.randn(64, 128, generator=g), dim=1).to(DEV).requires_grad_(True)
    k = torch.nn.functional.normalize(torch.randn(64, 128, generator=g), dim=1).to(DEV)       # momentum encoder: no grad
    loss = crit(q, k, None)
    loss.backward()
    assert q.grad is not None and torch.isfinite(q.grad).all() and loss.device.type == "cuda" and loss.dim() == 0
    with pytest.raises(NotImplementedError):
        crit(q, k, None, cluster_result={"emb2cluster": []}, index=torch.arange(64))
