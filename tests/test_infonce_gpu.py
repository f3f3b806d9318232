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


def _assert_rows_close(got, ref, precision):
    """Every gradient ROW on its own scale (a max-abs bar over the matrix would let small rows be arbitrarily wrong):
    relative L2 error <= 3e-2 and cosine >= 0.9995 on the bf16 path (a CPU emulation of the bf16 roundings gives
    <= 1.8e-2 / >= 0.99995 on these fixtures), 2e-4 / 0.999999 on the fp32 path.  Rows whose reference norm is below
    1e-6 of the largest row are compared on that floor."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    rn = np.linalg.norm(ref, axis=1)
    floor = max(rn.max(), 1e-30) * 1e-6
    rel = np.linalg.norm(got - ref, axis=1) / np.maximum(rn, floor)
    cos = (got * ref).sum(1) / np.maximum(np.linalg.norm(got, axis=1) * rn, 1e-300)
    rtol, ctol = (3e-2, 0.9995) if precision == "bf16" else (2e-4, 0.999999)
    assert rel.max() <= rtol, f"worst row: relative L2 error {rel.max():.3e} at row {rel.argmax()}"
    live = rn > floor
    assert cos[live].min() >= ctol, f"worst row: cosine {cos[live].min():.6f} at row {np.flatnonzero(live)[cos[live].argmin()]}"


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
    _assert_rows_close(dq.numpy(), z["dq"], "fp32")
    _assert_rows_close(dk.numpy(), z["dk"], "fp32")


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
    _assert_rows_close(dq.numpy(), z["dq"], "bf16")
    _assert_rows_close(dk.numpy(), z["dk"], "bf16")


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
    _assert_rows_close(dq.numpy(), rdq.numpy(), "bf16")
    _assert_rows_close(dk.numpy(), rdk.numpy(), "bf16")


def test_batch_8192_beyond_one_pass_of_partials():
    """2N = 16 384 rows: 64 tile rows -- more row slots (128) and column slots (up to 504) than one batch of loads of
    the symmetric forward's row kernel covers.  Loss and a sample of gradient rows against the closed form in float64
    on the GPU, fed the same bf16-rounded embeddings."""
    n, dim, temp = 8192, 128, 0.05
    g = torch.Generator(device=DEV).manual_seed(77)
    q = torch.nn.functional.normalize(torch.randn(n, dim, generator=g, device=DEV), dim=1).bfloat16().float().requires_grad_(True)
    k = torch.nn.functional.normalize(torch.randn(n, dim, generator=g, device=DEV) * 0.5 + q.detach(), dim=1).bfloat16().float().requires_grad_(True)
    loss = drs_b200.NCELoss({"temperature": temp, "precision": "bf16"})(q, k, None)
    loss.backward()
    f = torch.cat([q.detach(), k.detach()]).double()
    lse = torch.empty(2 * n, dtype=torch.float64, device=DEV)
    for r0 in range(0, 2 * n, 2048):                      # row blocks: the 16 384^2 float64 matrix never exists
        sm = (f[r0:r0 + 2048] @ f.T) / temp
        sm[torch.arange(2048, device=DEV), torch.arange(r0, r0 + 2048, device=DEV)] = float("-inf")
        lse[r0:r0 + 2048] = torch.logsumexp(sm, 1)
    pos = (f[:n] * f[n:]).sum(1) / temp
    ref = (lse.sum() - 2 * pos.sum()) / 2
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item())
    rows = torch.tensor([0, 1, 255, 256, 4095, 4096, 8191], device=DEV)      # dq rows: d/dq_i = sum_j H_ij f_j over both halves
    p_rows = torch.exp((f[rows] @ f.T) / temp - lse[rows, None])             # P[i, :]
    p_cols = torch.exp((f[rows] @ f.T) / temp - lse[None, :])                # P[:, i] (S symmetric)
    h = p_rows + p_cols
    h[torch.arange(len(rows), device=DEV), rows] = 0.0
    h[torch.arange(len(rows), device=DEV), rows + n] -= 2.0
    ref_dq = (h @ f) / (2 * temp)
    got = q.grad[rows].double()
    rel = (got - ref_dq).norm(dim=1) / ref_dq.norm(dim=1)
    assert rel.max().item() <= 3e-2, rel


def test_module_is_stateless_and_k_no_grad_ok():
    crit = drs_b200.NCELoss({"temperature": 0.05})
    assert len(crit.state_dict()) == 0
    g = torch.Generator().manual_seed(3)
    q = torch.nn.functional.normalize(torch.randn(64, 128, generator=g), dim=1).to(DEV).requires_grad_(True)
    k = torch.nn.functional.normalize(torch.randn(64, 128, generator=g), dim=1).to(DEV)       # momentum encoder: no grad
    loss = crit(q, k, None)
    loss.backward()
    assert q.grad is not None and torch.isfinite(q.grad).all() and loss.device.type == "cuda" and loss.dim() == 0


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "moco_*.npz"))))
def test_moco_infonce_matches_reference_golden(path, precision):
    """InfoNCE.forward (contrastive_loss.py:26-44) vs the reference's own outputs."""
    z = np.load(path)
    q = torch.from_numpy(z["q"]).to(DEV).requires_grad_(True)
    k = torch.from_numpy(z["k"]).to(DEV).requires_grad_(True)
    crit = drs_b200.InfoNCE({"temperature": float(z["temperature"]), "precision": precision})
    loss = crit(q, k, torch.from_numpy(z["queue"]).to(DEV))
    loss.backward()
    tol, gtol = (1e-5, 1e-4) if precision == "fp32" else (2e-2, 3e-2)
    assert abs(loss.item() - float(z["loss"])) <= tol * abs(float(z["loss"])) + 1e-6
    scale = max(np.abs(z["dq"]).max(), np.abs(z["dk"]).max())
    np.testing.assert_allclose(q.grad.cpu().numpy(), z["dq"], rtol=0, atol=gtol * scale)
    np.testing.assert_allclose(k.grad.cpu().numpy(), z["dk"], rtol=0, atol=gtol * scale)
    _assert_rows_close(q.grad.cpu().numpy(), z["dq"], precision)
    _assert_rows_close(k.grad.cpu().numpy(), z["dk"], precision)


def _proto_inputs(z):
    """the selection of contrastive_loss.py:101-112,122-123 with make_golden.py's fixed sampler"""
    import sys
    sys.path.insert(0, GOLDEN)
    from proto_inputs import selected_from_fixture
    protos, temps = selected_from_fixture(z)
    return [torch.from_numpy(p).to(DEV) for p in protos], [torch.from_numpy(t).to(DEV) for t in temps]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "proto_*.npz"))))
def test_proto_nce_matches_reference_golden(path, precision):
    """NCELoss._compute_proto_loss (contrastive_loss.py:112-134) vs the reference's own outputs.  Prototype counts
    22 and 84 (ragged), 88, and the reference's own 128 + 3072 = 3200 (config.yaml:29-30,87), all on both paths:
    the tcgen05 path takes any count (padded pitch + tensor maps that stop at the true extent)."""
    z = np.load(path)
    protos, temps = _proto_inputs(z)
    q = torch.from_numpy(z["q"]).to(DEV).requires_grad_(True)
    loss = drs_b200.proto_nce_loss(q, protos, temps, precision)
    (loss * 2.0).backward()
    tol, gtol = (1e-5, 1e-4) if precision == "fp32" else (2e-2, 3e-2)
    assert abs(loss.item() - float(z["loss"])) <= tol * abs(float(z["loss"])) + 1e-5
    scale = np.abs(z["dq"]).max() * 2.0
    np.testing.assert_allclose(q.grad.cpu().numpy(), 2.0 * z["dq"], rtol=0, atol=gtol * scale)
    _assert_rows_close(q.grad.cpu().numpy(), 2.0 * z["dq"], precision)


def test_nceloss_with_cluster_result_adds_proto_loss():
    """NCELoss.forward with cluster_result (contrastive_loss.py:137-141): info loss + proto loss on
    the module's own prototype selection, checked against the oracle on that same selection."""
    g = torch.Generator().manual_seed(5)
    n, dim, ncl, r = 64, 128, [96, 160], 24
    q = torch.nn.functional.normalize(torch.randn(n, dim, generator=g), dim=1)
    k = torch.nn.functional.normalize(torch.randn(n, dim, generator=g) * 0.5 + q, dim=1)
    index = torch.randperm(4 * n, generator=g)[:n]
    cr = {"emb2cluster": [], "centroids": [], "density": []}
    for c in ncl:
        e2c = torch.randint(0, c, (4 * n,), generator=g)
        e2c[0] = c - 1
        cr["emb2cluster"].append(e2c)
        cr["centroids"].append(torch.nn.functional.normalize(torch.randn(c, dim, generator=g), dim=1).to(DEV))
        cr["density"].append((torch.rand(c, generator=g) * 0.1 + 0.02).to(DEV))
    cfg = {"temperature": 0.05, "precision": "fp32", "cluster": {"num_cluster": ncl, "num_neg_proto": r}}
    crit = drs_b200.NCELoss(cfg)
    qd = q.to(DEV).requires_grad_(True)
    import random
    random.seed(1126)                                              # the negatives come from the global stream (:4,:109)
    loss = crit(qd, k.to(DEV), None, cluster_result=cr, index=index)
    loss.backward()
    twin = drs_b200.NCELoss(cfg)
    random.seed(1126)                                              # same seed -> same selection
    protos, temps = twin.select_prototypes(cr, index)
    assert all(p.shape == (n + r, dim) for p in protos)
    l_info, dq_info, _ = infonce.nce_info_loss(q, k, None, 0.05, dtype=torch.float64)
    l_proto, dq_proto = infonce.proto_loss(q, [p.cpu() for p in protos], [t.cpu() for t in temps], dtype=torch.float64)
    ref = (l_info + l_proto).item()
    assert abs(loss.item() - ref) <= 1e-5 * abs(ref)
    rdq = dq_info + dq_proto
    assert (qd.grad.cpu().double() - rdq).abs().max().item() <= 1e-4 * rdq.abs().max().item()


@pytest.mark.parametrize("n,dim", [(1024, 128), (512, 768), (2048, 64)])
def test_symmetric_gradient_matrix_matches_the_full_computation(n, dim):
    """Backward at whole 256-row tiles computes only the blocks of H = dL/dS on and above the diagonal.  Mode 2 (default)
    stores them once and the dF = H F GEMM reads the blocks below the diagonal TRANSPOSED out of the stored ones
    (MN-major tcgen05 operand); mode 1 writes every block twice (itself and its transpose); mode 0 computes all of H.
    Modes 1 and 2 must give the same gradients (same bf16 H values, same accumulation order); mode 0 evaluates the
    blocks below the diagonal from their own rows' side (2^(y - L_j) (c + u_j v_i) instead of the transpose of
    2^(y - L_i) (c + u_i v_j), GradLogitEpilogue), so single bf16 values of H may round the other way: equal to a few
    1e-4 of the gradient scale.  All match the oracle."""
    g = torch.Generator().manual_seed(5)
    q = torch.nn.functional.normalize(torch.randn(n, dim, generator=g), dim=1)
    k = torch.nn.functional.normalize(torch.randn(n, dim, generator=g) * 0.5 + q, dim=1)
    res = {}
    try:
        for mode in (2, 1, 0):
            drs_b200.set_option("tune.symmetric_grad", mode)
            res[mode] = _run(q, k, None, 0.05, "bf16")
    finally:
        drs_b200.set_option("tune.symmetric_grad", 2)
    loss0, dq0, dk0 = res[0]
    scale = dq0.abs().max().item()
    for mode in (1, 2):
        loss, dq, dk = res[mode]
        assert loss.item() == loss0.item()
        assert (dq - dq0).abs().max().item() <= 1e-3 * scale and (dk - dk0).abs().max().item() <= 1e-3 * scale, mode
    assert (res[1][1] - res[2][1]).abs().max().item() <= 1e-6 * scale and (res[1][2] - res[2][2]).abs().max().item() <= 1e-6 * scale
    rl, rdq, rdk = infonce.nce_info_loss(q, k, None, 0.05, dtype=torch.float64)
    assert (res[2][1].double() - rdq).abs().max().item() <= 3e-2 * rdq.abs().max().item()
    assert (res[2][2].double() - rdk).abs().max().item() <= 3e-2 * rdk.abs().max().item()
    _assert_rows_close(res[2][1].numpy(), rdq.numpy(), "bf16")
    _assert_rows_close(res[2][2].numpy(), rdk.numpy(), "bf16")


@pytest.mark.parametrize("temp,spread,klen", [(0.05, 0.0, 0), (0.05, 0.0, 256), (0.05, 0.6, 256), (0.01, 0.6, 256), (0.004, 0.8, 0)])
def test_shared_exponential_form_and_its_fallback(temp, spread, klen):
    """GradLogitEpilogue evaluates 2^(y - L_i) once per score and derives the second softmax term from it through
    per-row / per-column factors while the LSEs lie within 2^30 of a reference row, and falls back to two exponentials per
    score elsewhere.  Rows of very different norms and temperatures far below the reference's 0.05 put warps and
    column chunks on both sides of that bound: the result must equal the two-exponential form (debug.flags = 16, the
    formulation the goldens pinned) to bf16 rounding of single H values, and match the oracle fed the same
    bf16-rounded embeddings."""
    n, dim = 512, 128
    g = torch.Generator().manual_seed(23)
    scale_rows = 1.0 + spread * (2.0 * torch.rand(n, 1, generator=g) - 1.0)
    q = torch.nn.functional.normalize(torch.randn(n, dim, generator=g), dim=1) * scale_rows
    k = torch.nn.functional.normalize(torch.randn(n, dim, generator=g) * 0.5 + q, dim=1) * scale_rows.flip(0)
    queue = torch.nn.functional.normalize(torch.randn(dim, klen, generator=g), dim=0) if klen else None
    q, k = q.bfloat16().float(), k.bfloat16().float()
    queue = queue.bfloat16().float() if klen else None
    try:
        drs_b200.set_option("debug.flags", 16)
        loss_two, dq_two, dk_two = _run(q, k, queue, temp, "bf16")
    finally:
        drs_b200.set_option("debug.flags", 0)
    loss, dq, dk = _run(q, k, queue, temp, "bf16")
    assert loss.item() == loss_two.item()
    scale = max(dq_two.abs().max().item(), dk_two.abs().max().item())
    assert torch.isfinite(dq).all() and torch.isfinite(dk).all()
    assert (dq - dq_two).abs().max().item() <= 2e-3 * scale and (dk - dk_two).abs().max().item() <= 2e-3 * scale
    rl, rdq, rdk = infonce.nce_info_loss(q, k, queue, temp, dtype=torch.float64)
    assert abs(loss.item() - rl.item()) <= 1e-3 * abs(rl.item())
    rscale = max(rdq.abs().max().item(), rdk.abs().max().item())
    assert (dq.double() - rdq).abs().max().item() <= 3e-2 * rscale and (dk.double() - rdk).abs().max().item() <= 3e-2 * rscale
    if temp >= 0.01:   # (at T = 0.004 single rows of the bf16 path -- either form -- are off by tens of per cent: H in bf16)
        _assert_rows_close(dq.numpy(), rdq.numpy(), "bf16")
        _assert_rows_close(dk.numpy(), rdk.numpy(), "bf16")


@pytest.mark.parametrize("n,dim,klen,temp", [(128, 128, 512, 0.05), (512, 64, 0, 0.05), (1024, 128, 256, 0.07), (2048, 64, 0, 0.05),
                                             (512, 128, 0, 0.01)])
def test_symmetric_forward_matches_the_full_matrix_forward(n, dim, klen, temp):
    """Forward over the tiles of F F^T on and above the diagonal only (SymLseEpilogue: row sums and, through a
    transposing warp butterfly, column sums against one bounded reference) against the full-matrix forward with
    running maxima: same loss to fp32 summation order, same gradients (they depend on the LSEs), both equal to the
    oracle.  Forced on (tune.symmetric_lse = 2) because by default it is used only where it pays (2N >= 6144 at dim = 768).  T = 0.01 makes the
    logits' span exceed the bound: the symmetric launch then walks the full matrix itself."""
    g = torch.Generator().manual_seed(31)
    q = torch.nn.functional.normalize(torch.randn(n, dim, generator=g), dim=1)
    k = torch.nn.functional.normalize(torch.randn(n, dim, generator=g) * 0.5 + q, dim=1)
    queue = torch.nn.functional.normalize(torch.randn(dim, klen, generator=g), dim=0) if klen else None
    res = {}
    try:
        for mode in (2, 0):
            drs_b200.set_option("tune.symmetric_lse", mode)
            res[mode] = _run(q, k, queue, temp, "bf16")
    finally:
        drs_b200.set_option("tune.symmetric_lse", 1)
    (loss_s, dq_s, dk_s), (loss_f, dq_f, dk_f) = res[2], res[0]
    assert abs(loss_s.item() - loss_f.item()) <= 2e-6 * abs(loss_f.item())
    scale = max(dq_f.abs().max().item(), dk_f.abs().max().item())
    assert (dq_s - dq_f).abs().max().item() <= 2e-3 * scale and (dk_s - dk_f).abs().max().item() <= 2e-3 * scale
    qb, kb = q.bfloat16().float(), k.bfloat16().float()
    rl, rdq, rdk = infonce.nce_info_loss(qb, kb, queue.bfloat16().float() if klen else None, temp, dtype=torch.float64)
    assert abs(loss_s.item() - rl.item()) <= 1e-4 * abs(rl.item())
    _assert_rows_close(dq_s.numpy(), rdq.numpy(), "bf16")
    _assert_rows_close(dk_s.numpy(), rdk.numpy(), "bf16")


def test_round_robin_tile_order_gives_the_same_bits():
    """tune.triangle_order = 1 deals the tiles of the symmetric GEMMs round-robin instead of in contiguous pieces: which
    cluster computes a tile must not matter (partials are placed by tile coordinates, sums run in a fixed order)."""
    n, dim = 1024, 128
    g = torch.Generator().manual_seed(41)
    q = torch.nn.functional.normalize(torch.randn(n, dim, generator=g), dim=1)
    k = torch.nn.functional.normalize(torch.randn(n, dim, generator=g) * 0.5 + q, dim=1)
    res = {}
    try:
        drs_b200.set_option("tune.symmetric_lse", 2)
        for order in (0, 1):
            drs_b200.set_option("tune.triangle_order", order)
            res[order] = _run(q, k, None, 0.05, "bf16")
    finally:
        drs_b200.set_option("tune.triangle_order", 0)
        drs_b200.set_option("tune.symmetric_lse", 1)
    assert res[0][0].item() == res[1][0].item()
    assert torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])


def test_loss_step_is_cuda_graph_capturable():
    """Forward + backward of NCELoss captured once in a CUDA graph and replayed on new embeddings: same loss and
    gradients as the eager call (a trainer can take the ~20 small launches off the host)."""
    n, dim = 512, 128
    g = torch.Generator(device=DEV).manual_seed(11)
    q = torch.nn.functional.normalize(torch.randn(n, dim, generator=g, device=DEV), dim=1).requires_grad_(True)
    k = torch.nn.functional.normalize(torch.randn(n, dim, generator=g, device=DEV), dim=1).requires_grad_(True)
    crit = drs_b200.NCELoss({"temperature": 0.05})
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            q.grad = k.grad = None
            crit(q, k, None).backward()
    torch.cuda.current_stream().wait_stream(side)
    q.grad = k.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        loss = crit(q, k, None)
        loss.backward()
    for seed in (1, 2):
        gg = torch.Generator(device=DEV).manual_seed(seed)
        with torch.no_grad():
            q.copy_(torch.nn.functional.normalize(torch.randn(n, dim, generator=gg, device=DEV), dim=1))
            k.copy_(torch.nn.functional.normalize(torch.randn(n, dim, generator=gg, device=DEV) * 0.3 + q, dim=1))
        graph.replay()
        torch.cuda.synchronize()
        got = (loss.item(), q.grad.clone(), k.grad.clone())
        q2, k2 = q.detach().clone().requires_grad_(True), k.detach().clone().requires_grad_(True)
        l2 = crit(q2, k2, None)
        l2.backward()
        assert got[0] == l2.item() and torch.equal(got[1], q2.grad) and torch.equal(got[2], k2.grad)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_split_k_queue_gradient_at_the_reference_shapes(precision):
    """The reference's own shapes (config.yaml: batch 128, 128-d, queue 12 544): dq through the queue is a
    [128 x 12544] x [12544 x 128] product -- one output tile, long K -- computed as K slices on many clusters and
    summed in slice order.  Same gradients as the unsplit product (tune.k_split = 1), bit-identical from run to run,
    and both within the usual bars of the oracle."""
    n, dim, klen = 128, 128, 12544
    g = torch.Generator().manual_seed(7)
    q = torch.nn.functional.normalize(torch.randn(n, dim, generator=g), dim=1)
    k = torch.nn.functional.normalize(torch.randn(n, dim, generator=g) * 0.5 + q, dim=1)
    queue = torch.nn.functional.normalize(torch.randn(dim, klen, generator=g), dim=0)
    rl, rdq, rdk = infonce.nce_info_loss(q, k, queue, 0.05, dtype=torch.float64)
    l1, dq1, dk1 = _run(q, k, queue, 0.05, precision)
    l2, dq2, dk2 = _run(q, k, queue, 0.05, precision)
    assert torch.equal(dq1, dq2) and torch.equal(dk1, dk2) and torch.equal(l1, l2)
    drs_b200.set_option("tune.k_split", 1)
    try:
        l0, dq0, dk0 = _run(q, k, queue, 0.05, precision)
    finally:
        drs_b200.set_option("tune.k_split", 0)
    assert torch.equal(l0, l1) and torch.equal(dk0, dk1)       # the split only touches dq
    torch.testing.assert_close(dq1, dq0, rtol=1e-4, atol=1e-6 * float(dq0.abs().max()))
    _assert_rows_close(dq1.numpy(), rdq.numpy(), precision)
    _assert_rows_close(dk1.numpy(), rdk.numpy(), precision)
