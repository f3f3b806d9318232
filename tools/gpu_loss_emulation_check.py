import sys
sys.path.insert(0, "/root/repo")
import torch, drs_b200 as drs
dev = torch.device("cuda:0")
unit = lambda x: torch.nn.functional.normalize(x, dim=1)
for (n, dim, klen, temp) in [(1024, 16, 64, 0.05), (1024, 16, 0, 0.05), (1024, 64, 64, 0.05), (256, 16, 64, 0.05)]:
    g = torch.Generator(device=dev).manual_seed(7)
    q = unit(torch.randn(n, dim, generator=g, device=dev)).requires_grad_(True)
    k = unit(torch.randn(n, dim, generator=g, device=dev) * 0.5 + q.detach()).requires_grad_(True)
    queue = torch.nn.functional.normalize(torch.randn(dim, klen, generator=g, device=dev), dim=0) if klen else None
    loss = drs.NCELoss({"temperature": temp, "precision": "bf16"})(q, k, queue)
    loss.backward()
    def ref(qv, kv, qu, round_h):
        qd, kd = qv.double().requires_grad_(True), kv.double().requires_grad_(True)
        f = torch.cat([qd, kd]); n2 = 2 * n
        sm = (f @ f.T) / temp
        sm = sm.masked_fill(torch.eye(n2, dtype=torch.bool, device=dev), float("-inf"))
        logits = sm if qu is None else torch.cat([sm, torch.cat([qd @ qu.double(), qd @ qu.double()]) / temp], 1)
        pos = torch.arange(n2, device=dev).roll(n)
        L = torch.logsumexp(logits, 1)
        l = (L - sm[torch.arange(n2, device=dev), pos]).sum() / 2
        if not round_h:
            l.backward(); return l.item(), qd.grad, kd.grad
        # emulate: H rounded to bf16, dF = H F (F bf16 values)
        with torch.no_grad():
            P = torch.exp(sm - L[:, None]); oh = torch.zeros_like(P); oh[torch.arange(n2, device=dev), pos] = 1
            H = ((P + P.T - 2 * oh) / (2 * temp)).float().bfloat16().double()
            dF = H @ f
            if qu is not None:
                lq = (qd @ qu.double()) / temp
                W = ((torch.exp(lq - L[:n, None]) + torch.exp(lq - L[n:, None])) / (2 * temp)).float().bfloat16().double()
                dF[:n] += W @ qu.double().T
        return l.item(), dF[:n], dF[n:]
    l0, dq0, dk0 = ref(q.detach(), k.detach(), queue, False)
    qb, kb = q.detach().bfloat16().float(), k.detach().bfloat16().float()
    qub = queue.bfloat16().float() if queue is not None else None
    l1, dq1, dk1 = ref(qb, kb, qub, True)
    def rowerr(a, b):
        return ((a.double() - b).norm(dim=1) / b.norm(dim=1).clamp_min(1e-30)).max().item()
    print(f"n={n} dim={dim} queue={klen}: vs fp64-of-fp32-inputs: loss rel {abs(loss.item()-l0)/abs(l0):.2e} row err dq {rowerr(q.grad,dq0):.3e} dk {rowerr(k.grad,dk0):.3e} | "
          f"vs bf16 emulation: loss rel {abs(loss.item()-l1)/abs(l1):.2e} row err dq {rowerr(q.grad,dq1):.3e} dk {rowerr(k.grad,dk1):.3e} | emulation vs exact: dq {rowerr(dq1,dq0):.3e}")
