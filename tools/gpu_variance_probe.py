import os, sys
sys.path.insert(0, "/root/repo")
import torch, drs_b200 as drs
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
nc = 12_000_000
c = torch.empty(nc, 768, dtype=torch.bfloat16, device=dev)
for r0 in range(0, nc, 1 << 20):
    r1 = min(nc, r0 + (1 << 20))
    c[r0:r1] = torch.nn.functional.normalize(torch.randn(r1 - r0, 768, generator=g, device=dev), dim=1)
qa = torch.nn.functional.normalize(torch.randn(2048, 768, generator=g, device=dev), dim=1).bfloat16()
for nq in (512, 256, 1024, 512):
    q = qa[:nq].contiguous()
    prof = []
    for _ in range(40):
        drs.search(q, c, 10, profile=prof)
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in prof]
    clk = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
    pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000
    print(nq, "first5", [round(x, 2) for x in ms[:5]], "last5", [round(x, 2) for x in ms[-5:]], "min", round(min(ms), 2), "max", round(max(ms), 2), "clk", clk, "W", pw, flush=True)
