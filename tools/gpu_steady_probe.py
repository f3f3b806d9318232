#!/usr/bin/env python
"""Perf probe (test tooling): does a short timing window see the steady state?  Runs ONE claim count back to back for a
few seconds and prints the scan time, SM clock and board power over time.
    python tools/gpu_steady_probe.py [corpus rows] [csv of claim counts] [seconds per count] [cta group or 0]
Each count starts from a cooled-down chip (1 s idle)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import drs_b200 as drs  # noqa: E402


def main():
    import pynvml as nv
    nv.nvmlInit()
    h = nv.nvmlDeviceGetHandleByIndex(0)
    dev = torch.device("cuda:0")
    nc = int(sys.argv[1]) if len(sys.argv) > 1 else 25_000_000
    nqs = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "64,128,256,512").split(",")]
    secs = float(sys.argv[3]) if len(sys.argv) > 3 else 3.0
    cg = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    g = torch.Generator(device=dev).manual_seed(1)
    c = torch.empty(nc, 768, dtype=torch.bfloat16, device=dev)
    for r0 in range(0, nc, 1 << 20):
        r1 = min(nc, r0 + (1 << 20))
        c[r0:r1] = torch.nn.functional.normalize(torch.randn(r1 - r0, 768, generator=g, device=dev), dim=1)
    qa = torch.nn.functional.normalize(torch.randn(max(nqs), 768, generator=g, device=dev), dim=1).bfloat16()
    drs.set_option("search.cta_group", cg)
    if os.environ.get("READ_BASELINE"):
        # a plain read-only pass over the same bytes (torch reduction), for the power/clock comparison
        ci = c.view(torch.int32)
        time.sleep(1.0)
        t_start = time.perf_counter()
        marks = []
        while time.perf_counter() - t_start < secs:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(8):
                ci.max()
            e1.record()
            mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
            torch.cuda.synchronize()
            marks.append((time.perf_counter() - t_start, e0.elapsed_time(e1) / 8, mhz, pw))
        pick = [marks[0], marks[len(marks) // 4], marks[len(marks) // 2], marks[-1]]
        print("torch max() over the corpus bytes | " + "  ".join(f"t={t:4.2f}s {ms:6.3f} ms {mhz:4d} MHz {pw:4.0f} W" for t, ms, mhz, pw in pick), flush=True)
    t_hbm = nc * 768 * 2 / 6.551e12 * 1e3
    for nq in nqs:
        q = qa[:nq].contiguous()
        t_mma = 2.0 * nq * nc * 768 / 1389.5e12 * 1e3
        roof = max(t_hbm, t_mma)
        drs.search(q, c, 10)
        torch.cuda.synchronize()
        time.sleep(1.0)
        t_start = time.perf_counter()
        marks = []
        while time.perf_counter() - t_start < secs:
            prof = []
            for _ in range(8):
                drs.search(q, c, 10, profile=prof)
            mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)          # sampled while the 8 scans are in flight
            pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
            torch.cuda.synchronize()
            ms = sum(a.elapsed_time(b) for a, b in prof) / len(prof)
            marks.append((time.perf_counter() - t_start, ms, mhz, pw))
        pick = [marks[0], marks[len(marks) // 8], marks[len(marks) // 4], marks[len(marks) // 2], marks[-1]]
        tail = marks[len(marks) // 2:]
        steady = sum(m[1] for m in tail) / len(tail)
        print(f"nq={nq:5d} roof {roof:6.3f} ms | " + "  ".join(f"t={t:4.2f}s {ms:6.3f} ms {mhz:4d} MHz {pw:4.0f} W" for t, ms, mhz, pw in pick)
              + f" | steady {steady:6.3f} ms = {roof / steady * 100:5.1f} % of roof", flush=True)


if __name__ == "__main__":
    main()
