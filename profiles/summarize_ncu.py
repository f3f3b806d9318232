#!/usr/bin/env python
"""Turn ncu exports into the committed summaries of this directory.
    python profiles/summarize_ncu.py raw <x_raw.csv> <out_kernel_metrics.json>       `ncu -i rep --page raw --csv`
    python profiles/summarize_ncu.py source <x_source.csv> <out_hotspots.md> [kernel index]   `--page source --csv --print-source sass`
    python profiles/summarize_ncu.py launches <launches.csv> <out_launch_shares.md>  `ncu --metrics gpu__time_duration.sum --csv`
(the .ncu-rep files themselves are too large to bring back from the GPU box; tools/ncu_capture.sh exports these CSVs there)"""
import csv
import json
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
]


def raw(path, out):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    summary = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP:
                d[h] = f"{v} {u}".strip()
        summary.append(d)
    json.dump(summary, open(out, "w"), indent=1)
    print(json.dumps(summary, indent=1)[:2500])


def source(path, out, which=0):
    rows = list(csv.reader(open(path)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    a = starts[which]
    b = starts[which + 1] if which + 1 < len(starts) else len(rows)
    hdr = rows[a + 1]
    ix = {h: j for j, h in enumerate(hdr)}
    rs = [r for r in rows[a + 2:b] if len(r) == len(hdr)]
    samp = sum(int(r[ix["# Samples"]]) for r in rs) or 1
    inst = sum(int(r[ix["Instructions Executed"]]) for r in rs) or 1
    stall = [h for h in hdr if h.startswith("stall_") and "(Not" not in h]
    tot = {h: sum(int(r[ix[h]] or 0) for r in rs) for h in stall}
    with open(out, "w") as f:
        f.write(f"kernel: `{rows[a][1][:160]}`\n\nwarp-instructions executed: {inst}, stall samples: {samp}\n\n")
        f.write("stall reasons (share of samples): " + ", ".join(f"{h[6:]} {100 * v / samp:.1f} %" for h, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v) + "\n\n")
        f.write("| SASS line | samples | executed | instruction | stalls |\n|---|---|---|---|---|\n")
        top = sorted(range(len(rs)), key=lambda n: -int(rs[n][ix["# Samples"]]))[:24]
        for n in sorted(top):
            r = rs[n]
            st = ", ".join(f"{h[6:]} {r[ix[h]]}" for h in stall if int(r[ix[h]] or 0) > 0)
            f.write(f"| {n} | {100 * int(r[ix['# Samples']]) / samp:.2f} % | {r[ix['Instructions Executed']]} | `{r[1].strip()[:90]}` | {st} |\n")
    print(open(out).read()[:3000])


def launches(path, out):
    lrows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    lh = lrows[0]
    ik, iv = lh.index("Kernel Name"), lh.index("Metric Value")
    tot, cnt = {}, {}
    for r in lrows[1:]:
        name = r[ik].split("(")[0]
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        tot[name] = tot.get(name, 0.0) + v
        cnt[name] = cnt.get(name, 0) + 1
    whole = sum(tot.values())
    with open(out, "w") as f:
        f.write("| kernel | launches | total ms | avg ms | share of device time |\n|---|---|---|---|---|\n")
        for name, t in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write(f"| `{name}` | {cnt[name]} | {t / 1e6:.3f} | {t / 1e6 / cnt[name]:.4f} | {100 * t / whole:.3f} % |\n")
    print(open(out).read())


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "raw":
        raw(sys.argv[2], sys.argv[3])
    elif mode == "source":
        source(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 0)
    else:
        launches(sys.argv[2], sys.argv[3])
