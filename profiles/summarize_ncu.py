#!/usr/bin/env python
"""Turn an ncu report (.ncu-rep, read with `ncu -i ... --page raw --csv`) and a launch-list CSV
into the committed summaries of this directory.
    python profiles/summarize_ncu.py <prof.ncu-rep> <launches.csv> <out_prefix>"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__warps_issue_stalled_long_scoreboard_per_warp_active.pct",
]


def main():
    rep, launches, out = sys.argv[1:4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    summary = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP:
                d[h] = f"{v} {u}".strip()
        summary.append(d)
    with open(out + "_kernel_metrics.json", "w") as f:
        json.dump(summary, f, indent=1)
    # launch list: total device time per kernel name and its share of the profiled window
    lrows = list(csv.reader(l for l in open(launches) if l.startswith('"')))
    lh = lrows[0]
    ik, iv = lh.index("Kernel Name"), lh.index("Metric Value")
    tot, cnt = {}, {}
    for r in lrows[1:]:
        name = r[ik].split("(")[0]
        tot[name] = tot.get(name, 0.0) + float(r[iv].replace(",", ""))
        cnt[name] = cnt.get(name, 0) + 1
    whole = sum(tot.values())
    with open(out + "_launch_shares.md", "w") as f:
        f.write("| kernel | launches | total ms | avg ms | share of device time |\n|---|---|---|---|---|\n")
        for name, t in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write(f"| `{name}` | {cnt[name]} | {t / 1e6:.3f} | {t / 1e6 / cnt[name]:.4f} | {100 * t / whole:.3f} % |\n")
    print(open(out + "_launch_shares.md").read())
    print(json.dumps(summary, indent=1)[:3000])


if __name__ == "__main__":
    main()
