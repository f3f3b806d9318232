#!/bin/bash
# launch list (device time per kernel) of one probe: tools/gpu_launch_list.sh <out-prefix> <python args...>
out=$1; shift
python "$@" > gpurun_out/${out}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${out}_launches.csv python "$@" > gpurun_out/${out}_ncu.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/${out}_launches.csv") if l.startswith('"'))]
h=rows[0]; ik,iv=h.index("Kernel Name"),h.index("Metric Value")
seq=[(r[ik].split("(")[0][-60:],float(r[iv].replace(",",""))) for r in rows[1:]]
# print the last 40 launches in order
for n,(k,v) in enumerate(seq[-40:]): print(f"{n:3d} {v/1e3:10.1f} us  {k}")
PY
