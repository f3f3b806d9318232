#!/bin/bash
# One `ncu --set full` capture on the GPU box, brought back as CSV (the .ncu-rep files are too big for gpurun_out):
#   tools/ncu_capture.sh <name> <skip> <count> <python args...>
# writes gpurun_out/<name>_plain.log, gpurun_out/<name>_raw.csv (all metrics per captured launch) and
# gpurun_out/<name>_source.csv (per-SASS-instruction samples / stalls).  Runs the same command without ncu first.
name=$1; skip=$2; count=$3; shift 3
python "$@" > gpurun_out/${name}_plain.log 2>&1 || { echo "plain run failed"; tail -n 5 gpurun_out/${name}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:gemm_nt_tc -s $skip -c $count -o /tmp/${name} python "$@" > gpurun_out/${name}_ncu.log 2>&1
ncu -i /tmp/${name}.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
ncu -i /tmp/${name}.ncu-rep --page source --csv --print-source sass > gpurun_out/${name}_source.csv 2>/dev/null
tail -n 1 gpurun_out/${name}_plain.log
ls -la gpurun_out/${name}_raw.csv gpurun_out/${name}_source.csv | awk '{print $5, $9}'
