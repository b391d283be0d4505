#!/bin/bash
# launch list only (per-kernel durations) of the default bench workload
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline $*"
FILT='regex:update_|pml_|mur_|excite|probe|nf2ff|ts_add|energy'
mkdir -p gpurun_out
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$FILT" -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv
