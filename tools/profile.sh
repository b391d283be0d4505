#!/bin/bash
# ncu recipe of /opt/skills/guides/B200_PROFILING.md: plain run first, then the launch list, then full captures.
set -u
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
FILT='regex:update_|pml_|mur_|excite|probe|nf2ff|ts_add|energy'
mkdir -p gpurun_out
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$FILT" -c 300 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv
# plain E/H launches are the ones with the big grid: skip the warm-up step launches, take 2
ncu --set full --clock-control none --import-source on -k regex:"update_e_kernel.*false, true|update_e_kernel<4, 0, 1>" -s 4 -c 2 -f -o gpurun_out/prof_update_e $CMD > gpurun_out/ncu_e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"update_h_kernel<4, 0, 1>" -s 4 -c 2 -f -o gpurun_out/prof_update_h $CMD > gpurun_out/ncu_h.log 2>&1
ls -la gpurun_out | grep -E "ncu-rep|csv"
