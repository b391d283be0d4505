#!/bin/bash
# ncu recipe of /opt/skills/guides/B200_PROFILING.md: plain run first, then the launch list, then full captures.
#   tools/profile.sh [tag]     (outputs under gpurun_out/, copy the summaries you want judged into profiles/)
set -u
TAG=${1:-r01b}
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
FILT='regex:update_|pml_k|mur_|excite|probe|nf2ff|ts_add|energy'
mkdir -p gpurun_out
timeout 600 $CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$FILT" -c 400 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_$TAG.csv
# the dominant kernel (fused H->E launch over the plain region) and one launch of each slab kind
timeout 900 ncu --set full --clock-control none --import-source on -k regex:update_he5_kernel -s 2 -c 1 -f -o gpurun_out/prof_he5_$TAG $CMD > gpurun_out/ncu_he5_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"update_[eh]_kernel<\(int\)4, \(int\)[12]" -s 10 -c 5 -f \
    -o gpurun_out/prof_slabs_$TAG $CMD > gpurun_out/ncu_slabs_$TAG.log 2>&1
ls -la gpurun_out | grep -E "ncu-rep|csv"
