#!/bin/bash
# full ncu capture of the plain (non-PML, row-compressed) E and H volume launches of the default bench workload
set -u
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:update_e_kernel<\(int\)4, \(bool\)0, \(bool\)1>' -s 3 -c 2 -f -o gpurun_out/prof_update_e $CMD > gpurun_out/ncu_e.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:update_h_kernel<\(int\)4, \(bool\)0, \(bool\)1>' -s 3 -c 2 -f -o gpurun_out/prof_update_h $CMD > gpurun_out/ncu_h.log 2>&1
tail -3 gpurun_out/ncu_e.log
ls -la gpurun_out | grep -E "ncu-rep"
