#!/bin/bash
# Round-2 ncu recipe (/opt/skills/guides/B200_PROFILING.md): plain run first, then the launch list of every kernel of the
# path on three workloads, then one full capture of the dominant kernels.  Outputs under gpurun_out/ (copy into profiles/).
set -u
TAG=${1:-r02}
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
mkdir -p gpurun_out
for W in "patch100m" "cube --cube-n 768" "config1" "patch100m --seeded"; do
  N=$(echo $W | tr -d ' -' )
  timeout 600 python tools/prof_step.py --workload $W > gpurun_out/plain_${TAG}_$N.log 2>&1 || { echo "plain run failed: $W"; tail -5 gpurun_out/plain_${TAG}_$N.log; continue; }
  tail -1 gpurun_out/plain_${TAG}_$N.log
  timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/launches_${TAG}_$N.csv \
      python tools/prof_step.py --workload $W > gpurun_out/ncu_launches_${TAG}_$N.log 2>&1
  python tools/summarize_launches.py gpurun_out/launches_${TAG}_$N.csv | tee gpurun_out/launches_${TAG}_$N.txt
done
# full captures: the fused H->E launch (compressed operator), the plain fusion on the seeded operator, one x-slab launch
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:update_he6_kernel -c 1 -f \
    -o gpurun_out/prof_${TAG}_he6 python tools/prof_step.py --workload patch100m > gpurun_out/ncu_${TAG}_he6.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:update_he_kernel -c 1 -f \
    -o gpurun_out/prof_${TAG}_he_seeded python tools/prof_step.py --workload patch100m --seeded > gpurun_out/ncu_${TAG}_he_seeded.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:update_he6_kernel -c 1 -f \
    -o gpurun_out/prof_${TAG}_he6_cube python tools/prof_step.py --workload cube --cube-n 768 > gpurun_out/ncu_${TAG}_he6_cube.log 2>&1
for R in he6 he_seeded he6_cube; do
  ncu -i gpurun_out/prof_${TAG}_$R.ncu-rep --page raw --csv > gpurun_out/ncu_raw_${TAG}_$R.csv 2>/dev/null
done
ls -la gpurun_out | grep -E "${TAG}.*(ncu-rep|csv)"
# one merged slab launch of each pass
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:update_slabs_kernel -c 2 -f \
    -o gpurun_out/prof_${TAG}_slabs python tools/prof_step.py --workload patch100m > gpurun_out/ncu_${TAG}_slabs.log 2>&1
ncu -i gpurun_out/prof_${TAG}_slabs.ncu-rep --page raw --csv > gpurun_out/ncu_raw_${TAG}_slabs.csv 2>/dev/null
