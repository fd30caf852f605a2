#!/bin/bash
# durations of the table-preparation kernels for 512 cosmologies (ncu launch list), per library variant
# usage (GPU box): bash tools/prep_time.sh [lib names...]
libs="$@"; [ -z "$libs" ] && libs="libccl_b200"
for v in $libs; do
  CCL_B200_LIB=$PWD/ccl_b200/$v.so ncu --metrics gpu__time_duration.sum --clock-control none -k regex:bicubic -c 3 --csv --log-file gpurun_out/pf_$v.csv python bench.py --ncosmo 512 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
  echo $v; grep bicubic gpurun_out/pf_$v.csv | awk -F'","' '{print "   " $5 "  " $NF}' | tr -d '"' | cut -c1-90
done
