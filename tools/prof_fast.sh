# usage: bash tools/prof_fast.sh <tag> [lib name]: ncu --set full of the fast spline kernel at 64 cosmologies
TAG=$1; LIB=${2:-libccl_b200}
CCL_B200_LIB=$PWD/ccl_b200/$LIB.so ncu --set full --clock-control none --import-source on -k regex:limber_spline_fast_kernel -c 1 -f \
  -o gpurun_out/prof_$TAG python bench.py --ncosmo 64 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --from-parameters 0 --qag-sample 0 --no-y10 > gpurun_out/ncu_$TAG.log 2>&1
echo ncu rc=$?
