# usage: bash tools/prof_qag.sh <tag> [fast|general]: ncu --set full of the 'qag_quad' kernel at 16 cosmologies
TAG=$1; MODE=${2:-fast}
if [ $MODE = fast ]; then unset CCL_B200_NO_QAG_FAST; K=limber_qag_fast_kernel; else export CCL_B200_NO_QAG_FAST=1; K=limber_qag_kernel; fi
ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -f -o gpurun_out/prof_$TAG \
  python bench.py --method qag_quad --ncosmo 16 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --from-parameters 0 --qag-sample 0 --no-y10 > gpurun_out/ncu_$TAG.log 2>&1
echo ncu rc=$?
