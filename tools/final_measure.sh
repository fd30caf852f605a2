#!/bin/bash
# Round-end measurement batch on one B200 (run through gpurun); outputs under gpurun_out/.
# usage: bash tools/final_measure.sh [tag]
TAG=${1:-final}
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/pytest_gpu_$TAG.txt; cat $O/pytest_gpu_$TAG.txt
python __graft_entry__.py --smoke > $O/smoke_$TAG.txt 2>&1; tail -2 $O/smoke_$TAG.txt
# the driver's two arms, default flags (stated config: 4096 LHS cosmologies, device-built tables; legs: e2e,
# from_parameters, qag, y10, cpu_baseline)
python bench.py --steps 5 --warmup 3 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; cut -c1-400 $O/bench_$TAG.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err; cut -c1-300 $O/bench_ref_$TAG.json
# dominant kernel: DRAM traffic + executed fp64 at the bench workload, then one --set full capture per kernel
bash tools/ncu_traffic.sh 4096 | tail -1
ncu --set full --clock-control none --import-source on -k regex:limber_spline_fast_kernel -c 1 -f -o $O/prof_spline_$TAG \
    python bench.py --ncosmo 64 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --from-parameters 0 --qag-sample 0 --no-y10 \
    > $O/ncu_spline_$TAG.log 2>&1; echo ncu-spline rc=$?
ncu --set full --clock-control none --import-source on -k regex:limber_qag_fast_kernel -c 1 -f -o $O/prof_qag_$TAG \
    python bench.py --ncosmo 64 --steps 1 --warmup 1 --method qag_quad --no-e2e --no-cpu-baseline --from-parameters 0 --qag-sample 0 --no-y10 \
    > $O/ncu_qag_$TAG.log 2>&1; echo ncu-qag rc=$?
# launch list of the bench command (reduced batch so that the serialised pass stays short)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --ncosmo 512 --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launches_$TAG.log 2>&1; echo ncu-launches rc=$?
python bench.py --steps 3 --warmup 3 --method qag_quad --no-e2e --no-cpu-baseline --from-parameters 0 --qag-sample 0 --no-y10 \
    > $O/bench_qag_4096_$TAG.json 2> $O/bench_qag_4096_$TAG.err; cut -c1-200 $O/bench_qag_4096_$TAG.json
python tools/bench_api_call.py > $O/api_call_$TAG.txt 2>&1; tail -4 $O/api_call_$TAG.txt
python tools/bench_y10.py > $O/y10_$TAG.txt 2>&1; tail -1 $O/y10_$TAG.txt
# multi-GPU (gpurun --gpus 8): bash tools/scale_run.sh "1 2 4 8"
