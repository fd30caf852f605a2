#!/bin/bash
# DRAM traffic of the dominant kernel at the bench workload (one ncu pass, per launch) -> profiles/traffic.json
# usage (on the GPU box): bash tools/ncu_traffic.sh [ncosmo]
NC=${1:-4096}
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum --clock-control none \
    -k regex:limber_spline_fast_kernel -c 1 --csv --log-file gpurun_out/traffic.csv \
    python bench.py --ncosmo $NC --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --from-parameters 0 --qag-sample 0 --no-y10 > gpurun_out/traffic.log 2>&1
python - <<PY
import csv, json
rows = [r for r in csv.reader(open("gpurun_out/traffic.csv")) if len(r) > 5]
h = rows[0]
vals = {}
for r in rows[1:]:
    d = dict(zip(h, r))
    v = float(d["Metric Value"].replace(",", ""))
    u = d["Metric Unit"]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    vals[d["Metric Name"]] = v * mult
out = {"kernel": "limber_spline_fast_kernel", "ncosmo": $NC, "method": "spline",
       "dram_bytes_per_launch": vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"],
       "dram_bytes_read": vals["dram__bytes_read.sum"], "dram_bytes_write": vals["dram__bytes_write.sum"],
       "executed_fp64_flops_per_launch": vals.get("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", 0)
           + vals.get("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", 0)
           + 2 * vals.get("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", 0),
       "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, one launch, bench.py --ncosmo $NC"}
json.dump(out, open("gpurun_out/traffic.json", "w"), indent=1)
print(out)
PY
