# usage (gpurun --gpus N): bash tools/scale_run.sh "1 2 4 8" [extra bench flags]: one bench line per N into gpurun_out/scale_r2_n<N>.json
NS=$1; shift
for n in $NS; do
  if [ $n = 1 ]; then
    python bench.py --gpus 1 --steps 5 --warmup 3 "$@" > gpurun_out/scale_r2_n$n.json 2> gpurun_out/scale_r2_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n \
      bench.py --gpus $n --steps 5 --warmup 3 "$@" > gpurun_out/scale_r2_n$n.json 2> gpurun_out/scale_r2_n$n.err
  fi
  echo "N=$n rc=$?"; tail -2 gpurun_out/scale_r2_n$n.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/scale_r2_n$n.json"))
    e = d.get("e2e") or {}; y = d.get("y10") or {}
    print("N=%d value %.3e ms %.2f e2e %.3e (%.1f ms) y10 %.3e (%.3f ms, same %s)" % (d["n_gpus"], d["value"], d["ms_per_step"],
          e.get("value", 0), e.get("ms_per_step", 0), y.get("value", 0), y.get("ms_per_pass", 0), y.get("identical_to_unsplit")))
except Exception as ex:
    print("no line:", ex)
PY
done
