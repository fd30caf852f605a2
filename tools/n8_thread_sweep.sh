for cfg in "2 2" "1 2" "2 1" "3 3" "1 1"; do set -- $cfg
  CCL_B200_NSTAGE=$1 CCL_B200_NCOMMIT=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2961$1 bench.py --gpus 8 --steps 4 --warmup 3 --no-cpu-baseline --from-parameters 0 --qag-sample 0 --no-y10 2>/dev/null | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('NSTAGE $1 NCOMMIT $2: value ms %.2f  e2e ms %.2f  e2e %.3e' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['value']))
"
done
