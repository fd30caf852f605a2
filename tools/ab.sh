# usage: bash tools/ab.sh [lib names...]; default: the in-tree build (variants: tools/build_variants.sh).
libs="$@"; [ -z "$libs" ] && libs="libccl_b200"
for v in $libs; do
  echo "== $v"
  lib=$PWD/ccl_b200/$v.so; extra=""
  CCL_B200_LIB=$lib python bench.py --ncosmo ${NCOSMO:-512} --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --from-parameters 0 --qag-sample 0 --no-y10 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('ms/step %.2f  frac %.4f  value %.3e fail %d maxerr %s'%(d['ms_per_step'],d['roofline']['frac'],d['value'],d.get('run',d['config']).get('failed_cl_entries'),(d.get('cpu_baseline') or {}).get('max_rel_err_gpu_vs_cpu_on_sample')))
    else: print(ln.rstrip()[:200])
"
done
