# 'qag_quad' resident pass: grid-sharing kernel of each library variant, then the general kernel (CCL_B200_NO_QAG_FAST=1)
# usage: [NCOSMO=64] bash tools/ab_qag.sh [lib names...]
libs="$@"; [ -z "$libs" ] && libs="libccl_b200"
for v in $libs general; do
  if [ $v = general ]; then export CCL_B200_NO_QAG_FAST=1; lib=$PWD/ccl_b200/libccl_b200.so; else unset CCL_B200_NO_QAG_FAST; lib=$PWD/ccl_b200/$v.so; fi
  echo "== qag $v"
  CCL_B200_LIB=$lib python bench.py --method qag_quad --ncosmo ${NCOSMO:-64} --steps 2 --warmup 2 --no-e2e --from-parameters 0 --qag-sample 0 --no-y10 ${QAG_EXTRA:---no-cpu-baseline} 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('ms/step %.2f  frac %.4f  value %.3e fail %d nodes %.3e maxerr %s computed %s'%(d['ms_per_step'],d['roofline']['frac'],d['value'],d['run']['failed_cl_entries'],d['roofline']['integrand_nodes_per_launch'],(d.get('cpu_baseline') or {}).get('max_rel_err_gpu_vs_cpu_on_sample'), d['run'].get('qag_abscissae_computed')))
    else: print(ln.rstrip()[:300])
"
done
