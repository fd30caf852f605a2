# usage: bash tools/ab.sh [lib names...]; default: the in-tree build
python -m pytest tests -m gpu -q 2>&1 | tail -3
libs="$@"; [ -z "$libs" ] && libs="libccl_b200"
for v in $libs; do
  echo "== $v"
  CCL_B200_LIB=$PWD/ccl_b200/$v.so python bench.py --ncosmo 512 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('ms/step %.2f  frac %.4f  value %.3e fail %d'%(d['ms_per_step'],d['roofline']['frac'],d['value'],d['config']['failed_cl_entries']))
    else: print(ln.rstrip()[:200])
"
done
