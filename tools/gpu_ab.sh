mkdir -p gpurun_out
Q="--no-cpu-baseline --no-ref-cuda --no-traffic --no-other-variants --steps 10 --warmup 3"
for lib in default n1_4 n1_3 n2_2 n2_2_n1_3 n1_6; do
for w in "cfg4 uniform" "cfg4 biphone"; do
  set -- $w
  if [ $lib = default ]; then L=""; else L="$PWD/pychain_b200/lib/libchain_b200_$lib.so"; fi
  CHAIN_B200_LIB=$L timeout 200 python bench.py --config $1 --variant $2 $Q > gpurun_out/ab_${1}_${2}_${lib}.json 2>/dev/null
done; done
for f in gpurun_out/ab_*.json; do echo -n "$f  "; python -c "
import json,sys
try:
  d=json.load(open('$f')); print(round(d['ms_per_step'],3), d['objf'], d['ok'])
except Exception as e: print('ERR', e)
"; done
