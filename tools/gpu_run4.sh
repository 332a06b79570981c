mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_den_shared.py tests/test_gpu_edges.py tests/test_gpu_generic.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r2d_pytest.txt
tail -4 gpurun_out/r2d_pytest.txt
Q="--no-cpu-baseline --no-ref-cuda --no-traffic --steps 10 --warmup 3"
for lib in default nf1_4 nf2_4; do
for v in uniform biphone; do
  if [ $lib = default ]; then L=""; else L="$PWD/pychain_b200/lib/libchain_b200_$lib.so"; fi
  CHAIN_B200_LIB=$L timeout 200 python bench.py --variant $v $Q > gpurun_out/r2d_q_${v}_${lib}.json 2> gpurun_out/r2d_q_${v}_${lib}.err
done; done
for f in gpurun_out/r2d_q_*.json; do echo $f; python -c "
import json,sys
try:
  d=json.load(open('$f')); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['objf'], d['ok'])
except Exception as e: print('ERR', e)
"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:den_shared -s 2 -c 1 -f -o gpurun_out/r2d_ncu_biphone python bench.py --variant biphone --probe > gpurun_out/r2d_ncu_biphone.log 2>&1
ls -la gpurun_out/r2d*.ncu-rep
