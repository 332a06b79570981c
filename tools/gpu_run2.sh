mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2b_pytest.txt
tail -5 gpurun_out/r2b_pytest.txt
Q="--no-cpu-baseline --no-ref-cuda --no-traffic --steps 10 --warmup 3"
for v in uniform biphone skew; do
  timeout 200 python bench.py --variant $v $Q > gpurun_out/r2b_q_${v}_nf3.json 2> gpurun_out/r2b_q_${v}_nf3.err
  CHAIN_B200_LIB=$PWD/pychain_b200/lib/libchain_b200_nf4.so timeout 200 python bench.py --variant $v $Q > gpurun_out/r2b_q_${v}_nf4.json 2> gpurun_out/r2b_q_${v}_nf4.err
done
CHAIN_B200_MAX_CLASSES=0 timeout 200 python bench.py --variant biphone $Q > gpurun_out/r2b_q_biphone_nocls.json 2> gpurun_out/r2b_q_biphone_nocls.err
for f in gpurun_out/r2b_q_*.json; do echo $f; python -c "
import json,sys
try:
  d=json.load(open('$f')); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['objf'], d['ok'])
except Exception as e: print('ERR', e)
"; done
timeout 500 python bench.py > gpurun_out/r2b_bench_uniform.json 2> gpurun_out/r2b_bench_uniform.err
timeout 500 python bench.py --variant biphone > gpurun_out/r2b_bench_biphone.json 2> gpurun_out/r2b_bench_biphone.err
cut -c1-1500 gpurun_out/r2b_bench_biphone.json
