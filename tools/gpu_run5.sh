mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2e_pytest.txt
tail -8 gpurun_out/r2e_pytest.txt
Q="--no-cpu-baseline --no-ref-cuda --no-traffic --steps 10 --warmup 3"
for v in uniform biphone skew; do
  timeout 200 python bench.py --variant $v $Q > gpurun_out/r2e_q_cfg4_${v}.json 2> gpurun_out/r2e_q_cfg4_${v}.err
done
for c in cfg2 cfg3 cfg5; do
  timeout 200 python bench.py --config $c $Q > gpurun_out/r2e_q_${c}.json 2> gpurun_out/r2e_q_${c}.err
done
for f in gpurun_out/r2e_q_*.json; do echo $f; python -c "
import json,sys
try:
  d=json.load(open('$f')); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['objf'], d['ok'], d['gpu_launches'])
except Exception as e: print('ERR', e)
"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:den_shared -s 2 -c 1 -f -o gpurun_out/r2e_ncu_uniform python bench.py --variant uniform --probe > gpurun_out/r2e_ncu_uniform.log 2>&1
ls -la gpurun_out/r2e*.ncu-rep
