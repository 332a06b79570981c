mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2c_pytest.txt
tail -6 gpurun_out/r2c_pytest.txt
Q="--no-cpu-baseline --no-ref-cuda --no-traffic --steps 10 --warmup 3"
for v in uniform biphone skew; do
  timeout 200 python bench.py --variant $v $Q > gpurun_out/r2c_q_${v}.json 2> gpurun_out/r2c_q_${v}.err
done
for f in gpurun_out/r2c_q_*.json; do echo $f; python -c "
import json,sys
try:
  d=json.load(open('$f')); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['objf'], d['ok'])
except Exception as e: print('ERR', e)
"; done
for v in biphone uniform; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:den_shared -s 2 -c 1 -f -o gpurun_out/r2c_ncu_$v python bench.py --variant $v --probe > gpurun_out/r2c_ncu_$v.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
