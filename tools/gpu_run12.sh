mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2k_pytest.txt
tail -3 gpurun_out/r2k_pytest.txt
Q="--no-cpu-baseline --no-ref-cuda --no-traffic --no-other-variants --steps 10 --warmup 3"
for c in cfg1 cfg2 cfg3 cfg5 cfg4; do
  timeout 200 python bench.py --config $c $Q > gpurun_out/r2k_q_${c}.json 2>/dev/null
  CHAIN_B200_NO_ALPHA_PREFETCH=1 timeout 200 python bench.py --config $c $Q > gpurun_out/r2k_q_${c}_noprefetch.json 2>/dev/null
done
timeout 200 python bench.py --config cfg5 --scaling strong $Q > gpurun_out/r2k_q_cfg5strong.json 2>/dev/null
for f in gpurun_out/r2k_q_*.json; do echo -n "$f  "; python -c "
import json,sys
try:
  d=json.load(open('$f')); print(round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d['objf'], d['ok'])
except Exception as e: print('ERR', e)
"; done
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_case.py > gpurun_out/r2k_racecheck.txt 2>&1; tail -2 gpurun_out/r2k_racecheck.txt
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_case.py > gpurun_out/r2k_memcheck.txt 2>&1; tail -1 gpurun_out/r2k_memcheck.txt
