mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2h_pytest.txt
tail -5 gpurun_out/r2h_pytest.txt
Q="--no-cpu-baseline --no-ref-cuda --no-traffic --no-other-variants --steps 10 --warmup 3"
for w in "cfg4 uniform" "cfg4 biphone" "cfg4 skew" "cfg5 uniform" "cfg2 uniform" "cfg3 uniform" "cfg1 uniform"; do
  set -- $w
  timeout 200 python bench.py --config $1 --variant $2 $Q > gpurun_out/r2h_q_${1}_${2}.json 2> gpurun_out/r2h_q_${1}_${2}.err
done
CHAIN_B200_EMBED_CODES=1 timeout 200 python bench.py --config cfg4 --variant uniform $Q > gpurun_out/r2h_q_cfg4_uniform_embed.json 2>/dev/null
CHAIN_B200_EMBED_CODES=1 timeout 200 python bench.py --config cfg5 --variant uniform $Q > gpurun_out/r2h_q_cfg5_uniform_embed.json 2>/dev/null
CHAIN_B200_EMBED_CODES=0 timeout 200 python bench.py --config cfg4 --variant biphone $Q > gpurun_out/r2h_q_cfg4_biphone_compact.json 2>/dev/null
CHAIN_B200_EMBED_CODES=0 timeout 200 python bench.py --config cfg4 --variant skew $Q > gpurun_out/r2h_q_cfg4_skew_compact.json 2>/dev/null
for f in gpurun_out/r2h_q_*.json; do echo -n "$f  "; python -c "
import json,sys
try:
  d=json.load(open('$f')); print(round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d['objf'], d['ok'])
except Exception as e: print('ERR', e)
"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4 -c 40 --csv --log-file gpurun_out/r2h_launch_list.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-traffic --no-other-variants > gpurun_out/r2h_launch_list.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_case.py > gpurun_out/r2h_racecheck.txt 2>&1; tail -3 gpurun_out/r2h_racecheck.txt
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_case.py > gpurun_out/r2h_memcheck.txt 2>&1; tail -2 gpurun_out/r2h_memcheck.txt
