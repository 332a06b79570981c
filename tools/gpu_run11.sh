mkdir -p gpurun_out
CHAIN_B200_EMBED_CODES=1 timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_den_shared.py tests/test_gpu_edges.py tests/test_gpu_fused_loss.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r2j_pytest_embed.txt
tail -3 gpurun_out/r2j_pytest_embed.txt
Q="--no-cpu-baseline --no-ref-cuda --no-traffic --no-other-variants --steps 10 --warmup 3"
for c in cfg1 cfg3 cfg5; do
  timeout 200 python bench.py --config $c $Q > gpurun_out/r2j_q_${c}_default.json 2>/dev/null
  CHAIN_B200_EMBED_CODES=1 timeout 200 python bench.py --config $c $Q > gpurun_out/r2j_q_${c}_embed_res.json 2>/dev/null
  CHAIN_B200_EMBED_CODES=1 CHAIN_B200_NO_RESIDENT_RECORDS=1 timeout 200 python bench.py --config $c $Q > gpurun_out/r2j_q_${c}_embed_nores.json 2>/dev/null
done
timeout 200 python bench.py --config cfg5 --scaling strong $Q > gpurun_out/r2j_q_cfg5strong_default.json 2>/dev/null
CHAIN_B200_EMBED_CODES=1 timeout 200 python bench.py --config cfg5 --scaling strong $Q > gpurun_out/r2j_q_cfg5strong_embed_res.json 2>/dev/null
for f in gpurun_out/r2j_q_*.json; do echo -n "$f  "; python -c "
import json,sys
try:
  d=json.load(open('$f')); print(round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d['objf'], d['ok'])
except Exception as e: print('ERR', e)
"; done
