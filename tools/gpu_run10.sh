mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_generic.py tests/test_gpu_den_shared.py tests/test_gpu_parity_at_size.py -m gpu -q -k "log or cfg2 or golden or cfg3" 2>&1 | tail -8 > gpurun_out/r2i_pytest.txt
tail -4 gpurun_out/r2i_pytest.txt
Q="--no-cpu-baseline --no-ref-cuda --no-traffic --no-other-variants --steps 10 --warmup 3"
for c in cfg2 cfg3; do
  timeout 200 python bench.py --config $c $Q > gpurun_out/r2i_q_$c.json 2> gpurun_out/r2i_q_$c.err
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 24 --csv --log-file gpurun_out/r2i_launches_$c.csv python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-traffic --no-other-variants > /dev/null 2>&1
done
for f in gpurun_out/r2i_q_*.json; do echo -n "$f  "; python -c "
import json,sys
try:
  d=json.load(open('$f')); print(round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d['objf'], d['ok'])
except Exception as e: print('ERR', e)
"; done
grep -o '"[a-zA-Z_:0-9<>, ]*kernel[^"]*"[^\n]*' gpurun_out/r2i_launches_cfg3.csv | head -0
python - <<'PY'
import csv
for c in ("cfg2","cfg3"):
    rows=[r for r in csv.reader(open(f"gpurun_out/r2i_launches_{c}.csv")) if len(r)>10 and r[0].isdigit()]
    print(c)
    for r in rows[:14]: print("  ", r[4][:70], r[-1], "ns grid", r[8], "block", r[7])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_generic -s 3 -c 1 -f -o gpurun_out/r2i_ncu_num python bench.py --config cfg3 --probe > gpurun_out/r2i_ncu_num.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:den_shared_log -s 3 -c 1 -f -o gpurun_out/r2i_ncu_logden python bench.py --config cfg2 --probe > gpurun_out/r2i_ncu_logden.log 2>&1
ls -la gpurun_out/r2i*.ncu-rep
