# every BASELINE config through bench.py (1 GPU) -> gpurun_out/r1_all_configs.jsonl
rm -f gpurun_out/r1_all_configs.jsonl
for cfg in cfg1 cfg2 cfg3 cfg4 cfg5; do
  python bench.py --config $cfg --steps 20 --warmup 5 2>gpurun_out/bench_$cfg.err | tail -1 >> gpurun_out/r1_all_configs.jsonl
done
python bench.py --config cfg4 --variant skew --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/r1_all_configs.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/r1_all_configs.jsonl'):
    try: d=json.loads(l)
    except Exception as e: print("bad line", l[:200]); continue
    cb=d.get("cpu_baseline") or {}
    print(d["config"]["workload"][:60], "| %.3f ms | %.2f Mframes/s | e2e %.2f M | frac %.3f | cpu %s (%s cores) | ok %s" % (d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6, d["roofline"]["frac"], cb.get("value"), cb.get("cores"), d["ok"]))
PY
