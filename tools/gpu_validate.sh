# The round's validation + measurement pass on one B200 (gpurun -- 'bash tools/gpu_validate.sh'):
# GPU tests, the two bench arms as the driver runs them, every BASELINE config, launch lists, ncu of the dominant
# kernel, sanitizers.  Everything lands in gpurun_out/r2_*; the summaries under profiles/ are made from it.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_pytest.txt
tail -5 gpurun_out/r2_pytest.txt
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
timeout 900 python bench.py > gpurun_out/r2_bench_cfg4_biphone.json 2> gpurun_out/r2_bench_cfg4_biphone.err
timeout 900 python bench.py --variant uniform --no-other-variants > gpurun_out/r2_bench_cfg4_uniform.json 2> gpurun_out/r2_bench_cfg4_uniform.err
rm -f gpurun_out/r2_all_configs.jsonl
for cfg in cfg1 cfg2 cfg3 cfg5; do
  timeout 600 python bench.py --config $cfg --steps 20 --warmup 5 2>gpurun_out/r2_bench_$cfg.err | tail -1 >> gpurun_out/r2_all_configs.jsonl
done
cat gpurun_out/r2_bench_cfg4_biphone.json gpurun_out/r2_bench_cfg4_uniform.json >> gpurun_out/r2_all_configs.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/r2_all_configs.jsonl'):
    try: d=json.loads(l)
    except Exception as e: print("bad line", l[:200]); continue
    cb=d.get("cpu_baseline") or {}; rc=d.get("reference_cuda") or {}
    print(d["config"]["workload"][:70], "| %.3f ms | %.2f Mframes/s | e2e %.2f M | frac %.3f | cpu %s | refcuda %s | objf_rel_err %s grad %s | ok %s" % (d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6, d["roofline"]["frac"], cb.get("value"), rc.get("value"), d.get("objf_rel_err"), d.get("grad_max_abs_err"), d["ok"]))
PY
Q="--steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-traffic --no-other-variants"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4 -c 40 --csv --log-file gpurun_out/r2_launch_list.csv python bench.py $Q > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_smoke_launches.csv python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:den_shared -s 2 -c 1 -f -o gpurun_out/r2_ncu_biphone python bench.py --probe > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:den_shared -s 2 -c 1 -f -o gpurun_out/r2_ncu_uniform python bench.py --variant uniform --probe > /dev/null 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_case.py > gpurun_out/r2_memcheck.txt 2>&1; tail -1 gpurun_out/r2_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_case.py > gpurun_out/r2_racecheck.txt 2>&1; tail -1 gpurun_out/r2_racecheck.txt
