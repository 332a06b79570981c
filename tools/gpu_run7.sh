mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2g_pytest.txt
tail -5 gpurun_out/r2g_pytest.txt
# the two bench arms as the driver runs them (biphone headline)
timeout 600 python bench.py --impl reference > gpurun_out/r2g_bench_reference_arm.json 2> gpurun_out/r2g_bench_reference_arm.err
timeout 900 python bench.py > gpurun_out/r2g_bench_cfg4_biphone.json 2> gpurun_out/r2g_bench_cfg4_biphone.err
timeout 900 python bench.py --variant uniform --no-other-variants > gpurun_out/r2g_bench_cfg4_uniform.json 2> gpurun_out/r2g_bench_cfg4_uniform.err
# every BASELINE config
rm -f gpurun_out/r2g_all_configs.jsonl
for cfg in cfg1 cfg2 cfg3 cfg5; do
  timeout 600 python bench.py --config $cfg --steps 20 --warmup 5 --no-other-variants 2>gpurun_out/r2g_bench_$cfg.err | tail -1 >> gpurun_out/r2g_all_configs.jsonl
done
cat gpurun_out/r2g_bench_cfg4_biphone.json gpurun_out/r2g_bench_cfg4_uniform.json >> gpurun_out/r2g_all_configs.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/r2g_all_configs.jsonl'):
    try: d=json.loads(l)
    except Exception as e: print("bad line", l[:200]); continue
    cb=d.get("cpu_baseline") or {}; rc=d.get("reference_cuda") or {}
    print(d["config"]["workload"][:70], "| %.3f ms | %.2f Mframes/s | e2e %.2f M | frac %.3f | cpu %s | refcuda %s | objf_rel_err %s grad %s | ok %s" % (d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6, d["roofline"]["frac"], cb.get("value"), rc.get("value"), d.get("objf_rel_err"), d.get("grad_max_abs_err"), d["ok"]))
PY
# launch list of the bench command + full ncu of the dominant kernel (biphone)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/r2g_launch_list.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-traffic --no-other-variants > gpurun_out/r2g_launch_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:den_shared -s 2 -c 1 -f -o gpurun_out/r2g_ncu_biphone python bench.py --probe > gpurun_out/r2g_ncu_biphone.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:den_shared -s 2 -c 1 -f -o gpurun_out/r2g_ncu_uniform python bench.py --variant uniform --probe > gpurun_out/r2g_ncu_uniform.log 2>&1
# smoke() launch list (what the driver records)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_smoke_launches.csv python __graft_entry__.py smoke > gpurun_out/r2g_smoke.log 2>&1
# sanitizer on a small end-to-end case
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_case.py > gpurun_out/r2g_memcheck.txt 2>&1; tail -3 gpurun_out/r2g_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_case.py > gpurun_out/r2g_racecheck.txt 2>&1; tail -3 gpurun_out/r2g_racecheck.txt
ls -la gpurun_out/r2g*
