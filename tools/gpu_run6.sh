mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_den_shared.py tests/test_gpu_edges.py tests/test_gpu_generic.py tests/test_gpu_fused_loss.py tests/test_gpu_dropin.py -m gpu -q 2>&1 | tail -25 > gpurun_out/r2f_pytest.txt
tail -6 gpurun_out/r2f_pytest.txt
Q="--no-cpu-baseline --no-ref-cuda --no-traffic --no-other-variants --steps 10 --warmup 3"
for lib in A_inl_3_5 B_inl_4_3a C_call_3_5 D_call_4_5 E_nw8_8_10 F_nw12_6_8 G_call_nofence; do
for w in "cfg4 uniform" "cfg4 biphone" "cfg5 uniform"; do
  set -- $w
  CHAIN_B200_LIB=$PWD/pychain_b200/lib/libchain_b200_$lib.so timeout 200 python bench.py --config $1 --variant $2 $Q > gpurun_out/r2f_q_${1}_${2}_${lib}.json 2> gpurun_out/r2f_q_${1}_${2}_${lib}.err
done; done
timeout 200 python bench.py --config cfg3 $Q > gpurun_out/r2f_q_cfg3_default.json 2> gpurun_out/r2f_q_cfg3_default.err
CHAIN_B200_NO_RS=1 timeout 200 python bench.py --config cfg3 $Q > gpurun_out/r2f_q_cfg3_nors.json 2> gpurun_out/r2f_q_cfg3_nors.err
timeout 200 python bench.py --config cfg2 $Q > gpurun_out/r2f_q_cfg2_default.json 2> gpurun_out/r2f_q_cfg2_default.err
for f in gpurun_out/r2f_q_*.json; do echo -n "$f  "; python -c "
import json,sys
try:
  d=json.load(open('$f')); print(round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d['objf'], d['ok'])
except Exception as e: print('ERR', e)
"; done
