# multi-GPU: 2-rank NCCL equality test + strong scaling of cfg5 (global batch 4096 fixed) at 1/2/4/8 GPUs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r2_multi_pytest.txt
cat gpurun_out/r2_multi_pytest.txt
Q="--no-cpu-baseline --no-ref-cuda --no-traffic --no-other-variants --steps 20 --warmup 5"
timeout 300 python bench.py --gpus 1 --config cfg5 --scaling strong $Q > gpurun_out/r2_strong_cfg5_n1.json 2> gpurun_out/r2_strong_cfg5_n1.err
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --config cfg5 --scaling strong $Q > gpurun_out/r2_strong_cfg5_n$n.json 2> gpurun_out/r2_strong_cfg5_n$n.err
done
for n in 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n $Q > gpurun_out/r2_weak_cfg4_n$n.json 2> gpurun_out/r2_weak_cfg4_n$n.err
done
for f in gpurun_out/r2_strong_cfg5_n*.json gpurun_out/r2_weak_cfg4_n*.json; do echo -n "$f  "; python -c "
import json
try:
  d=json.load(open('$f')); print(d['n_gpus'], round(d['ms_per_step'],3), round(d['value']/1e6,2), 'Mframes/s', d['config']['global_batch'], d['objf'])
except Exception as e: print('ERR', e)
"; done
