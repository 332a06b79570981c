"""Two ranks over NCCL: the batch-sharded objective equals the single-GPU run on the same global batch
(SURVEY.md 4 item 4 / 8(e)).  The only cross-utterance operations of the reference are the objf sum
(chain-computation.cc:229) and the frame count (loss.py:104): here ONE all-reduce of the three floats the
finalize kernel leaves in `Result.stats`.  Needs >= 2 GPUs (skipped on the 1-GPU test box; run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pychain_b200 import ChainGraphBatch, synth, engine
from pychain_b200 import distributed as pd

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case():
    spec = synth.workload_spec("cfg5", B=64, T=60)
    den, _, x, _ = synth.build_workload(spec)
    lengths = synth.make_lengths(spec["B"], spec["T"], 7, True)
    return spec, den, x, lengths


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    spec, den, x, lengths = _case()
    xs, ls, _, idx = pd.shard_batch(x, lengths, rank, world)
    gb = ChainGraphBatch(den, xs.size(0))
    gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_PROB)
    r = engine.forward_backward_prob(gt, xs.to(dev), ls, spec["S"], 1e-5, x_is_exp=False, clamp=True)
    objf, frames, ok = pd.reduce_result(r)          # ONE all-reduce, no other launch
    torch.save(dict(objf=objf.item(), frames=frames.item(), ok=bool(ok), idx=idx, grad=r.grad.cpu(), logp=r.logp.cpu()),
               os.path.join(out_dir, "rank%d.pt" % rank))
    dist.destroy_process_group()


def test_two_rank_nccl_shards_equal_one_gpu(cuda, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    spec, den, x, lengths = _case()
    gb = ChainGraphBatch(den, x.size(0))
    gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_PROB)
    full = engine.forward_backward_prob(gt, x.to(cuda), lengths, spec["S"], 1e-5, x_is_exp=False, clamp=True)
    parts = [torch.load(os.path.join(str(tmp_path), "rank%d.pt" % r)) for r in range(world)]
    for p in parts:   # every rank holds the same global numbers
        assert abs(p["objf"] - full.objf.item()) < 1e-6 * abs(full.objf.item())
        assert p["frames"] == float(lengths.sum()) and p["ok"]
        # each utterance's arithmetic does not depend on which other utterances share its group
        assert torch.equal(p["logp"], full.logp.cpu()[p["idx"]])
        assert torch.equal(p["grad"], full.grad.cpu()[p["idx"]])
    assert sorted(torch.cat([p["idx"] for p in parts]).tolist()) == list(range(x.size(0)))
