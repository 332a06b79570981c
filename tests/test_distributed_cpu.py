"""world_size-2 gloo test of the batch-shard host logic (pychain_b200.distributed):
strided sharding + the single all-reduce of {objf, frames, not ok}.  The local
compute is the CPU oracle here (the CUDA engine needs a GPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pychain_b200 import ChainGraph, ChainGraphBatch, synth
from pychain_b200 import distributed as pd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    from oracle import chain_oracle as co
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = ChainGraph(synth.random_den_fst(12, 40, 8, seed=3))
    x, lengths = synth.make_inputs(7, 20, 8, seed=4, variable=True)
    x = x.clamp(-30, 30)

    def compute(xs, ls):
        o, _, _, ok, _ = co.den_c(ChainGraphBatch(g, xs.size(0)), xs.exp(), ls, 1e-5)
        return torch.tensor(o), ok

    objf, frames, ok, idx = pd.sharded_objf(compute, x, lengths, rank, world)
    # every rank must hold the same global numbers
    out[rank] = (objf.item(), frames.item(), bool(ok), idx.tolist(), [int(v) for v in lengths[idx]])
    dist.destroy_process_group()


def test_sharded_objf_two_ranks_gloo():
    from oracle import chain_oracle as co
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    g = ChainGraph(synth.random_den_fst(12, 40, 8, seed=3))
    x, lengths = synth.make_inputs(7, 20, 8, seed=4, variable=True)
    o_all, _, _, _, _ = co.den_c(ChainGraphBatch(g, 7), x.clamp(-30, 30).exp(), lengths, 1e-5)
    assert abs(out[0][0] - out[1][0]) < 1e-6 and abs(out[0][0] - o_all) < 1e-4 * abs(o_all)
    assert out[0][1] == out[1][1] == float(lengths.sum())
    assert out[0][2] and out[1][2]
    assert out[0][3] == [0, 2, 4, 6] and out[1][3] == [1, 3, 5]
    for r in range(world):  # shards stay sorted descending (ChainFunction's requirement)
        assert out[r][4] == sorted(out[r][4], reverse=True)


def test_shard_batch_reorders_numerator_graphs():
    nb = synth.make_num_batch(5, 40, 16, seed=5, min_phones=3, max_phones=6)
    x, lengths = synth.make_inputs(5, 40, 16, seed=6, variable=True)
    probs = nb.forward_transition_probs.clone()
    xs, ls, ng, idx = pd.shard_batch(x, lengths, 1, 2, nb)
    assert idx.tolist() == [1, 3] and xs.shape[0] == 2 and ng.batch_size == 2
    assert torch.equal(ng.forward_transition_probs, probs[[1, 3]])
    assert nb.forward_transition_probs.shape[0] == 5  # the caller's batch is untouched
