"""Batch-sharded LF-MMI objective across GPUs (one process per GPU).

Utterances are independent (chain-computation.h:33-35 of the reference), so the
minibatch shards with NO data-path collective: every rank runs the kernels on
its own utterances and keeps its slice of the gradient local.  The only exchange
is ONE all-reduce(sum) of three floats {objf, frames, not_ok} -- the reference
has no distributed code at all; this is the 8xB200 layout
BASELINE.json asks for.

`shard_indices` gives rank r the utterances r, r+W, r+2W, ... of the
length-sorted batch, so every rank sees the same length mix and its shard stays
sorted descending (what `ChainFunction` requires).
"""
import torch
import torch.distributed as dist

__all__ = ["shard_indices", "shard_batch", "reduce_stats", "reduce_result", "sharded_objf"]


def shard_indices(batch_size, rank, world_size):
    return torch.arange(rank, batch_size, world_size)


def shard_batch(x, lengths, rank, world_size, num_graphs=None):
    """Slice a (length-sorted) global batch for `rank`.  `num_graphs` (a
    per-utterance ChainGraphBatch) is re-ordered to the same selection."""
    idx = shard_indices(x.size(0), rank, world_size)
    xs = x.index_select(0, idx.to(x.device))
    ls = lengths.index_select(0, idx.to(lengths.device))
    if num_graphs is not None:
        import copy
        num_graphs = copy.copy(num_graphs)
        num_graphs.reorder(idx)
        num_graphs.batch_size = int(idx.numel())
    return xs, ls, num_graphs, idx


def reduce_stats(objf, frames, ok, group=None):
    """The single collective: sum {objf, frames, not ok} over ranks.  Returns
    (objf_sum tensor, frames_sum tensor, all_ok tensor) on objf's device."""
    buf = torch.stack([objf.reshape(()).float(),
                       torch.as_tensor(float(frames), device=objf.device),
                       (1 - ok.reshape(-1)[0].float()) if torch.is_tensor(ok) else torch.as_tensor(0.0 if ok else 1.0, device=objf.device)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf[0], buf[1], buf[2] == 0


def reduce_result(r, group=None):
    """The same single collective for an engine `Result`: its `stats` view already holds
    {objf, valid frames, not ok} (written by the finalize kernel), so the step costs ONE all-reduce and no
    other launch or host-to-device write.  Returns (objf_sum, frames_sum, all_ok) tensors."""
    buf = r.stats
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf[0], buf[1], buf[2] == 0


def sharded_objf(compute, x, lengths, rank, world_size, group=None):
    """objf / frames of the global batch from per-rank shards.  `compute(xs, ls)`
    -> (objf tensor, ok) runs the local forward-backward (the CUDA engine in
    production; the tests pass the CPU oracle to exercise this host logic under
    gloo)."""
    xs, ls, _, idx = shard_batch(x, lengths, rank, world_size)
    objf, ok = compute(xs, ls)
    return reduce_stats(objf, int(ls.sum()), ok, group=group) + (idx,)
