"""The reference's CPU path on a BASELINE-sized batch, sharded over host processes.

TEST INFRASTRUCTURE ONLY (see chain_oracle.c header): used by the at-size parity tests
(tests/test_gpu_parity_at_size.py), by bench.py's `cpu_baseline` / `--impl reference` legs and by
nothing in the product package.

The reference extension's arc loops are single-threaded and it holds the GIL for the whole call
(chain-computation.cc:150-174, 289-309; pychain.cc has no gil_scoped_release), so the only way it
can use the box's cores is one process per shard.  Utterances are independent
(chain-computation.h:33-35): a contiguous shard of the length-sorted batch is itself a valid call
(lengths sorted descending; the shard's input is trimmed to its own longest utterance because the
reference requires max(lengths) == T, SURVEY.md Appx C-1).

Every worker REGENERATES its inputs from the workload description (`spec`): the graph from its seed
and the utterances from the blocked generator `pychain_b200.synth.input_block`, i.e. exactly the
tensors the GPU arm holds, without shipping a 410 MB tensor to 64 processes.

spec keys: mode ('den' | 'logden' | 'loss'), S, K, D, T, B, variant, graph_seed, x_seed, block,
variable, coeff, num_seed (loss only), x_scale (optional multiplier applied before the clamp).
"""
import os
import time

import numpy as np
import torch


def build_inputs(spec, start, count):
    """(den graph, numerator batch or None, x[count,T,D], lengths[count]) for utterances start..start+count."""
    from pychain_b200 import synth
    return synth.build_workload(spec, start, count)


def _worker(task):
    spec, start, count, reps, grad_for = task
    torch.set_num_threads(1)
    from pychain_b200 import ChainGraphBatch
    from oracle import chain_oracle as co
    den, num, x, lengths = build_inputs(spec, start, count)
    Tm = int(lengths.max())
    xc = x[:, :Tm].clamp(-30, 30).contiguous()   # loss.py:30; the shard's own T (Appx C-1)
    use_ref = co.ref_available()
    mode = spec["mode"]
    gb = ChainGraphBatch(den, count).materialize()   # what graph.py:101-120 hands the extension
    ex = xc.exp() if mode != "logden" else None      # loss.py:43
    walls, out = [], None
    for _ in range(reps):
        t0 = time.perf_counter()
        if mode == "logden":
            if use_ref:
                o, lg, ok = co.num_ref(gb, xc, lengths)
            else:
                o, _, lg, ok, _ = co.num_c(gb, xc, lengths)
            res = dict(objf=float(o), grad=lg, grad_is_log=True, ok=bool(ok))
        else:
            if use_ref:
                o, g, ok = co.den_ref(gb, ex, lengths, spec["coeff"])
            else:
                o, _, g, ok, _ = co.den_c(gb, ex, lengths, spec["coeff"])
            res = dict(objf=float(o), grad=g, grad_is_log=False, ok=bool(ok))
            if mode == "loss":
                if use_ref:
                    o2, lg2, ok2 = co.num_ref(num, xc, lengths)
                else:
                    o2, _, lg2, ok2, _ = co.num_c(num, xc, lengths)
                res.update(objf_num=float(o2), grad_num=lg2, ok=res["ok"] and bool(ok2))
        walls.append(time.perf_counter() - t0)
        out = res
    grads = {}
    for b in grad_for:
        if start <= b < start + count:
            g = out["grad"][b - start]
            g = g.exp() if out["grad_is_log"] else g
            if mode == "loss":   # d(num - den)/dx is what ChainLoss differentiates: keep both parts
                g = torch.stack([g, out["grad_num"][b - start].exp()])
            full = torch.zeros(g.shape[:-2] + (spec["T"], spec["D"]))
            full[..., :Tm, :] = g
            grads[b] = full.numpy()
    return dict(walls=walls, objf=out["objf"], objf_num=out.get("objf_num"), ok=out["ok"], grads=grads,
                kind="reference" if use_ref else "port", start=start, count=count,
                frames=int(lengths.sum()))


def run(spec, sample=None, reps=1, grad_for=(), max_workers=64, drop_first=False):
    """Run the reference CPU path over utterances [0, sample) of `spec`, `reps` passes each.
    Returns dict(walls=[seconds per pass = slowest worker], workers, kind, objf, objf_num, ok, grads{b: ndarray},
    frames).  `drop_first` runs one extra (cold) pass first and discards it."""
    import multiprocessing as mp
    B = spec["B"]
    sample = min(sample or B, B)
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, sample, max_workers))
    per = [sample // workers + (1 if i < sample % workers else 0) for i in range(workers)]
    starts = [sum(per[:i]) for i in range(workers)]
    n = reps + (1 if drop_first else 0)
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers) as pool:
        res = pool.map(_worker, [(spec, s, c, n, tuple(grad_for)) for s, c in zip(starts, per) if c > 0])
    first = 1 if drop_first else 0
    walls = [max(r["walls"][i] for r in res) for i in range(first, n)]
    grads = {}
    for r in res:
        grads.update(r["grads"])
    return dict(walls=walls, workers=len(res), kind=res[0]["kind"], objf=sum(r["objf"] for r in res),
                objf_num=(sum(r["objf_num"] for r in res) if res[0]["objf_num"] is not None else None),
                ok=all(r["ok"] for r in res), grads=grads, frames=sum(r["frames"] for r in res), sample=sample)


def spec_for(cfg, variant="uniform", **over):
    """The BASELINE configuration `cfg` as bench.py runs it (same seeds => same tensors)."""
    from pychain_b200 import synth
    return synth.workload_spec(cfg, variant, **over)
