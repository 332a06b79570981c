"""Seeded synthetic workloads (OpenFst-free) for tests and bench.py.

These are the generators SURVEY.md §8(d) specifies; the reference repository
ships no data, tests or benchmark inputs, so every workload here is synthetic.

* `random_den_fst`  -- denominator-like graph: S states, K arcs; a backbone
  s -> (s+1)%S guarantees every state has >=1 in- and out-arc and that the
  graph is strongly connected; the other K-S arcs are uniform (variant
  "skew": destination ~ Zipf, to model biphone in-degree skew); pdf uniform in
  [0,D); arc / final log-weights make every state row-stochastic.  Variant
  "biphone" has the structure of a real chain denominator instead (states =
  phone pairs, the pdf of an arc is a function of its destination).
* `linear_num_fst`  -- per-utterance numerator: left-to-right HMM, each phone
  = 2 states, self-loops + forward arcs, final only on the last state.
* `make_inputs`     -- x = 2*randn(B,T,D) and descending lengths with max == T
  (the `pack_padded_sequence` contract of `loss.py:37-40`).
* `CONFIGS`         -- the five BASELINE.json configurations.
"""
from __future__ import annotations

import numpy as np
import torch

from .graph import Fst, ChainGraph, ChainGraphBatch

__all__ = ["random_den_fst", "linear_num_fst", "make_inputs", "make_lengths", "input_block", "make_num_batch", "CONFIGS",
           "den_graph_for", "workload_spec", "build_workload"]


def random_den_fst(S: int, K: int, D: int, seed: int = 0, variant: str = "uniform") -> Fst:
    assert K >= S, "need at least one backbone arc per state"
    rng = np.random.default_rng(seed)
    src = np.empty(K, dtype=np.int64)
    dst = np.empty(K, dtype=np.int64)
    src[:S] = np.arange(S)
    dst[:S] = (np.arange(S) + 1) % S
    src[S:] = rng.integers(0, S, size=K - S)
    if variant == "uniform":
        dst[S:] = rng.integers(0, S, size=K - S)
    elif variant == "skew":
        z = rng.zipf(1.2, size=K - S)
        dst[S:] = (z - 1) % S
        perm = rng.permutation(S)  # so that the popular states are not 0,1,2,...
        dst[S:] = perm[dst[S:]]
    elif variant == "biphone":
        return _biphone_den_fst(S, K, D, rng)
    else:
        raise ValueError(variant)
    pdf = rng.integers(0, D, size=K)
    out_deg = np.bincount(src, minlength=S)
    logw = -np.log(out_deg[src] + 1.0)
    fin = -np.log(out_deg + 1.0)
    return Fst(S, 0, src, dst, pdf, logw.astype(np.float32), fin.astype(np.float32))


def _biphone_den_fst(S, K, D, rng):
    """A denominator graph with the structure of a biphone chain model (1-state HMMs, bigram-like
    phone LM): state = (previous phone a, current phone b); it has a self loop (self-loop pdf of b
    after a) and arcs to states (b, c) for a sparse set of successors c, all arcs ENTERING (b, c)
    from another state carrying the same forward pdf of c after b.  Unlike the uniform variant the
    pdf of an arc is a function of its destination, as in real chain graphs."""
    P = max(2, int(np.floor(np.sqrt(S))))
    n_states = P * P
    fan = max(1, (K - n_states) // n_states)          # successors per state besides the self loop
    succ = [rng.choice(P, size=min(fan, P), replace=False) for _ in range(P)]  # phone b -> next phones c
    half = max(1, D // 2)
    src, dst, pdf = [], [], []
    for a in range(P):
        for b in range(P):
            st = a * P + b
            src.append(st); dst.append(st); pdf.append(half + (a * 7919 + b * 104729) % (D - half))  # self loop
            for c in succ[b]:
                src.append(st); dst.append(b * P + int(c)); pdf.append((b * 31337 + int(c) * 613) % half)
    for st in range(n_states, S):  # padding states up to S: unreachable, a self loop keeps them well formed
        src.append(st); dst.append(st); pdf.append(st % D)
    src, dst, pdf = np.asarray(src), np.asarray(dst), np.asarray(pdf)
    out_deg = np.bincount(src, minlength=S)
    logw = -np.log(out_deg[src] + 1.0)
    fin = -np.log(out_deg + 1.0)
    return Fst(S, 0, src, dst, pdf, logw.astype(np.float32), fin.astype(np.float32))


def linear_num_fst(n_phones: int, D: int, seed: int = 0, states_per_phone: int = 2) -> Fst:
    """Left-to-right numerator graph: state i has a self-loop (pdf a_i) and a
    forward arc to i+1 (pdf b_i); weights -log 2; final on the last state."""
    rng = np.random.default_rng(seed)
    S = n_phones * states_per_phone + 1
    src, dst, pdf, lw = [], [], [], []
    for i in range(S - 1):
        a, b = rng.integers(0, D, size=2)
        src += [i, i]
        dst += [i, i + 1]
        pdf += [int(a), int(b)]
        lw += [-np.log(2.0), -np.log(2.0)]
    fin = np.full(S, -np.inf, dtype=np.float32)
    fin[S - 1] = 0.0
    # a self loop on the last state so that any T >= S-1 has a path
    src.append(S - 1), dst.append(S - 1), pdf.append(int(rng.integers(0, D))), lw.append(0.0)
    return Fst(S, 0, src, dst, pdf, np.asarray(lw, dtype=np.float32), fin)


def input_block(T: int, D: int, seed: int, index: int, block: int) -> torch.Tensor:
    """Block `index` (utterances index*block .. +block) of a blocked input tensor: every block has its own
    generator, so a host process that only needs some utterances (the sharded CPU reference of
    bench.py / the at-size parity tests) can regenerate exactly those without the whole [B,T,D] tensor."""
    g = torch.Generator().manual_seed(seed * 1000003 + 7 * index + 1)
    return 2.0 * torch.randn(block, T, D, generator=g, dtype=torch.float32)


def make_lengths(B: int, T: int, seed: int = 0, variable: bool = False) -> torch.Tensor:
    """lengths [B] int64, sorted descending, max == T (the `pack_padded_sequence` contract, loss.py:37-40)."""
    if not variable:
        return torch.full((B,), T, dtype=torch.long)
    g = torch.Generator().manual_seed(seed + 77)
    lengths = torch.randint(max(1, T // 2), T + 1, (B,), generator=g, dtype=torch.long)
    lengths[0] = T
    return torch.sort(lengths, descending=True)[0]


def make_inputs(B: int, T: int, D: int, seed: int = 0, variable: bool = False, device="cpu", block: int = 0):
    """x [B,T,D] fp32 (unnormalised, like chain nnet outputs) and lengths [B]
    int64, sorted descending, max == T.  With `block` > 0 the tensor is generated in independent
    blocks of `block` utterances (see `input_block`)."""
    if block:
        nb = (B + block - 1) // block
        x = torch.cat([input_block(T, D, seed, i, block) for i in range(nb)], 0)[:B].contiguous()
        return x.to(device), make_lengths(B, T, seed, variable)
    g = torch.Generator().manual_seed(seed)
    x = 2.0 * torch.randn(B, T, D, generator=g, dtype=torch.float32)
    if variable:
        lengths = torch.randint(max(1, T // 2), T + 1, (B,), generator=g, dtype=torch.long)
        lengths[0] = T
        lengths, _ = torch.sort(lengths, descending=True)
    else:
        lengths = torch.full((B,), T, dtype=torch.long)
    return x.to(device), lengths


def make_num_batch(B: int, T: int, D: int, seed: int = 0, min_phones: int = 10, max_phones: int = 40):
    """Per-utterance numerator graphs (log domain), batched the reference way
    (`graph.py:122-175`).  Number of phones is capped so a path fits in T/2."""
    rng = np.random.default_rng(seed)
    graphs = []
    cap = max(1, (T // 2 - 1) // 2)
    for b in range(B):
        n = int(rng.integers(min(min_phones, cap), min(max_phones, cap) + 1))
        graphs.append(ChainGraph(linear_num_fst(n, D, seed=seed * 100003 + b), log_domain=True))
    ms = max(g.num_states for g in graphs)
    mt = max(g.num_transitions for g in graphs)
    return ChainGraphBatch(graphs, max_num_transitions=mt, max_num_states=ms)


# BASELINE.json "configs", made concrete
CONFIGS = {
    "cfg1": dict(S=10, K=20, B=2, T=50, D=16, variable=False, what="toy den (plumbing)"),
    "cfg2": dict(S=1000, K=5000, B=64, T=150, D=256, variable=False, what="den ~1k states/5k arcs"),
    "cfg3": dict(S=1000, K=5000, B=128, T=150, D=512, variable=True, what="full ChainLoss num+den, variable T"),
    "cfg4": dict(S=5000, K=40000, B=512, T=200, D=1000, variable=False, what="biphone den ~5k states/40k arcs (headline)"),
    "cfg5": dict(S=1000, K=5000, B=512, T=150, D=512, variable=False, B_global=4096,
                 what="per-GPU shard of the B=4096 batch"),
}


def workload_spec(cfg: str, variant: str = "uniform", rank: int = 0, **over) -> dict:
    """The BASELINE configuration `cfg` as a seeded description every arm of bench.py and the at-size parity
    tests regenerate their tensors from (graph seed 1000+idx, inputs 2000+idx in blocks of 8 utterances)."""
    c = CONFIGS[cfg]
    idx = int(cfg[3:])
    mode = {"cfg2": "logden", "cfg3": "loss"}.get(cfg, "den")
    spec = dict(cfg=cfg, mode=mode, S=c["S"], K=c["K"], D=c["D"], T=c["T"], B=c["B"], variant=variant,
                graph_seed=1000 + idx, x_seed=2000 + idx + 7919 * rank, block=8, variable=c["variable"], coeff=1e-5,
                num_seed=52 + 1000 * rank)
    spec.update(over)
    return spec


def build_workload(spec: dict, start: int = 0, count: int = None):
    """(den ChainGraph, numerator ChainGraphBatch or None, x [count,T,D], lengths [count]) for utterances
    start..start+count of `spec`; only the blocks of x that are needed are generated."""
    T, D, B, blk = spec["T"], spec["D"], spec["B"], spec["block"]
    count = B - start if count is None else count
    lengths = make_lengths(B, T, spec["x_seed"], spec.get("variable", False))[start:start + count]
    b0, b1 = start // blk, (start + count + blk - 1) // blk
    x = torch.cat([input_block(T, D, spec["x_seed"], i, blk) for i in range(b0, b1)], 0)
    x = x[start - b0 * blk: start - b0 * blk + count].contiguous()
    if spec.get("x_scale"):
        x = x * float(spec["x_scale"])
    fst = random_den_fst(spec["S"], spec["K"], D, seed=spec["graph_seed"], variant=spec.get("variant", "uniform"))
    den = ChainGraph(fst, log_domain=(spec["mode"] == "logden"),
                     initial_mode=spec.get("initial_mode", "fst"), final_mode=spec.get("final_mode", "fst"))
    num = None
    if spec["mode"] == "loss":
        import copy
        num = copy.copy(make_num_batch(B, T, D, seed=spec["num_seed"]))
        if start != 0 or count != B:
            num.reorder(torch.arange(start, start + count))
            num.batch_size = count
    return den, num, x, lengths


def den_graph_for(cfg: str, variant: str = "uniform", initial_mode: str = "fst", final_mode: str = "fst") -> ChainGraph:
    c = CONFIGS[cfg]
    idx = int(cfg[3:])
    fst = random_den_fst(c["S"], c["K"], c["D"], seed=1000 + idx, variant=variant)
    return ChainGraph(fst, initial_mode=initial_mode, final_mode=final_mode, log_domain=False)
