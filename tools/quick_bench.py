"""Throwaway-style timing sweep (not the judged bench): den fwd+bwd per config / G / path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pychain_b200 import synth, engine, ChainGraphBatch

def run(cfg, path, G=None, iters=5, variant="uniform", Bover=None):
    c = synth.CONFIGS[cfg]
    B = Bover or c["B"]
    g = synth.den_graph_for(cfg, variant=variant)
    gb = ChainGraphBatch(g, B)
    x, lengths = synth.make_inputs(B, c["T"], c["D"], seed=2000, variable=c["variable"])
    xd = x.cuda()
    gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_PROB)
    if G: os.environ["CHAIN_B200_G"] = str(G)
    else: os.environ.pop("CHAIN_B200_G", None)
    for _ in range(2):
        r = engine.forward_backward_prob(gt, xd, lengths, gb.num_states, 1e-5, x_is_exp=False, clamp=True, path=path)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        r = engine.forward_backward_prob(gt, xd, lengths, gb.num_states, 1e-5, x_is_exp=False, clamp=True, path=path)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    F = int(lengths.sum())
    alg = F * (12 * c["D"] + 8 * c["S"])
    print(f"{cfg} {variant} B={B} path={path} G={G}: {ms:.3f} ms  {F/ms/1e3:.2f} Mframes/s  alg {alg/ms/1e6:.1f} GB/s  objf {r.objf.item():.3f} ok {int(r.ok.item())}", flush=True)

if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg4", "cfg2", "cfg5", "cfg3"]
    for cfg in which:
        for G in (4, 2, 1):
            run(cfg, "auto", G)
        if cfg != "cfg4":
            run(cfg, "generic")
    if "cfg4" in which:
        run("cfg4", "auto", 4, variant="skew")
        run("cfg4", "generic", iters=2)
