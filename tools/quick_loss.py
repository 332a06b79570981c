"""Timing of the other BASELINE configs: cfg2 via the log-domain entry, cfg3 full ChainLoss."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pychain_b200 import synth, engine, ChainGraph, ChainGraphBatch, ChainLoss

def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

# cfg2: leaky-free denominator through forward_backward_log_domain (shared graph, generic kernel)
c = synth.CONFIGS["cfg2"]
g = ChainGraph(synth.random_den_fst(c["S"], c["K"], c["D"], seed=1002), log_domain=True)
gb = ChainGraphBatch(g, c["B"])
x, lengths = synth.make_inputs(c["B"], c["T"], c["D"], seed=2002)
xd = x.cuda()
gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_LOG)
ms = timeit(lambda: engine.forward_backward_log(gt, xd, lengths, gb.num_states, grad_as_log=False, clamp=True))
print(f"cfg2 log-domain den (shared-graph log kernel): {ms:.3f} ms  {c['B']*c['T']/ms/1e3:.2f} Mframes/s")

# cfg3: full ChainLoss, variable T
c = synth.CONFIGS["cfg3"]
den = synth.den_graph_for("cfg3")
nb = synth.make_num_batch(c["B"], c["T"], c["D"], seed=52)
x, lengths = synth.make_inputs(c["B"], c["T"], c["D"], seed=2003, variable=True)
crit = ChainLoss(den, 1e-5, avg=True)
xd = x.cuda().requires_grad_(True)
def step():
    xd.grad = None
    loss = crit(xd, lengths, nb)
    loss.backward()
ms = timeit(step)
F = int(lengths.sum())
print(f"cfg3 ChainLoss fwd+bwd (num S<={nb.num_states} K<={nb.num_transitions}): {ms:.3f} ms  {F/ms/1e3:.2f} Mframes/s")
gtn = tuple(getattr(nb, n) for n in engine._GRAPH_NAMES_LOG)
xc = x.cuda()
ms_n = timeit(lambda: engine.forward_backward_log(gtn, xc, lengths, nb.num_states, grad_as_log=False, clamp=True))
gbd = ChainGraphBatch(den, c["B"]); gtd = tuple(getattr(gbd, n) for n in engine._GRAPH_NAMES_PROB)
ms_d = timeit(lambda: engine.forward_backward_prob(gtd, xc, lengths, gbd.num_states, 1e-5, x_is_exp=False, clamp=True))
print(f"   numerator alone {ms_n:.3f} ms, denominator alone {ms_d:.3f} ms")
ChainLoss.overlap_streams = False
ms2 = timeit(step)
print(f"   ChainLoss without stream overlap: {ms2:.3f} ms")
ChainLoss.overlap_streams = True
ms3 = timeit(step)
print(f"   ChainLoss with stream overlap again: {ms3:.3f} ms")
