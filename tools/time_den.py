"""Time the shared-graph denominator on an arbitrary graph / batch shape (CUDA events, resident inputs).
usage: time_den.py S K D B T [variant] [G]   (G forces the group width through CHAIN_B200_G)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
S, K, D, B, T = (int(v) for v in sys.argv[1:6])
variant = sys.argv[6] if len(sys.argv) > 6 else "uniform"
if len(sys.argv) > 7:
    os.environ["CHAIN_B200_G"] = sys.argv[7]
import torch
from pychain_b200 import ChainGraph, ChainGraphBatch, synth, engine
dev = torch.device("cuda:0")
g = ChainGraph(synth.random_den_fst(S, K, D, seed=5, variant=variant))
gb = ChainGraphBatch(g, B)
gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_PROB)
x, lengths = synth.make_inputs(B, T, D, seed=9)
x = x.to(dev)
run = lambda: engine.forward_backward_prob(gt, x, lengths, g.num_states, 1e-5, x_is_exp=False, clamp=True)
for _ in range(5):
    r = run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 30
e0.record()
for _ in range(n):
    r = run()
e1.record()
torch.cuda.synchronize()
print("S %d K %d D %d B %d T %d %s G=%s: %.4f ms  objf %.6f ok %d" % (S, K, D, B, T, variant, os.environ.get("CHAIN_B200_G", "auto"),
                                                                      e0.elapsed_time(e1) / n, float(r.objf), int(r.ok)))
