"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pychain_b200 import ChainGraph, ChainGraphBatch, ChainLoss, synth, engine
dev = torch.device("cuda:0")
for G in (1, 2, 4):
    os.environ["CHAIN_B200_G"] = str(G)
    for variant, D in (("uniform", 24), ("skew", 19), ("biphone", 12)):
        den = ChainGraph(synth.random_den_fst(45, 700, D, seed=3, variant=variant))
        nb = synth.make_num_batch(5, 24, D, seed=7, min_phones=3, max_phones=5)
        x, lengths = synth.make_inputs(5, 24, D, seed=11, variable=True)
        xd = x.to(dev).requires_grad_(True)
        loss = ChainLoss(den, 1e-5)(xd, lengths, nb)
        loss.backward()
        torch.cuda.synchronize()
        print(G, variant, float(loss))
