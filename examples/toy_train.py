#!/usr/bin/env python
"""Minimal LF-MMI training loop on synthetic data with pychain_b200 (needs one CUDA device).

    python examples/toy_train.py [--steps 20]

The only pychain-specific lines are the ones a pychain user already has: build a denominator
`ChainGraph`, batch the per-utterance numerator graphs with `ChainGraphBatch`, call `ChainLoss`.
"""
import argparse
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pychain_b200 import ChainGraph, ChainLoss, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    D, B, T, F = 128, 16, 60, 40                       # pdfs, utterances, frames, feature dim
    den = ChainGraph(synth.random_den_fst(300, 1500, D, seed=1), initial_mode="leaky", final_mode="ones")
    num = synth.make_num_batch(B, T, D, seed=2, min_phones=4, max_phones=10)
    torch.manual_seed(0)
    feats = torch.randn(B, T, F, device=dev)
    lengths = torch.sort(torch.randint(T // 2, T + 1, (B,)), descending=True).values
    lengths[0] = T
    model = nn.Sequential(nn.Linear(F, 256), nn.ReLU(), nn.Linear(256, D)).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=2e-3)
    criterion = ChainLoss(den, leaky_coefficient=1e-5, avg=True)
    first = None
    for step in range(a.steps):
        opt.zero_grad(set_to_none=True)
        loss = criterion(model(feats), lengths, num)
        loss.backward()
        opt.step()
        if first is None:
            first = loss.item()
        if step % 5 == 0 or step == a.steps - 1:
            print("step %3d  LF-MMI loss per frame %.4f" % (step, loss.item()))
    assert loss.item() < first, "the loss should go down on a fixed batch"
    print("ok: %.4f -> %.4f" % (first, loss.item()))


if __name__ == "__main__":
    main()
