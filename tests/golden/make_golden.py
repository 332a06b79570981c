#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference extension's CPU path
(oracle/_ref/pychain_C.so built by oracle/build_ref.py from /root/reference).

The reference repository has no tests or golden vectors;
these fixtures pin the oracle (and, on the GPU box where /root/reference does
not exist, the CUDA engine) to outputs of the reference itself.  Run where
/root/reference exists:   python tests/golden/make_golden.py
Inputs are regenerated from seeds by pychain_b200.synth, so a fixture only
stores the case description and the reference outputs.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from pychain_b200 import ChainGraph, ChainGraphBatch, synth  # noqa: E402

DEN_CASES = {
    # name: (S, K, D, B, T, variable, coeff, initial_mode, final_mode, variant, seed)
    "den_cfg1": (10, 20, 16, 2, 50, False, 1e-5, "fst", "fst", "uniform", 1001),
    "den_ragged": (10, 20, 16, 3, 50, True, 1e-5, "fst", "fst", "uniform", 1002),
    "den_leaky_ones": (10, 20, 16, 3, 50, True, 0.1, "leaky", "ones", "uniform", 1003),
    "den_skew": (37, 300, 19, 5, 33, True, 1e-5, "fst", "fst", "skew", 1004),
    "den_mid": (120, 700, 48, 4, 40, True, 1e-5, "leaky", "ones", "uniform", 1005),
}
NUM_CASES = {
    # name: (B, T, D, variable, seed)
    "num_small": (3, 50, 16, True, 3),
    "num_mid": (6, 64, 40, True, 4),
}
LOGDEN_CASES = {
    # leaky-free denominator through the log-domain entry (BASELINE cfg2 flavour)
    "logden_small": (60, 400, 24, 4, 40, True, 1006),
}


def den_inputs(case):
    S, K, D, B, T, variable, coeff, im, fm, variant, seed = case
    g = ChainGraph(synth.random_den_fst(S, K, D, seed=seed, variant=variant), initial_mode=im, final_mode=fm)
    x, lengths = synth.make_inputs(B, T, D, seed=seed + 1, variable=variable)
    return g, x.clamp(-30, 30), lengths, coeff


def num_inputs(case):
    B, T, D, variable, seed = case
    nb = synth.make_num_batch(B, T, D, seed=seed, min_phones=3, max_phones=8)
    x, lengths = synth.make_inputs(B, T, D, seed=seed + 10, variable=variable)
    return nb, x.clamp(-30, 30), lengths


def logden_inputs(case):
    S, K, D, B, T, variable, seed = case
    g = ChainGraph(synth.random_den_fst(S, K, D, seed=seed), log_domain=True)
    x, lengths = synth.make_inputs(B, T, D, seed=seed + 1, variable=variable)
    return ChainGraphBatch(g, B), x.clamp(-30, 30), lengths


def main():
    from oracle import chain_oracle as co
    if not co.ref_available():
        raise SystemExit("oracle/_ref/pychain_C.so missing: run `python oracle/build_ref.py` where /root/reference exists")
    for name, case in DEN_CASES.items():
        g, x, lengths, coeff = den_inputs(case)
        gb = ChainGraphBatch(g, x.size(0))
        objf, grad, ok = co.den_ref(gb, x.exp(), lengths, coeff)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), objf=np.float32(objf.item()), grad=grad.numpy(),
                            ok=np.bool_(ok), x_checksum=np.float64(x.double().sum().item()),
                            leaky_checksum=np.float64(g.leaky_probs.double().sum().item()))
        print(name, objf.item(), ok)
    for name, case in NUM_CASES.items():
        nb, x, lengths = num_inputs(case)
        objf, lg, ok = co.num_ref(nb, x, lengths)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), objf=np.float32(objf.item()), log_grad=lg.numpy(),
                            ok=np.bool_(ok), x_checksum=np.float64(x.double().sum().item()))
        print(name, objf.item(), ok)
    for name, case in LOGDEN_CASES.items():
        gb, x, lengths = logden_inputs(case)
        objf, lg, ok = co.num_ref(gb, x, lengths)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), objf=np.float32(objf.item()), log_grad=lg.numpy(),
                            ok=np.bool_(ok), x_checksum=np.float64(x.double().sum().item()))
        print(name, objf.item(), ok)


if __name__ == "__main__":
    main()
