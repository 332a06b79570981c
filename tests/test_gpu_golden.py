"""The CUDA engine against the committed reference outputs (tests/golden/, produced
by the unmodified reference extension's CPU path): every path that can run a case."""
import os

import numpy as np
import pytest
import torch

from pychain_b200 import ChainGraphBatch, engine
from golden_cases import mg, load

pytestmark = pytest.mark.gpu


def rel(a, b):
    return abs(float(a) - float(b)) / max(1e-30, abs(float(b)))


@pytest.mark.parametrize("name", sorted(mg.DEN_CASES))
@pytest.mark.parametrize("path", ["generic", "G1", "G2", "G4"])
def test_den_kernels_match_reference_golden(cuda, name, path):
    gold = load(name)
    g, x, lengths, coeff = mg.den_inputs(mg.DEN_CASES[name])
    gb = ChainGraphBatch(g, x.size(0))
    gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_PROB)
    if path != "generic":
        os.environ["CHAIN_B200_G"] = path[1]
    try:
        r = engine.forward_backward_prob(gt, x.to(cuda), lengths, gb.num_states, coeff, x_is_exp=False, clamp=True,
                                         path="generic" if path == "generic" else "auto")
    finally:
        os.environ.pop("CHAIN_B200_G", None)
    assert rel(r.objf.item(), gold["objf"]) < 5e-6          # BASELINE target: 1e-4
    assert np.abs(r.grad.cpu().numpy() - gold["grad"]).max() < 2e-5
    assert bool(r.ok.item()) == bool(gold["ok"])


@pytest.mark.parametrize("name", sorted(mg.NUM_CASES) + sorted(mg.LOGDEN_CASES))
@pytest.mark.parametrize("path", ["generic", "G1", "G4"])
def test_log_domain_kernel_matches_reference_golden(cuda, name, path):
    gold = load(name)
    graphs, x, lengths = mg.num_inputs(mg.NUM_CASES[name]) if name in mg.NUM_CASES else mg.logden_inputs(mg.LOGDEN_CASES[name])
    gt = tuple(getattr(graphs, n) for n in engine._GRAPH_NAMES_LOG)
    if path != "generic":
        os.environ["CHAIN_B200_G"] = path[1]
    try:
        r = engine.forward_backward_log(gt, x.to(cuda), lengths, graphs.num_states, grad_as_log=True,
                                        path="generic" if path == "generic" else "auto")
    finally:
        os.environ.pop("CHAIN_B200_G", None)
    assert rel(r.objf.item(), gold["objf"]) < 2e-6
    lg = r.grad.cpu().numpy()
    assert np.abs(np.exp(lg) - np.exp(gold["log_grad"])).max() < 2e-5
    pad = np.isneginf(gold["log_grad"]).all(-1)
    assert np.isneginf(lg[pad]).all()
