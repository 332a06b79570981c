"""The drop-in is real: the UNMODIFIED reference `pychain/loss.py` (staged, never committed, next to the
reference extension by oracle/build_ref.py) runs on top of

  (a) the Python drop-in modules (`install_as_pychain()`: pychain.graph, pychain_C -> this package), and
  (b) the C++ binding of INTEGRATION.md section 3 (integration/pychain_binding.cc: pybind11 -> the C ABI),

and its ChainLoss / ChainFunction outputs are compared with the reference extension itself (oracle/_ref, CPU path).
Also: a MATERIALISED denominator batch (what the reference's own graph.py:101-120 builds with `.repeat`) is
recognised and takes the packed shared-graph kernels, so the stock layout costs the same as the stride-0 one.
"""
import importlib.util
import os
import sys
import time

import pytest
import torch

import pychain_b200
from pychain_b200 import ChainGraph, ChainGraphBatch, synth, engine
from oracle import chain_oracle as co
from oracle import build_ref

pytestmark = pytest.mark.gpu


def rel(a, b):
    return abs(float(a) - float(b)) / max(1e-30, abs(float(b)))


def _load_reference_loss(pychain_C_module):
    """Execute the staged reference loss.py with `pychain.graph` = this package's graph module and `pychain_C` =
    the given native-module replacement."""
    path = build_ref.STAGED_LOSS
    if not os.path.isfile(path):
        pytest.skip("the reference's loss.py is not staged (oracle/build_ref.py was not run where /root/reference exists)")
    saved = {k: sys.modules.get(k) for k in ("pychain", "pychain.graph", "pychain.loss", "pychain_C")}
    try:
        for k in saved:
            sys.modules.pop(k, None)
        pychain_b200.install_as_pychain()
        sys.modules["pychain_C"] = pychain_C_module
        spec = importlib.util.spec_from_file_location("reference_pychain_loss", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _case(B=12, T=40, D=48, S=80, K=500, seed=91):
    den = ChainGraph(synth.random_den_fst(S, K, D, seed=seed, variant="biphone"))
    nb = synth.make_num_batch(B, T, D, seed=seed + 1, min_phones=3, max_phones=8)
    x, lengths = synth.make_inputs(B, T, D, seed=seed + 2, variable=True)
    xc = x.clamp(-30, 30)
    gbm = ChainGraphBatch(den, B).materialize()
    if co.ref_available():
        o_den, g_den, _ = co.den_ref(gbm, xc.exp(), lengths, 1e-5)
        o_num, lg_num, _ = co.num_ref(nb, xc, lengths)
    else:
        o_den, _, g_den, _, _ = co.den_c(gbm, xc.exp(), lengths, 1e-5)
        o_num, _, lg_num, _, _ = co.num_c(nb, xc, lengths)
    n = float(lengths.sum())
    return den, nb, x, lengths, -(float(o_num) - float(o_den)) / n, (g_den - lg_num.exp()) / n


@pytest.mark.parametrize("native", ["python-shim", "cpp-binding"])
def test_unmodified_reference_loss_runs_on_the_drop_in(cuda, native):
    if native == "python-shim":
        from pychain_b200 import pychain_C as native_mod
    else:
        from integration import build_binding
        native_mod = build_binding.load()
    ref_loss = _load_reference_loss(native_mod)
    assert ref_loss.pychain_C is native_mod and ref_loss.ChainGraphBatch is pychain_b200.ChainGraphBatch
    den, nb, x, lengths, want_loss, want_grad = _case()
    xd = x.to(cuda).requires_grad_(True)
    crit = ref_loss.ChainLoss(den, leaky_coefficient=1e-5, avg=True)      # the reference's class, unmodified
    loss = crit(xd, lengths, nb)
    loss.backward()
    assert rel(loss.item(), want_loss) < 1e-5
    assert (xd.grad.cpu() - want_grad).abs().max().item() < 1e-6
    # ChainFunction alone (numerator, log domain), upstream gradient != 1
    xd2 = x.to(cuda).requires_grad_(True)
    obj = ref_loss.ChainFunction.apply(xd2, lengths, nb)
    (obj * 0.25).backward()
    assert torch.isfinite(xd2.grad).all() and xd2.grad.abs().sum().item() > 0


def test_materialised_identical_batch_takes_the_packed_kernels(cuda):
    """`ChainGraphBatch.initialized_by_one` of the reference repeats the graph B times (graph.py:101-120); the
    engine recognises such a batch (identical rows) and runs the shared-graph kernels on row 0: same bits as the
    stride-0 batch, and nowhere near the per-sequence kernel's time."""
    c = synth.CONFIGS["cfg5"]
    B, T, D, S = 96, c["T"], c["D"], c["S"]
    g = synth.den_graph_for("cfg5")
    x, lengths = synth.make_inputs(B, T, D, seed=5, variable=True)
    xd = x.to(cuda)
    shared = ChainGraphBatch(g, B)
    mat = ChainGraphBatch(g, B).materialize()
    assert mat.forward_transitions.stride(0) != 0
    gs = tuple(getattr(shared, n) for n in engine._GRAPH_NAMES_PROB)
    gm = tuple(getattr(mat, n) for n in engine._GRAPH_NAMES_PROB)

    def timed(gt, path=None, n=5):
        r = engine.forward_backward_prob(gt, xd, lengths, S, 1e-5, x_is_exp=False, clamp=True, path=path)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            r = engine.forward_backward_prob(gt, xd, lengths, S, 1e-5, x_is_exp=False, clamp=True, path=path)
        torch.cuda.synchronize()
        return r, (time.perf_counter() - t0) / n

    r_s, t_s = timed(gs)
    r_m, t_m = timed(gm)
    r_g, t_g = timed(gm, path="generic")
    assert torch.equal(r_s.objf, r_m.objf) and torch.equal(r_s.grad, r_m.grad)      # same kernel, same blob
    assert (r_g.grad - r_m.grad).abs().max().item() < 2e-5
    assert t_m < 0.5 * (t_s + t_g), (t_s, t_m, t_g)     # on the packed side of the gap (host-side row check included)
    # a batch whose rows differ must NOT be taken for shared
    mat2 = ChainGraphBatch(g, B).materialize()
    mat2.forward_transition_probs[B // 2] *= 0.5
    mat2.backward_transition_probs[B // 2] *= 0.5
    g2 = tuple(getattr(mat2, n) for n in engine._GRAPH_NAMES_PROB)
    r2 = engine.forward_backward_prob(g2, xd, lengths, S, 1e-5, x_is_exp=False, clamp=True)
    assert not torch.equal(r2.logp[B // 2], r_s.logp[B // 2]) and torch.allclose(r2.logp[0], r_s.logp[0], rtol=1e-5)
