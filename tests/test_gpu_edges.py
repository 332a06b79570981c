"""Edge cases on the GPU: zero-length and one-frame utterances, a batch smaller than a
group, graphs with unreachable / arc-less states, error codes instead of crashes."""
import ctypes
import os

import pytest
import torch

from pychain_b200 import ChainGraph, ChainGraphBatch, Fst, synth, engine, _cabi
from pychain_b200 import pychain_C as new_C
from oracle import chain_oracle as co

pytestmark = pytest.mark.gpu


def rel(a, b):
    return abs(float(a) - float(b)) / max(1e-30, abs(float(b)))


@pytest.mark.parametrize("path", ["generic", "auto"])
def test_zero_length_and_single_frame(cuda, path):
    g = ChainGraph(synth.random_den_fst(20, 80, 12, seed=3))
    B, T, D = 5, 9, 12
    x, _ = synth.make_inputs(B, T, D, seed=4)
    lengths = torch.tensor([9, 1, 0, 4, 0])
    gb = ChainGraphBatch(g, B)
    o_c, lp_c, g_c, ok_c, _ = co.den_c(gb, x.clamp(-30, 30).exp(), lengths, 1e-5)
    gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_PROB)
    r = engine.forward_backward_prob(gt, x.to(cuda), lengths, gb.num_states, 1e-5, x_is_exp=False, clamp=True, path=path)
    assert rel(r.objf.item(), o_c) < 5e-6
    assert (r.logp.cpu() - lp_c).abs().max() < 1e-4
    assert (r.grad.cpu() - g_c).abs().max().item() < 2e-5
    assert r.grad[2].abs().max().item() == 0 and r.grad[4].abs().max().item() == 0
    assert int(r.ok.item()) == 1
    # T == 1
    r1 = engine.forward_backward_prob(gt[:], x[:, :1].contiguous().to(cuda), torch.tensor([1, 1, 1, 1, 1]), gb.num_states,
                                      1e-5, x_is_exp=False, clamp=True, path=path)
    o1, _, g1, _, _ = co.den_c(gb, x[:, :1].clamp(-30, 30).exp(), torch.tensor([1, 1, 1, 1, 1]), 1e-5)
    assert rel(r1.objf.item(), o1) < 5e-6 and (r1.grad.cpu() - g1).abs().max().item() < 2e-5


def test_graph_with_dead_states(cuda):
    """States without in-arcs / out-arcs and pdfs no arc uses: the rows must still be written (zeros)."""
    #   0 -> 1 -> 2 (final), 3 isolated, 4 has only an out-arc into 2
    fst = Fst(5, 0, [0, 1, 1, 4], [1, 1, 2, 2], [0, 2, 3, 5], [-0.1, -0.7, -0.7, -0.2], [-9.0, -9.0, -0.3, -9.0, -9.0])
    g = ChainGraph(fst)
    B, T, D = 3, 6, 8
    x, lengths = synth.make_inputs(B, T, D, seed=5, variable=True)
    gb = ChainGraphBatch(g, B)
    o_c, _, g_c, _, _ = co.den_c(gb, x.clamp(-30, 30).exp(), lengths, 1e-3)
    gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_PROB)
    for path in ("generic", "auto"):
        for G in ("1", "4"):
            os.environ["CHAIN_B200_G"] = G
            try:
                r = engine.forward_backward_prob(gt, x.to(cuda), lengths, 5, 1e-3, x_is_exp=False, clamp=True, path=path)
            finally:
                os.environ.pop("CHAIN_B200_G")
            assert rel(r.objf.item(), o_c) < 5e-6
            assert (r.grad.cpu() - g_c).abs().max().item() < 2e-5


def test_error_codes_not_crashes(cuda):
    lib = _cabi.lib()
    g = ChainGraph(synth.random_den_fst(10, 30, 8, seed=1))
    gb = ChainGraphBatch(g, 2)
    x, lengths = synth.make_inputs(2, 5, 8, seed=2)
    gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_PROB)
    with pytest.raises(_cabi.ChainB200Error, match="invalid argument"):   # leaky coefficient must be in (0,1), chain-computation.cc:68
        engine.forward_backward_prob(gt, x.to(cuda), lengths, 10, 0.0, x_is_exp=False, clamp=True, path="generic")
    with pytest.raises(_cabi.ChainB200Error, match="invalid argument"):
        engine.forward_backward_prob(gt, x.to(cuda), lengths, 10, 1.5, x_is_exp=False, clamp=True)
    with pytest.raises(RuntimeError, match="must be contiguous"):          # CHECK_CONTIGUOUS, pychain.cc:42-54
        new_C.forward_backward(*gt, gb.start_state, x.to(cuda).transpose(1, 2).contiguous().transpose(1, 2), None, lengths, 10, 1e-5)
    with pytest.raises(ValueError):
        engine.forward_backward_prob(gt, x.to(cuda), torch.tensor([5]), 10, 1e-5, x_is_exp=False, path="generic")
    # workspace too small / missing -> status code
    xd = x.to(cuda)
    dev = [t[0].contiguous().to(cuda) for t in gt]
    out = torch.empty(1, device=cuda); grad = torch.empty_like(xd); ok = torch.empty(1, dtype=torch.int32, device=cuda)
    ld = lengths.to(cuda)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    st = lib.chain_b200_forward_backward(*[p(t) for t in dev], p(xd), p(ld), 2, 5, 8, 10, 30, 1e-5, _cabi.GRAPH_SHARED, 1.0,
                                         p(out), None, p(grad), p(ok), None, None, 0, 0, None)
    assert st == _cabi.ERR_WORKSPACE
    # more pdfs than the pdf-sort kernel can count on chip -> UNSUPPORTED, not a launch failure
    assert lib.chain_b200_forward_backward(*[p(t) for t in dev], p(xd), p(ld), 2, 5, 60000, 10, 30, 1e-5, _cabi.GRAPH_SHARED, 1.0,
                                           p(out), None, p(grad), p(ok), None, p(grad), 1 << 40, 0, None) == _cabi.ERR_UNSUPPORTED
    torch.cuda.synchronize()


def test_graph_too_large_for_shared_memory(cuda):
    """19k states: 3S+D floats exceed 227 KB, so the general kernel keeps its rows in global memory;
    the packed shared-graph kernels still fit at G=1.  Both must match the oracle."""
    S, K, D, B, T = 19000, 60000, 40, 2, 5
    g = ChainGraph(synth.random_den_fst(S, K, D, seed=8))
    x, lengths = synth.make_inputs(B, T, D, seed=9, variable=True)
    gb = ChainGraphBatch(g, B)
    o_c, _, g_c, _, _ = co.den_c(gb, x.clamp(-30, 30).exp(), lengths, 1e-5)
    gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_PROB)
    for path in ("generic", "auto"):
        r = engine.forward_backward_prob(gt, x.to(cuda), lengths, S, 1e-5, x_is_exp=False, clamp=True, path=path)
        assert rel(r.objf.item(), o_c) < 5e-6
        assert (r.grad.cpu() - g_c).abs().max().item() < 2e-5
        assert int(r.ok.item()) == 1
    # 40k states do not fit the shared-graph kernels at all: the engine falls back by itself
    S2 = 40000
    g2 = ChainGraph(synth.random_den_fst(S2, 90000, D, seed=10))
    gb2 = ChainGraphBatch(g2, B)
    o2, _, gc2, _, _ = co.den_c(gb2, x.clamp(-30, 30).exp(), lengths, 1e-5)
    gt2 = tuple(getattr(gb2, n) for n in engine._GRAPH_NAMES_PROB)
    r = engine.forward_backward_prob(gt2, x.to(cuda), lengths, S2, 1e-5, x_is_exp=False, clamp=True)
    assert rel(r.objf.item(), o2) < 5e-6 and (r.grad.cpu() - gc2).abs().max().item() < 2e-5


def test_ok_flag_and_verbose_report(cuda, capfd):
    """Numerical trouble clears `ok` and is reported on stderr like the reference's self-check
    (chain-computation.cc:363-390); set_verbose_level(1) reports the frame-0 check even when it passes."""
    g = ChainGraph(synth.random_den_fst(12, 40, 8, seed=2))
    B, T, D = 3, 7, 8
    gb = ChainGraphBatch(g, B).materialize()
    x, lengths = synth.make_inputs(B, T, D, seed=3)
    args = lambda xx: (gb.forward_transitions, gb.forward_transition_indices, gb.forward_transition_probs,  # noqa: E731
                       gb.backward_transitions, gb.backward_transition_indices, gb.backward_transition_probs,
                       gb.leaky_probs, gb.initial_probs, gb.final_probs, gb.start_state,
                       xx.exp().to(cuda), None, lengths, gb.num_states, 1e-5)
    objf, grad, ok = new_C.forward_backward(*args(x))
    assert ok.device.type == "cpu" and ok.dtype == torch.bool and ok.is_pinned()   # filled asynchronously: no sync in the call
    torch.cuda.synchronize()
    new_C.poll_reports()
    assert bool(ok.item()) and capfd.readouterr().err == ""
    new_C.set_verbose_level(1)
    try:
        objf, grad, ok = new_C.forward_backward(*args(x))
        err = capfd.readouterr().err
        assert bool(ok.item()) and "alpha-beta product" in err and "log-prob-deriv sum" in err
    finally:
        new_C.set_verbose_level(0)
    # verbose level: EVERY frame is re-checked (chain-computation.cc:337); a frame with NaN input is named
    new_C.set_verbose_level(1)
    try:
        bad_mid = x.clone()
        bad_mid[1, 4, :] = float("nan")
        objf, grad, ok = new_C.forward_backward(*args(bad_mid))
        err = capfd.readouterr().err
        assert not bool(ok.item()) and "On time 4, log-prob-deriv sum" in err
    finally:
        new_C.set_verbose_level(0)
    bad = x.clone()
    bad[1, 2, :] = float("nan")
    objf, grad, ok = new_C.forward_backward(*args(bad))
    torch.cuda.synchronize()
    new_C.poll_reports()      # the quiet-level check is asynchronous: the failure line is printed once its copy has landed
    err = capfd.readouterr().err
    assert not bool(ok.item()) and "Excessive error detected" in err
    for path in ("generic", "auto"):  # both kernels flag it on the device
        gs = ChainGraphBatch(g, B)
        gt = tuple(getattr(gs, n) for n in engine._GRAPH_NAMES_PROB)
        r = engine.forward_backward_prob(gt, bad.to(cuda), lengths, gs.num_states, 1e-5, x_is_exp=False, clamp=True, path=path)
        assert int(r.ok.item()) == 0


def test_cuda_graph_capture(cuda):
    """Every C-ABI call is asynchronous and allocation-free, so a forward-backward can be captured in
    a CUDA graph and replayed (inputs already on the device; outputs live in the graph's pool)."""
    g = ChainGraph(synth.random_den_fst(120, 700, 32, seed=4))
    B, T, D = 6, 25, 32
    x, lengths = synth.make_inputs(B, T, D, seed=5, variable=True)
    gb = ChainGraphBatch(g, B).to(cuda)
    gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_PROB)
    xd, ld = x.to(cuda), lengths.to(cuda)
    eager = engine.forward_backward_prob(gt, xd, ld, gb.num_states, 1e-5, x_is_exp=False, clamp=True)
    torch.cuda.synchronize()
    static_x = xd.clone()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        engine.forward_backward_prob(gt, static_x, ld, gb.num_states, 1e-5, x_is_exp=False, clamp=True)  # warm-up: packs, caches
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(graph):
        r = engine.forward_backward_prob(gt, static_x, ld, gb.num_states, 1e-5, x_is_exp=False, clamp=True)
    static_x.copy_(xd * 0.5)
    graph.replay()
    torch.cuda.synchronize()
    want = engine.forward_backward_prob(gt, xd * 0.5, ld, gb.num_states, 1e-5, x_is_exp=False, clamp=True)
    assert torch.equal(r.objf, want.objf) and torch.equal(r.grad, want.grad)
    static_x.copy_(xd)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(r.objf, eager.objf) and torch.equal(r.grad, eager.grad)
