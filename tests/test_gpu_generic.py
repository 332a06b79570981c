"""GPU parity of the per-sequence-graph kernels (chain_b200_forward_backward /
_log_domain) against the oracle (C restatement) and, when the prebuilt
reference extension travelled with the snapshot, against the reference's own
CPU and CUDA paths.  Tolerances: objf rel <= 1e-4 required
(we assert 2e-6), grad max-abs <= 2e-5 vs the exact CPU paths; vs the
reference CUDA denominator path 3e-4 because that path stochastically rounds
contributions below 2^-14 (chain-kernels.cu:53-87)."""
import pytest
import torch

from pychain_b200 import ChainGraph, ChainGraphBatch, synth, engine
from pychain_b200 import pychain_C as new_C
from oracle import chain_oracle as co

pytestmark = pytest.mark.gpu

OBJF_RTOL = 2e-6
GRAD_ATOL = 2e-5


def rel(a, b):
    return abs(float(a) - float(b)) / max(1e-30, abs(float(b)))


def _den_case(S, K, D, B, T, variable, seed, initial_mode="fst", final_mode="fst", variant="uniform"):
    g = ChainGraph(synth.random_den_fst(S, K, D, seed=seed, variant=variant), initial_mode=initial_mode, final_mode=final_mode)
    x, lengths = synth.make_inputs(B, T, D, seed=seed + 1, variable=variable)
    return g, x.clamp(-30, 30), lengths


@pytest.mark.parametrize("case", [
    dict(S=10, K=20, D=16, B=2, T=50, variable=False, coeff=1e-5),                      # BASELINE cfg1
    dict(S=10, K=20, D=16, B=3, T=50, variable=True, coeff=1e-5),
    dict(S=10, K=20, D=16, B=3, T=50, variable=True, coeff=0.1, initial_mode="leaky", final_mode="ones"),
    dict(S=37, K=300, D=19, B=5, T=33, variable=True, coeff=1e-5, variant="skew"),
    dict(S=300, K=1500, D=64, B=4, T=40, variable=True, coeff=1e-5),
])
@pytest.mark.parametrize("shared", [True, False])
def test_den_generic_vs_oracle(cuda, case, shared):
    case = dict(case)
    coeff = case.pop("coeff")
    g, x, lengths = _den_case(seed=11, **case)
    B = x.size(0)
    gb = ChainGraphBatch(g, B)
    o_c, lp_c, g_c, ok_c, diag_c = co.den_c(gb, x.exp(), lengths, coeff)
    if not shared:
        gb.materialize()
    gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_PROB)
    # (a) the reference-compatible call: exp'ed input
    r = engine.forward_backward_prob(gt, x.exp().to(cuda), lengths, gb.num_states, coeff, x_is_exp=True, path="generic")
    assert rel(r.objf.item(), o_c) < OBJF_RTOL
    assert (r.logp.cpu() - lp_c).abs().max() < 1e-4 * max(1.0, lp_c.abs().max().item())
    assert (r.grad.cpu() - g_c).abs().max().item() < GRAD_ATOL
    assert int(r.ok.item()) == int(ok_c) == 1
    assert torch.allclose(r.diag.cpu(), diag_c, rtol=1e-4)
    for b in range(B):  # padded frames are exactly zero
        assert r.grad[b, int(lengths[b]):].abs().max().item() == 0.0 if lengths[b] < x.size(1) else True
    # (b) the fused call: raw input, clamp + exp inside the kernel
    xr, _ = synth.make_inputs(B, x.size(1), x.size(2), seed=12, variable=False)
    xr = xr * 8.0  # make the clamp bite
    o2, _, g2, _, _ = co.den_c(gb if shared else gb, xr.clamp(-30, 30).exp(), lengths, coeff)
    r2 = engine.forward_backward_prob(gt, xr.to(cuda), lengths, gb.num_states, coeff, x_is_exp=False, clamp=True, path="generic")
    assert rel(r2.objf.item(), o2) < 5e-6
    assert (r2.grad.cpu() - g2).abs().max().item() < GRAD_ATOL


@pytest.mark.parametrize("B,T,D,variable", [(3, 50, 16, True), (8, 64, 40, True), (2, 30, 16, False)])
def test_num_generic_vs_oracle(cuda, B, T, D, variable):
    nb = synth.make_num_batch(B, T, D, seed=3, min_phones=3, max_phones=8)
    x, lengths = synth.make_inputs(B, T, D, seed=5, variable=variable)
    x = x.clamp(-30, 30)
    o_c, lp_c, lg_c, ok_c, diag_c = co.num_c(nb, x, lengths)
    gt = tuple(getattr(nb, n) for n in engine._GRAPH_NAMES_LOG)
    r = engine.forward_backward_log(gt, x.to(cuda), lengths, nb.num_states, grad_as_log=True)
    assert rel(r.objf.item(), o_c) < OBJF_RTOL
    lg = r.grad.cpu()
    assert (lg.exp() - lg_c.exp()).abs().max().item() < GRAD_ATOL
    # log values agree where there is non-negligible mass; -inf pattern agrees on padded frames
    m = lg_c > -20
    assert (lg[m] - lg_c[m]).abs().max().item() < 1e-3
    for b in range(B):
        if lengths[b] < T:
            assert torch.isneginf(lg[b, int(lengths[b]):]).all()
    assert int(r.ok.item()) == int(ok_c) == 1
    # fused variant: posteriors directly
    r2 = engine.forward_backward_log(gt, x.to(cuda), lengths, nb.num_states, grad_as_log=False)
    assert (r2.grad.cpu() - lg_c.exp()).abs().max().item() < GRAD_ATOL
    # each valid frame's posteriors sum to one
    rows = r2.grad.cpu().sum(-1)
    for b in range(B):
        assert (rows[b, :int(lengths[b])] - 1).abs().max().item() < 1e-4


@pytest.mark.parametrize("path", ["G1", "G4"])
def test_log_domain_shared_graph_with_split_rows(cuda, path):
    """A skewed graph (one state collects a large share of the arcs) through the log-domain shared
    kernel: heavy rows are split and their (max, sum-exp) pairs merged."""
    import os
    fst = synth.random_den_fst(200, 3000, 24, seed=1003, variant="skew")
    g = ChainGraph(fst, log_domain=True)
    B, T, D = 5, 20, 24
    gb = ChainGraphBatch(g, B)
    x, lengths = synth.make_inputs(B, T, D, seed=10, variable=True)
    x = x.clamp(-30, 30)
    o_c, lp_c, lg_c, _, _ = co.num_c(gb, x, lengths)
    gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_LOG)
    os.environ["CHAIN_B200_G"] = path[1]
    try:
        r = engine.forward_backward_log(gt, x.to(cuda), lengths, gb.num_states, grad_as_log=False)
    finally:
        os.environ.pop("CHAIN_B200_G", None)
    assert rel(r.objf.item(), o_c) < OBJF_RTOL
    assert (r.grad.cpu() - lg_c.exp()).abs().max().item() < GRAD_ATOL
    assert int(r.ok.item()) == 1


@pytest.mark.parametrize("path", ["generic", "G1", "G2", "G4"])
def test_log_domain_shared_graph(cuda, path):
    """BASELINE cfg2 flavour: a (leaky-free) denominator graph run through the log-domain entry,
    shared by the batch: the general kernel and the packed shared-graph log kernel (every G)."""
    import os
    fst = synth.random_den_fst(61, 400, 23, seed=1002)   # D % 4 != 0, S % 4 != 0
    g = ChainGraph(fst, log_domain=True)
    B, T, D = 7, 40, 23
    gb = ChainGraphBatch(g, B)
    x, lengths = synth.make_inputs(B, T, D, seed=9, variable=True)
    lengths[-1] = 0
    x = x * 5.0
    o_c, lp_c, lg_c, ok_c, _ = co.num_c(gb, x.clamp(-30, 30), lengths)
    gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_LOG)
    if path != "generic":
        os.environ["CHAIN_B200_G"] = path[1]
    try:
        r = engine.forward_backward_log(gt, x.to(cuda), lengths, gb.num_states, grad_as_log=True, clamp=True,
                                        path="generic" if path == "generic" else "auto")
        r2 = engine.forward_backward_log(gt, x.to(cuda), lengths, gb.num_states, grad_as_log=False, clamp=True,
                                         path="generic" if path == "generic" else "auto")
    finally:
        os.environ.pop("CHAIN_B200_G", None)
    assert rel(r.objf.item(), o_c) < OBJF_RTOL
    assert (r.logp.cpu() - lp_c).abs().max() < 1e-4 * max(1.0, lp_c.abs().max().item())
    assert (r.grad.cpu().exp() - lg_c.exp()).abs().max().item() < GRAD_ATOL
    assert (r2.grad.cpu() - lg_c.exp()).abs().max().item() < GRAD_ATOL
    m = lg_c > -20
    assert (r.grad.cpu()[m] - lg_c[m]).abs().max().item() < 1e-3
    for b in range(B):
        L = int(lengths[b])
        if L < T:
            assert torch.isneginf(r.grad[b, L:]).all() and r2.grad[b, L:].abs().max().item() == 0
    assert int(r.ok.item()) == 1


@pytest.mark.skipif(not co.ref_available(), reason="prebuilt reference extension (oracle/_ref) not present")
def test_against_reference_cpu_and_cuda(cuda):
    """The reference's own extension, unmodified, on the same tensors: its exact
    CPU path and its CUDA path (stochastic den gradient)."""
    g, x, lengths = _den_case(S=50, K=300, D=32, B=6, T=45, variable=True, seed=21)
    B = x.size(0)
    gb = ChainGraphBatch(g, B).materialize()
    ex = x.exp()
    o_cpu, g_cpu, ok_cpu = co.den_ref(gb, ex, lengths, 1e-5)
    o_cuda, g_cuda, ok_cuda = co.den_ref(gb, ex.to(cuda), lengths, 1e-5)
    mine = new_C.forward_backward(
        gb.forward_transitions, gb.forward_transition_indices, gb.forward_transition_probs,
        gb.backward_transitions, gb.backward_transition_indices, gb.backward_transition_probs,
        gb.leaky_probs, gb.initial_probs, gb.final_probs, gb.start_state,
        ex.to(cuda), co._batch_sizes(lengths, x.size(1)), lengths, gb.num_states, 1e-5)
    assert rel(mine[0].item(), o_cpu.item()) < OBJF_RTOL
    assert rel(mine[0].item(), o_cuda.item()) < 1e-5
    assert (mine[1].cpu() - g_cpu).abs().max().item() < GRAD_ATOL
    assert (mine[1].cpu() - g_cuda.cpu()).abs().max().item() < 3e-4
    assert bool(mine[2].item()) and ok_cpu and ok_cuda and mine[2].device.type == "cpu" and mine[2].dtype == torch.bool

    nb = synth.make_num_batch(B, x.size(1), x.size(2), seed=4, min_phones=3, max_phones=8)
    o_cpu, lg_cpu, _ = co.num_ref(nb, x, lengths)
    o_cuda, lg_cuda, _ = co.num_ref(nb, x.to(cuda), lengths)
    mine = new_C.forward_backward_log_domain(
        nb.forward_transitions, nb.forward_transition_indices, nb.forward_transition_probs,
        nb.backward_transitions, nb.backward_transition_indices, nb.backward_transition_probs,
        nb.initial_probs, nb.final_probs, nb.start_state,
        x.to(cuda), co._batch_sizes(lengths, x.size(1)), lengths, nb.num_states)
    assert rel(mine[0].item(), o_cpu.item()) < OBJF_RTOL
    assert rel(mine[0].item(), o_cuda.item()) < OBJF_RTOL
    assert (mine[1].cpu().exp() - lg_cpu.exp()).abs().max().item() < GRAD_ATOL
    assert (mine[1].cpu().exp() - lg_cuda.cpu().exp()).abs().max().item() < GRAD_ATOL
