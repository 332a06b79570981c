"""GPU parity of the packed shared-graph denominator kernels
(chain_b200_den_forward_backward) for every group width G, against the C oracle
at oracle-sized cases and against the generic kernel + size-independent
properties at BASELINE sizes."""
import os

import pytest
import torch

from pychain_b200 import ChainGraph, ChainGraphBatch, ChainLoss, ChainFunction, synth, engine
from oracle import chain_oracle as co

pytestmark = pytest.mark.gpu

OBJF_RTOL = 2e-6
GRAD_ATOL = 2e-5


def rel(a, b):
    return abs(float(a) - float(b)) / max(1e-30, abs(float(b)))


@pytest.fixture(params=[1, 2, 4])
def group_width(request):
    os.environ["CHAIN_B200_G"] = str(request.param)
    yield request.param
    os.environ.pop("CHAIN_B200_G", None)


CASES = [
    dict(S=10, K=20, D=16, B=2, T=50, variable=False, coeff=1e-5),                       # BASELINE cfg1
    dict(S=10, K=20, D=16, B=7, T=50, variable=True, coeff=1e-5),                        # ragged, B % G != 0
    dict(S=10, K=20, D=16, B=3, T=50, variable=True, coeff=0.1, initial_mode="leaky", final_mode="ones"),
    dict(S=37, K=300, D=19, B=5, T=33, variable=True, coeff=1e-5, variant="skew"),       # D % 4 != 0: non-TMA x path
    dict(S=301, K=1500, D=64, B=9, T=40, variable=True, coeff=1e-5),                     # S % 4 != 0
    dict(S=1700, K=60000, D=120, B=5, T=12, variable=True, coeff=1e-5, variant="biphone"),  # structured, in-degree 35: split rows
]


@pytest.mark.parametrize("case", CASES)
def test_den_shared_vs_oracle(cuda, case, group_width):
    case = dict(case)
    coeff = case.pop("coeff")
    B, T, D, variable = case.pop("B"), case.pop("T"), case["D"], case.pop("variable")
    g = ChainGraph(synth.random_den_fst(case.pop("S"), case.pop("K"), case.pop("D"), seed=31, variant=case.pop("variant", "uniform")), **case)
    x, lengths = synth.make_inputs(B, T, D, seed=32, variable=variable)
    x = x * 6.0  # the clamp bites
    gb = ChainGraphBatch(g, B)
    o_c, lp_c, g_c, ok_c, diag_c = co.den_c(gb, x.clamp(-30, 30).exp(), lengths, coeff)
    gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_PROB)
    # fused: raw input, clamp + exp in the kernel
    r = engine.forward_backward_prob(gt, x.to(cuda), lengths, gb.num_states, coeff, x_is_exp=False, clamp=True)
    assert rel(r.objf.item(), o_c) < 5e-6
    assert (r.logp.cpu() - lp_c).abs().max() < 2e-5 * max(1.0, lp_c.abs().max().item())
    assert (r.grad.cpu() - g_c).abs().max().item() < GRAD_ATOL
    assert int(r.ok.item()) == int(ok_c) == 1
    assert torch.allclose(r.diag.cpu(), diag_c, rtol=1e-4)
    for b in range(B):
        if lengths[b] < T:
            assert r.grad[b, int(lengths[b]):].abs().max().item() == 0.0
    # reference-compatible: exp'ed input
    r2 = engine.forward_backward_prob(gt, x.clamp(-30, 30).exp().to(cuda), lengths, gb.num_states, coeff, x_is_exp=True)
    assert rel(r2.objf.item(), o_c) < OBJF_RTOL
    assert (r2.grad.cpu() - g_c).abs().max().item() < GRAD_ATOL


def test_den_shared_lengths_unsorted_and_short(cuda, group_width):
    """The kernels do not need sorted lengths (only ChainFunction mirrors the
    reference's check); utterances inside one group may differ wildly."""
    g = ChainGraph(synth.random_den_fst(40, 200, 24, seed=41))
    B, T, D = 6, 30, 24
    x, _ = synth.make_inputs(B, T, D, seed=42)
    lengths = torch.tensor([3, 30, 1, 17, 30, 2])
    gb = ChainGraphBatch(g, B)
    o_c, lp_c, g_c, _, _ = co.den_c(gb, x.clamp(-30, 30).exp(), lengths, 1e-5)
    gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_PROB)
    r = engine.forward_backward_prob(gt, x.to(cuda), lengths, gb.num_states, 1e-5, x_is_exp=False, clamp=True)
    assert rel(r.objf.item(), o_c) < 5e-6
    assert (r.grad.cpu() - g_c).abs().max().item() < GRAD_ATOL


@pytest.mark.parametrize("cfg", ["cfg2", "cfg5"])
def test_den_shared_vs_generic_baseline_sizes(cuda, cfg):
    """At BASELINE sizes the oracle is too slow for a unit test; the two
    independent CUDA implementations must agree and the posteriors must be
    normalised (sum_n gamma(t,n) = 1 on valid frames, alpha'.beta' = 1)."""
    c = synth.CONFIGS[cfg]
    B = min(c["B"], 96)
    g = synth.den_graph_for(cfg)
    gb = ChainGraphBatch(g, B)
    x, lengths = synth.make_inputs(B, c["T"], c["D"], seed=2000, variable=True)
    gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_PROB)
    xd = x.to(cuda)
    ra = engine.forward_backward_prob(gt, xd, lengths, gb.num_states, 1e-5, x_is_exp=False, clamp=True, path="generic")
    for G in (1, 2, 4):
        os.environ["CHAIN_B200_G"] = str(G)
        try:
            rb = engine.forward_backward_prob(gt, xd, lengths, gb.num_states, 1e-5, x_is_exp=False, clamp=True)
        finally:
            os.environ.pop("CHAIN_B200_G", None)
        assert rel(rb.objf.item(), ra.objf.item()) < 2e-6
        assert (rb.grad - ra.grad).abs().max().item() < GRAD_ATOL
        rows = rb.grad.sum(-1).cpu()
        for b in range(B):
            L = int(lengths[b])
            assert (rows[b, :L] - 1).abs().max().item() < 2e-4
            assert rows[b, L:].abs().max().item() == 0.0 if L < c["T"] else True
        assert abs(rb.diag[0].item() - B) < 1e-3 * B and int(rb.ok.item()) == 1


def test_chain_loss_end_to_end(cuda):
    """ChainLoss forward + backward (BASELINE cfg3 flavour, small) against the
    oracle: loss = -(num - den)/sum(len), d loss/dx = (gamma_den - gamma_num)/sum(len)."""
    S, K, D, B, T = 60, 300, 32, 6, 40
    den = ChainGraph(synth.random_den_fst(S, K, D, seed=51))
    nb = synth.make_num_batch(B, T, D, seed=52, min_phones=3, max_phones=8)
    x, lengths = synth.make_inputs(B, T, D, seed=53, variable=True)
    xc = x.clamp(-30, 30)
    o_den, _, g_den, _, _ = co.den_c(ChainGraphBatch(den, B), xc.exp(), lengths, 1e-5)
    o_num, _, lg_num, _, _ = co.num_c(nb, xc, lengths)
    want_loss = -(o_num - o_den) / float(lengths.sum())
    want_grad = (g_den - lg_num.exp()) / float(lengths.sum())

    xd = x.to(cuda).requires_grad_(True)
    crit = ChainLoss(den, leaky_coefficient=1e-5, avg=True)
    loss = crit(xd, lengths, nb)
    loss.backward()
    assert rel(loss.item(), want_loss) < 1e-5
    assert (xd.grad.cpu() - want_grad).abs().max().item() < 1e-6
    # ChainFunction alone, with a non-default upstream gradient
    xd2 = x.to(cuda).requires_grad_(True)
    obj = ChainFunction.apply(xd2, lengths, nb)
    (obj * 0.5).backward()
    assert (xd2.grad.cpu() - 0.5 * lg_num.exp()).abs().max().item() < GRAD_ATOL


def test_cfg4_full_size_properties(cuda):
    """BASELINE cfg4 at full size (S=5000, K=40000, B=512, T=200, D=1000): the oracle would need
    minutes, so check what does not depend on size -- posteriors of every valid frame sum to one,
    alpha'.beta' = 1 at frame 0, padded frames are exactly zero, group widths agree, and a slice of the
    batch equals the same utterances run alone."""
    c = synth.CONFIGS["cfg4"]
    B, T, D, S = c["B"], c["T"], c["D"], c["S"]
    g = synth.den_graph_for("cfg4")
    gb = ChainGraphBatch(g, B)
    x, lengths = synth.make_inputs(B, T, D, seed=2004, variable=True)
    xd = x.to(cuda)
    gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_PROB)
    r4 = engine.forward_backward_prob(gt, xd, lengths, S, 1e-5, x_is_exp=False, clamp=True)
    assert int(r4.ok.item()) == 1 and torch.isfinite(r4.objf).item()
    rows = r4.grad.sum(-1)                                   # [B,T]
    mask = torch.arange(T, device=cuda)[None, :] < lengths.to(cuda)[:, None]
    assert (rows[mask] - 1).abs().max().item() < 5e-4
    assert r4.grad[~mask].abs().max().item() == 0.0
    assert abs(r4.diag[0].item() - B) < 1e-3 * B and abs(r4.diag[1].item() - B) < 1e-3 * B
    assert abs(r4.logp.sum().item() - r4.objf.item()) < 1e-3 * abs(r4.objf.item())
    os.environ["CHAIN_B200_G"] = "2"
    try:
        r2 = engine.forward_backward_prob(gt, xd, lengths, S, 1e-5, x_is_exp=False, clamp=True)
    finally:
        os.environ.pop("CHAIN_B200_G", None)
    assert rel(r2.objf.item(), r4.objf.item()) < 2e-6
    assert (r2.grad - r4.grad).abs().max().item() < GRAD_ATOL
    # utterances 100..107 alone (another grouping, another position in the batch): same numbers
    sub = slice(100, 108)
    gb8 = ChainGraphBatch(g, 8)
    gt8 = tuple(getattr(gb8, n) for n in engine._GRAPH_NAMES_PROB)
    r8 = engine.forward_backward_prob(gt8, xd[sub].contiguous(), lengths[sub], S, 1e-5, x_is_exp=False, clamp=True)
    assert (r8.logp - r4.logp[sub]).abs().max().item() < 2e-3
    assert (r8.grad - r4.grad[sub]).abs().max().item() < GRAD_ATOL


def test_bitwise_reproducible(cuda):
    """No atomics, fixed reduction orders: two runs give bit-identical objf and gradient
    (the reference's CUDA denominator gradient is stochastic, chain-kernels.cu:53-87)."""
    g = ChainGraph(synth.random_den_fst(200, 1600, 40, seed=61, variant="skew"))
    B, T, D = 13, 30, 40
    x, lengths = synth.make_inputs(B, T, D, seed=62, variable=True)
    gb = ChainGraphBatch(g, B)
    gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_PROB)
    nb = synth.make_num_batch(B, T, D, seed=63, min_phones=3, max_phones=8)
    gn = tuple(getattr(nb, n) for n in engine._GRAPH_NAMES_LOG)
    xd = x.to(cuda)
    for fn in (lambda: engine.forward_backward_prob(gt, xd, lengths, gb.num_states, 1e-5, x_is_exp=False, clamp=True),
               lambda: engine.forward_backward_prob(gt, xd, lengths, gb.num_states, 1e-5, x_is_exp=False, clamp=True, path="generic"),
               lambda: engine.forward_backward_log(gn, xd, lengths, nb.num_states, grad_as_log=False, clamp=True)):
        a, b = fn(), fn()
        assert torch.equal(a.objf, b.objf) and torch.equal(a.grad, b.grad) and torch.equal(a.logp, b.logp)
