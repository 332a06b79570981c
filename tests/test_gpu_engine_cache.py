"""The host-side caches of the engine (scratch per stream, one small-output block, uploaded lengths) change nothing
but the host time of a call: results are bitwise those of uncached calls, a scratch buffer that has to grow or is
shared by calls of different shapes is handled, outputs of successive calls do not alias, and an in-place edit of a
host lengths tensor is seen."""
import pytest
import torch

from pychain_b200 import ChainGraph, ChainGraphBatch, ChainLoss, synth, engine

pytestmark = pytest.mark.gpu


def _den_call(cuda, B, T, D=24, S=60, K=400, seed=5):
    g = ChainGraph(synth.random_den_fst(S, K, D, seed=3, variant="uniform"))
    gb = ChainGraphBatch(g, B)
    gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_PROB)
    x, lengths = synth.make_inputs(B, T, D, seed=seed, variable=True)
    r = engine.forward_backward_prob(gt, x.to(cuda), lengths, g.num_states, 1e-5, x_is_exp=False, clamp=True)
    return r


def test_cached_scratch_gives_bitwise_the_uncached_results(cuda, monkeypatch):
    shapes = [(3, 20), (17, 64), (5, 33), (17, 64), (40, 90), (3, 20)]   # grow, shrink, grow again
    monkeypatch.setenv("CHAIN_B200_CACHE_WS", "0")
    engine.release_workspaces()
    want = []
    for B, T in shapes:
        r = _den_call(cuda, B, T)
        want.append((r.objf.clone(), r.grad.clone(), r.logp.clone(), r.ok.clone()))
    assert not engine._ws_cache
    monkeypatch.setenv("CHAIN_B200_CACHE_WS", "1")
    got = [_den_call(cuda, B, T) for B, T in shapes]        # all results alive at once: outputs must not alias
    torch.cuda.synchronize()
    assert len(engine._ws_cache) == 1
    for (o, g, lp, ok), r in zip(want, got):
        assert torch.equal(o, r.objf) and torch.equal(g, r.grad) and torch.equal(lp, r.logp) and torch.equal(ok, r.ok)
        assert int(r.ok) == 1
    ptrs = {r.grad.data_ptr() for r in got} | {r.objf.data_ptr() for r in got}
    assert len(ptrs) == 2 * len(got)
    engine.release_workspaces()
    assert not engine._ws_cache


def test_small_outputs_share_one_block(cuda):
    r = _den_call(cuda, 6, 25)
    assert r.objf.dim() == 0 and r.ok.dtype == torch.int32 and r.ok.shape == (1,) and r.logp.shape == (6,)
    base = r.objf.data_ptr()
    assert r.ok.data_ptr() == base + 4 and r.diag.data_ptr() == base + 32 and r.logp.data_ptr() == base + 64
    # diag[2:5] = {objf, valid frames, not ok}: the buffer a sharded job all-reduces
    assert float(r.stats[0]) == float(r.objf) and float(r.stats[2]) == 0.0


def test_host_lengths_are_uploaded_once_and_edits_are_seen(cuda):
    D, S, K, B, T = 24, 60, 400, 5, 30
    g = ChainGraph(synth.random_den_fst(S, K, D, seed=3))
    gb = ChainGraphBatch(g, B)
    gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_PROB)
    x, _ = synth.make_inputs(B, T, D, seed=8)
    xd = x.to(cuda)
    lengths = torch.tensor([30, 28, 20, 20, 11])
    run = lambda ln: engine.forward_backward_prob(gt, xd, ln, S, 1e-5, x_is_exp=False, clamp=True)  # noqa: E731
    engine._lengths_cache.clear()
    a = run(lengths)
    n = len(engine._lengths_cache)
    b = run(lengths)
    assert len(engine._lengths_cache) == n == 1 and torch.equal(a.grad, b.grad)
    lengths[4] = 7                                           # in place: same address, new version
    c = run(lengths)
    d = run(torch.tensor([30, 28, 20, 20, 7]))
    assert torch.equal(c.grad, d.grad) and torch.equal(c.objf, d.objf) and not torch.equal(c.objf, a.objf)
    assert float(c.stats[1]) == 30 + 28 + 20 + 20 + 7


def test_chain_loss_steps_with_the_side_stream_reuse_their_scratch(cuda):
    """Numerator on the side stream, denominator on the current one: two scratch buffers, reused by every step;
    the steps give identical results and match a run without the caches."""
    D, B, T = 40, 9, 36
    den = ChainGraph(synth.random_den_fst(70, 420, D, seed=71))
    nb = synth.make_num_batch(B, T, D, seed=72, min_phones=3, max_phones=7)
    x, lengths = synth.make_inputs(B, T, D, seed=73, variable=True)
    crit = ChainLoss(den, 1e-5)
    engine.release_workspaces()

    def step():
        xd = x.to(cuda).requires_grad_(True)
        loss = crit(xd, lengths, nb)
        loss.backward()
        return loss.detach().clone(), xd.grad.clone()

    first = step()
    assert len(engine._ws_cache) == 2
    sizes = sorted(t.numel() for t in engine._ws_cache.values())
    for _ in range(5):
        nxt = step()
        assert torch.equal(first[0], nxt[0]) and torch.equal(first[1], nxt[1])
    assert sorted(t.numel() for t in engine._ws_cache.values()) == sizes
    import os
    os.environ["CHAIN_B200_CACHE_WS"] = "0"
    try:
        plain = step()
    finally:
        del os.environ["CHAIN_B200_CACHE_WS"]
    assert torch.equal(first[0], plain[0]) and torch.equal(first[1], plain[1])
