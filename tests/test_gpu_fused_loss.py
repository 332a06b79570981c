"""`ChainLoss` as one autograd node: the kernels write +-posteriors/sum(lengths), one
launch adds them and forms the loss, `grad * objf_grad` is an in-place kernel that does nothing for a plain
`loss.backward()`.  No host synchronisation => the whole forward + backward captures into a CUDA graph."""
import pytest
import torch

from pychain_b200 import ChainGraph, ChainGraphBatch, ChainLoss, ChainFunction, synth, engine
from oracle import chain_oracle as co

pytestmark = pytest.mark.gpu


def _case(B=9, T=36, D=40, S=70, K=420, seed=71, variant="uniform"):
    den = ChainGraph(synth.random_den_fst(S, K, D, seed=seed, variant=variant))
    nb = synth.make_num_batch(B, T, D, seed=seed + 1, min_phones=3, max_phones=7)
    x, lengths = synth.make_inputs(B, T, D, seed=seed + 2, variable=True)
    xc = x.clamp(-30, 30)
    o_den, _, g_den, _, _ = co.den_c(ChainGraphBatch(den, B), xc.exp(), lengths, 1e-5)
    o_num, _, lg_num, _, _ = co.num_c(nb, xc, lengths)
    return den, nb, x, lengths, (o_den - o_num), (g_den - lg_num.exp())


@pytest.mark.parametrize("avg", [True, False])
def test_fused_chain_loss_value_grad_and_upstream_factor(cuda, avg):
    den, nb, x, lengths, want_loss, want_grad = _case()
    s = 1.0 / float(lengths.sum()) if avg else 1.0
    xd = x.to(cuda).requires_grad_(True)
    loss = ChainLoss(den, 1e-5, avg=avg)(xd, lengths, nb)
    assert abs(loss.item() - want_loss * s) < 1e-5 * abs(want_loss * s)
    (loss * 0.3).backward(retain_graph=True)            # upstream factor != 1: chain_b200_scale_grad does the work
    assert (xd.grad.cpu() - 0.3 * s * want_grad).abs().max().item() < 2e-5 * s + 1e-7
    xd.grad = None
    (loss * 2.0).backward()                             # a second backward through the same node
    assert (xd.grad.cpu() - 2.0 * s * want_grad).abs().max().item() < 2e-5 * 2 * s + 1e-7


def test_no_aten_elementwise_between_our_kernels(cuda):
    """The launches of a ChainLoss forward + backward are this library's own (the engine counts them):
    pdf_sort + generic (numerator), den_shared, 2 x finalize, combine, scale_grad = 7."""
    den, nb, x, lengths, _, _ = _case()
    crit = ChainLoss(den, 1e-5)
    xd = x.to(cuda).requires_grad_(True)
    crit(xd, lengths, nb).backward()                    # warm-up: packs / uploads the graphs
    torch.cuda.synchronize()
    n0 = engine.launch_count()
    xd.grad = None
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        crit(xd, lengths, nb).backward()
        torch.cuda.synchronize()
    assert engine.launch_count() - n0 == 7
    names = [e.key for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA]
    kernels = [n for n in names if "memcpy" not in n.lower() and "memset" not in n.lower()]
    # the one launch that is not ours is autograd's own seed: `loss.backward()` fills a 1-element ones tensor
    foreign = [n for n in kernels if "chainb200" not in n and "FillFunctor" not in n]
    assert not foreign, foreign


def test_cuda_graph_capture_chain_loss(cuda):
    """ChainLoss forward + backward captured once and replayed on new inputs: bit-identical to eager."""
    den, nb, x, lengths, _, _ = _case(B=6, T=30)
    crit = ChainLoss(den, 1e-5, avg=True)
    ld = lengths.to(cuda)                                # device lengths: validated once, no sync afterwards
    xs = x.to(cuda).requires_grad_(True)
    for _ in range(2):                                   # warm-up outside the capture (packing, uploads, caches)
        xs.grad = None
        crit(xs, ld, nb).backward()
    torch.cuda.synchronize()
    eager_loss, eager_grad = crit(xs, ld, nb).detach().clone(), None
    xs.grad = None
    crit(xs, ld, nb).backward()
    eager_grad = xs.grad.detach().clone()
    xs.grad = None
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        static_loss = crit(xs, ld, nb)
        static_loss.backward()
    static_grad = xs.grad
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(static_loss, eager_loss) and torch.equal(static_grad, eager_grad)
    x2, _ = synth.make_inputs(6, 30, 40, seed=999)
    with torch.no_grad():
        xs.copy_(x2.to(cuda))
    g.replay()
    torch.cuda.synchronize()
    xe = x2.to(cuda).requires_grad_(True)
    le = crit(xe, ld, nb)
    le.backward()
    assert torch.equal(static_loss, le.detach()) and torch.equal(static_grad, xe.grad)


def test_overlapped_streams_are_race_free(cuda):
    """Numerator on a side stream vs everything on one stream: identical results over many iterations with tiny
    numerators and a changing input (the buffers of one step are recycled by the allocator for the next)."""
    den, nb, x, lengths, _, _ = _case(B=16, T=20, D=24, S=40, K=200)
    crit = ChainLoss(den, 1e-5)
    for it in range(40):
        xi = (x + 0.01 * it).to(cuda)
        outs = []
        for overlap in (True, False):
            ChainLoss.overlap_streams = overlap
            try:
                xd = xi.clone().requires_grad_(True)
                loss = crit(xd, lengths, nb)
                loss.backward()
                outs.append((loss.detach().clone(), xd.grad.clone()))
            finally:
                ChainLoss.overlap_streams = True
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]), it
