"""`ChainFunction` / `ChainLoss`: the autograd surface of the LF-MMI loss.

Same classes, arguments, semantics and error behaviour as the reference's
`pychain/loss.py` (ChainFunction :27-87, ChainLoss :90-105).  What changed is
underneath: every full-tensor pass the reference runs in PyTorch around its
native calls is fused into the kernels --

    input.clamp(-30, 30)        loss.py:30     -> CHAIN_B200_CLAMP_X (clamped while staging a frame)
    input.exp()                 loss.py:43     -> exp applied while staging a frame into shared memory
    log_probs_grad.exp()        loss.py:77     -> posteriors are written directly (no GRAD_AS_LOG)
    pack_padded_sequence(...)   loss.py:37-40  -> not needed (kernels take the lengths)
    grad * objf_grad            loss.py:85     -> chain_b200_scale_grad, a no-op on the device when objf_grad == 1
    -(num_objf - den_objf)      loss.py:102    \\
    objf / x_lengths.sum()      loss.py:104     > `ChainLoss` is ONE autograd node: the kernels write
    grad_num + grad_den (autograd)             /  +-posteriors / sum(lengths) and one launch
                                                  (chain_b200_combine_loss) adds them and forms the loss

and a shared denominator graph is packed once and kept on the device instead
of being replicated B times on the host and uploaded per call (loss.py:99,
graph.py:101-120, chain-computation.cc:77-89).  Nothing on this path synchronises
with the host, so `ChainLoss` forward + backward can be captured in a CUDA graph
(pass the lengths as a CUDA tensor for that; see tests/test_gpu_fused_loss.py).
"""
import threading
from collections import OrderedDict

import torch
import torch.nn as nn

from .graph import ChainGraphBatch
from . import engine

__all__ = ["ChainFunction", "ChainLoss"]

_tls = threading.local()

# validated host copies of CUDA `lengths` tensors: (data_ptr, version) -> (the tensor itself, host copy).
# Holding the tensor keeps its storage alive, so the key cannot be recycled for different contents.
_lengths_cache = OrderedDict()


def _validate_lengths(input, lengths):
    """The error behaviour of `pack_padded_sequence(..., enforce_sorted=True)` (loss.py:37-39)."""
    if lengths.dim() != 1 or lengths.numel() != input.size(0):
        raise RuntimeError("Expected `len(lengths)` to be equal to batch_size, but got {} (batch_size={})"
                           .format(lengths.numel(), input.size(0)))
    if lengths.numel() and (int(lengths.min()) <= 0):
        raise RuntimeError("Length of all samples has to be greater than 0, but found an element in "
                           "'lengths' that is <= 0")
    if lengths.numel() > 1 and bool((lengths[1:] > lengths[:-1]).any()):
        raise RuntimeError("`lengths` array must be sorted in decreasing order when `enforce_sorted` is True. "
                           "You can pass `enforce_sorted=False` to pack_padded_sequence and/or pack_sequence "
                           "to sidestep this requirement if you do not need ONNX exportability.")
    if lengths.numel() and int(lengths.max()) > input.size(1):
        raise RuntimeError("a sequence length ({}) exceeds the number of frames ({})"
                           .format(int(lengths.max()), input.size(1)))


def _check_lengths(input, input_lengths):
    """Host copy of the lengths, validated like the reference does by packing (without packing anything).
    CPU lengths (what the reference's recipes pass) are checked on every call, which costs no synchronisation.
    CUDA lengths are brought to the host ONCE per tensor (cached with the tensor), so that steady-state calls --
    and CUDA-graph capture -- never synchronise."""
    lengths = input_lengths.detach()
    if not lengths.is_cuda:
        _validate_lengths(input, lengths)
        return lengths
    key = (lengths.data_ptr(), lengths._version, tuple(lengths.shape))
    hit = _lengths_cache.get(key)
    if hit is None:
        host = lengths.cpu()
        _validate_lengths(input, host)
        hit = _lengths_cache[key] = (lengths, host)
        while len(_lengths_cache) > 16:
            _lengths_cache.popitem(last=False)
    else:
        _lengths_cache.move_to_end(key)
    return hit[1]


def _scaled(ctx, grad, objf_grad):
    """grad * objf_grad (loss.py:85) without an elementwise pass in the common case: the saved gradient is scaled
    in place by chain_b200_scale_grad, which leaves memory alone when objf_grad == 1.  A second backward through the
    same node (retain_graph) must undo the first factor, which takes the slow way."""
    if getattr(ctx, "prev_objf_grad", None) is not None:
        out = grad * (objf_grad / ctx.prev_objf_grad)
        return out
    ctx.prev_objf_grad = objf_grad.detach().clone()
    return engine.scale_grad_(grad, objf_grad)


class ChainFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, input_lengths, graphs, leaky_coefficient=1e-5):
        input = input.contiguous()
        B = input.size(0)
        if B != graphs.batch_size:
            raise ValueError(
                "input batch size ({}) does not equal to graph batch size ({})"
                .format(B, graphs.batch_size)
            )
        _check_lengths(input, input_lengths)
        lengths = input_lengths.detach()

        if not graphs.log_domain:  # usually for the denominator
            g = tuple(getattr(graphs, n) for n in engine._GRAPH_NAMES_PROB)
            r = engine.forward_backward_prob(g, input, lengths, graphs.num_states, leaky_coefficient,
                                             x_is_exp=False, clamp=True)
        else:  # usually for the numerator
            g = tuple(getattr(graphs, n) for n in engine._GRAPH_NAMES_LOG)
            r = engine.forward_backward_log(g, input, lengths, graphs.num_states, grad_as_log=False, clamp=True)
        ctx.save_for_backward(r.grad)
        ctx.mark_non_differentiable(r.ok)
        ChainFunction.last_ok = r.ok  # device int32[1]; not synchronised (the reference ignores `ok`, loss.py:44,62)
        return r.objf

    @staticmethod
    def backward(ctx, objf_grad):
        input_grad, = ctx.saved_tensors
        return _scaled(ctx, input_grad, objf_grad), None, None, None


class _ChainLossFunction(torch.autograd.Function):
    """The whole of `ChainLoss.forward` (loss.py:97-105) as one autograd node: numerator (side stream) and
    denominator kernels write -gamma_num * s and +gamma_den * s (s = 1/sum(lengths) with avg) and one launch adds
    the two and forms loss = -(num - den) * s."""

    @staticmethod
    def forward(ctx, input, input_lengths, num_graphs, den_graph, leaky_coefficient, avg, overlap):
        input = input.contiguous()
        B = input.size(0)
        if B != num_graphs.batch_size:
            raise ValueError(
                "input batch size ({}) does not equal to graph batch size ({})"
                .format(B, num_graphs.batch_size)
            )
        host_lengths = _check_lengths(input, input_lengths)
        lengths = input_lengths.detach()
        scale = 1.0 / float(host_lengths.sum()) if avg else 1.0
        # stride-0 views: O(1), not a B-fold copy (loss.py:99); built once per batch size and kept with the graph
        cache = den_graph.__dict__.setdefault("_batch_cache", {})
        den_graphs = cache.get(B)
        if den_graphs is None:
            if len(cache) > 8:
                cache.clear()
            den_graphs = cache[B] = ChainGraphBatch(den_graph, B)
        gd = tuple(getattr(den_graphs, n) for n in engine._GRAPH_NAMES_PROB)
        gn = tuple(getattr(num_graphs, n) for n in (engine._GRAPH_NAMES_LOG if num_graphs.log_domain
                                                    else engine._GRAPH_NAMES_PROB))
        # numerator and denominator are independent: the numerator kernels run on a side stream, concurrently with
        # the denominator (not while a CUDA graph is being captured: the allocator bookkeeping of a second stream
        # does not belong into a capture)
        side = None
        if overlap and input.is_cuda and not torch.cuda.is_current_stream_capturing():
            side = ChainLoss._side_streams.get(input.device)
            if side is None:
                side = ChainLoss._side_streams[input.device] = torch.cuda.Stream(device=input.device)
        if num_graphs.log_domain:
            r_num = engine.forward_backward_log(gn, input, lengths, num_graphs.num_states, grad_as_log=False,
                                                clamp=True, stream=side, grad_scale=-scale)
        else:
            r_num = engine.forward_backward_prob(gn, input, lengths, num_graphs.num_states, leaky_coefficient,
                                                 x_is_exp=False, clamp=True, stream=side, grad_scale=-scale)
        r_den = engine.forward_backward_prob(gd, input, lengths, den_graphs.num_states, leaky_coefficient,
                                             x_is_exp=False, clamp=True, grad_scale=scale)
        if side is not None:
            torch.cuda.current_stream(input.device).wait_stream(side)
        loss = engine.combine_loss(r_num, r_den, scale)   # r_den.grad += r_num.grad ; loss = (den - num) * s
        ctx.save_for_backward(r_den.grad)
        ChainFunction.last_ok = r_den.ok
        ChainLoss.last_ok_num = r_num.ok
        return loss

    @staticmethod
    def backward(ctx, loss_grad):
        grad, = ctx.saved_tensors
        return _scaled(ctx, grad, loss_grad), None, None, None, None, None, None


class ChainLoss(nn.Module):
    def __init__(self, den_graph, leaky_coefficient=1e-5, avg=True):
        super(ChainLoss, self).__init__()
        self.den_graph = den_graph
        self.avg = avg
        self.leaky_coefficient = leaky_coefficient

    overlap_streams = True
    _side_streams = {}
    last_ok_num = None

    def forward(self, x, x_lengths, num_graphs):
        return _ChainLossFunction.apply(x, x_lengths, num_graphs, self.den_graph, self.leaky_coefficient,
                                        self.avg, self.overlap_streams)
