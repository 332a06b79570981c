"""`ChainFunction` / `ChainLoss`: the autograd surface of the LF-MMI loss.

Same classes, arguments, semantics and error behaviour as the reference's
`pychain/loss.py` (ChainFunction :27-87, ChainLoss :90-105).  What changed is
underneath: the four full-tensor passes the reference runs in PyTorch around
its native call are fused into the kernels --

    input.clamp(-30, 30)      loss.py:30   -> CHAIN_B200_CLAMP_X (clamped while staging a frame)
    input.exp()               loss.py:43   -> exp applied while staging a frame into shared memory
    log_probs_grad.exp()      loss.py:77   -> posteriors are written directly (no GRAD_AS_LOG)
    pack_padded_sequence(...) loss.py:37-40-> not needed (kernels take the lengths)

and a shared denominator graph is packed once and kept on the device instead
of being replicated B times on the host and uploaded per call (loss.py:99,
graph.py:101-120, chain-computation.cc:77-89).
"""
import threading

import torch
import torch.nn as nn

from .graph import ChainGraphBatch
from . import engine

__all__ = ["ChainFunction", "ChainLoss"]

_tls = threading.local()


def _check_lengths(input, input_lengths):
    """The error behaviour of `pack_padded_sequence(..., enforce_sorted=True)`
    (loss.py:37-39) without packing anything."""
    lengths = input_lengths.detach().cpu()
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
    return lengths


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
        lengths = _check_lengths(input, input_lengths)

        # ChainLoss overlaps numerator and denominator: with a side stream set, the kernels are launched
        # there (outputs and workspace still come from the caller's stream pool; the caller joins the
        # side stream before using the results), and the autograd node stays on the caller's stream.
        side = getattr(_tls, "side_stream", None) if input.is_cuda else None
        if not graphs.log_domain:  # usually for the denominator
            g = tuple(getattr(graphs, n) for n in engine._GRAPH_NAMES_PROB)
            r = engine.forward_backward_prob(g, input, lengths, graphs.num_states, leaky_coefficient,
                                             x_is_exp=False, clamp=True, stream=side)
        else:  # usually for the numerator
            g = tuple(getattr(graphs, n) for n in engine._GRAPH_NAMES_LOG)
            r = engine.forward_backward_log(g, input, lengths, graphs.num_states, grad_as_log=False, clamp=True,
                                            stream=side)
        ctx.save_for_backward(r.grad)
        ctx.mark_non_differentiable(r.ok)
        ChainFunction.last_ok = r.ok  # device int32[1]; not synchronised (the reference ignores `ok`, loss.py:44,62)
        return r.objf

    @staticmethod
    def backward(ctx, objf_grad):
        input_grad, = ctx.saved_tensors
        input_grad = torch.mul(input_grad, objf_grad)

        return input_grad, None, None, None


class ChainLoss(nn.Module):
    def __init__(self, den_graph, leaky_coefficient=1e-5, avg=True):
        super(ChainLoss, self).__init__()
        self.den_graph = den_graph
        self.avg = avg
        self.leaky_coefficient = leaky_coefficient

    # numerator and denominator are independent: the numerator kernels run on a side stream,
    # concurrently with the denominator (they fit next to it / on the SMs it leaves idle)
    overlap_streams = True
    _side_streams = {}

    def forward(self, x, x_lengths, num_graphs):
        batch_size = x.size(0)
        den_graphs = ChainGraphBatch(self.den_graph, batch_size)  # stride-0 views: O(1), not a B-fold copy
        if self.overlap_streams and x.is_cuda:
            main = torch.cuda.current_stream(x.device)
            side = ChainLoss._side_streams.get(x.device)
            if side is None:
                side = ChainLoss._side_streams[x.device] = torch.cuda.Stream(device=x.device)
            _tls.side_stream = side
            try:
                num_objf = ChainFunction.apply(x, x_lengths, num_graphs)  # kernels go to `side`
            finally:
                _tls.side_stream = None
            den_objf = ChainFunction.apply(x, x_lengths, den_graphs, self.leaky_coefficient)
            main.wait_stream(side)
        else:
            den_objf = ChainFunction.apply(x, x_lengths, den_graphs, self.leaky_coefficient)
            num_objf = ChainFunction.apply(x, x_lengths, num_graphs)
        objf = -(num_objf - den_objf)
        if self.avg:
            objf = objf / x_lengths.sum()
        return objf
