"""Host-side launcher logic above the C ABI (include/chain_b200.h).

PyTorch is used for device memory (outputs, workspace), the current stream and
nothing else; all arithmetic happens in libchain_b200.so.  There is no CPU path:
a CPU `x` raises.
"""
from __future__ import annotations

import ctypes
import os
from collections import OrderedDict

import torch

from . import _cabi
from ._cabi import GRAPH_SHARED, X_IS_EXP, CLAMP_X, GRAD_AS_LOG, ChainB200Error

__all__ = ["forward_backward_prob", "forward_backward_log", "prepare_den_graph", "launch_count", "Result",
           "combine_loss", "scale_grad_"]

_GRAPH_NAMES_PROB = ("forward_transitions", "forward_transition_indices", "forward_transition_probs",
                     "backward_transitions", "backward_transition_indices", "backward_transition_probs",
                     "leaky_probs", "initial_probs", "final_probs")
_GRAPH_NAMES_LOG = tuple(n for n in _GRAPH_NAMES_PROB if n != "leaky_probs")

# launches issued by this process through the engine (bench.py's gpu_launches)
_launches = 0


def launch_count():
    return _launches


class Result(object):
    """Outputs of one forward-backward: all device tensors, nothing synchronised."""
    __slots__ = ("objf", "logp", "grad", "ok", "diag", "stats")

    def __init__(self, objf, logp, grad, ok, diag):
        # diag: the 8 floats finalize_kernel writes; [0:2] = the reference's self-check sums,
        # [2:5] = {objf, valid frames, not ok}: the one buffer a batch-sharded job all-reduces
        self.objf, self.logp, self.grad, self.ok = objf, logp, grad, ok
        self.diag, self.stats = diag[:2], diag[2:5]


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _require_cuda(x, what):
    if not x.is_cuda:
        raise ChainB200Error(
            "{}: nnet output is on '{}'. pychain_b200 is a CUDA (sm_100a) engine with no CPU fallback."
            .format(what, x.device))
    if x.dtype != torch.float32:
        raise ChainB200Error("{}: nnet output must be float32, got {}".format(what, x.dtype))
    if not x.is_contiguous():
        raise RuntimeError("{} must be contiguous".format(what))  # CHECK_CONTIGUOUS, pychain.cc:24


def _is_shared(tensors, B):
    """True when every graph tensor is a stride-0 batch view, i.e. ONE graph for the whole batch
    (`ChainGraphBatch(graph, B)`).  A batch of one is shared in the trivial sense."""
    return B == 1 or all(t.stride(0) == 0 for t in tensors)


def _record_side_stream(tstream, tensors):
    """Buffers allocated from the CURRENT stream's pool but used by kernels on `tstream`: tell the caching
    allocator, or it may hand a freed block to the current stream while the side-stream kernel still runs."""
    for t in tensors:
        if t is not None and t.is_cuda:
            t.record_stream(tstream)


def _graph_to_device(tensors, device, shared, names):
    """Graph tensors on `device` with the layout the C ABI wants: row 0 only
    (contiguous) when shared, the contiguous [B,...] batch otherwise."""
    out = []
    for t, name in zip(tensors, names):
        want = torch.int32 if "transition" in name and "probs" not in name else torch.float32
        if t.dtype != want:
            raise RuntimeError("{} must be {}, got {}".format(name, want, t.dtype))
        if shared:
            t = t[0]
        elif not t.is_contiguous():
            raise RuntimeError("{} must be contiguous".format(name))
        t = t.to(device, non_blocking=True).contiguous()
        out.append(t)
    return out


# graph tensors already uploaded: key = (data_ptrs, versions, device) -> device tensors.
# Keeps the per-call `.cuda()` of the reference (chain-computation.cc:77-89) out of the hot path.
_upload_cache = OrderedDict()
_UPLOAD_CACHE_MAX = 16


def _cached_upload(tensors, device, shared, names):
    if all(t.is_cuda for t in tensors):
        return _graph_to_device(tensors, device, shared, names)
    key = (tuple((t.data_ptr(), t._version, tuple(t.shape), t.stride(0)) for t in tensors), str(device), shared)
    hit = _upload_cache.get(key)
    if hit is not None:
        _upload_cache.move_to_end(key)
        return hit[1]
    dev = _graph_to_device(tensors, device, shared, names)
    _upload_cache[key] = (tensors, dev)  # hold the sources so data_ptr keys stay unique
    while len(_upload_cache) > _UPLOAD_CACHE_MAX:
        _upload_cache.popitem(last=False)
    return dev


_lengths_cache = OrderedDict()   # (data_ptr, version, numel, dtype, device) -> (the host tensor, its device copy)


def _lengths_dev(lengths, device, B):
    """int64 lengths on the device.  Host lengths (what the reference's recipes pass) are uploaded once per tensor
    *content* -- the numerator and the denominator call of one loss, and every step that reuses the tensor, share the
    copy (an in-place edit bumps `_version`; holding the source keeps its address from being recycled)."""
    if lengths.numel() != B:
        raise ValueError("sequence_lengths has {} entries for batch size {}".format(lengths.numel(), B))
    if lengths.is_cuda or torch.cuda.is_current_stream_capturing():
        return lengths.to(device=device, dtype=torch.int64, non_blocking=True).contiguous()
    key = (lengths.data_ptr(), lengths._version, lengths.numel(), lengths.dtype, str(device))
    hit = _lengths_cache.get(key)
    if hit is not None:
        _lengths_cache.move_to_end(key)
        return hit[1]
    dev = lengths.to(device=device, dtype=torch.int64, non_blocking=True).contiguous()
    _lengths_cache[key] = (lengths, dev)
    while len(_lengths_cache) > 16:
        _lengths_cache.popitem(last=False)
    return dev


# Scratch (alpha spill, per-frame totals) is internal to one forward-backward and dead when its kernels finish, so one
# grow-only buffer per (device, stream) serves every call launched on that stream: kernels of one stream run in order.
# torch.empty per call costs ~25 us each on this path (blocks that were used on the side stream are recycled through
# events), which matters when a whole ChainLoss step is 0.8 ms of device time.  CHAIN_B200_CACHE_WS=0 turns it off.
_ws_cache = {}


def _workspace(dev, nbytes, tstream, side):
    """(uint8 scratch tensor, came-from-cache)."""
    nbytes = max(int(nbytes), 1)
    if os.environ.get("CHAIN_B200_CACHE_WS", "1") == "0" or torch.cuda.is_current_stream_capturing():
        return torch.empty(nbytes, dtype=torch.uint8, device=dev), False
    key = (dev.index or 0, tstream.cuda_stream)
    t = _ws_cache.get(key)
    if t is None or t.numel() < nbytes:
        if t is not None and side:
            t.record_stream(tstream)   # the block goes back to the allocator: kernels of `tstream` may still use it
        t = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _ws_cache[key] = t
    return t, True


def release_workspaces():
    """Drop the cached scratch buffers (they are as large as the largest batch seen: B*(T+1)*(S+classes)*4 bytes)."""
    for (idx, _), t in list(_ws_cache.items()):
        with torch.cuda.device(idx):
            torch.cuda.synchronize()
    _ws_cache.clear()


def _small_outputs(B, dev):
    """objf, ok, diag[8] and logp[B] carved from ONE allocation: [0] objf | [1] ok (int32) | [8:16] diag | [16:] logp."""
    small = torch.empty(16 + B, dtype=torch.float32, device=dev)
    return small, small[0], small[1:2].view(torch.int32), small[8:16], small[16:16 + B]


def _launch_stream(dev, stream):
    """cudaStream_t to launch on.  With a side `stream`, everything enqueued so far on the current
    stream (input uploads, the producer of x) is ordered before the launch."""
    cur = torch.cuda.current_stream(dev)
    if stream is None or stream == cur:
        return cur, ctypes.c_void_p(cur.cuda_stream)
    stream.wait_stream(cur)
    return stream, ctypes.c_void_p(stream.cuda_stream)


def _run_generic(log_domain, graph_tensors, x, lengths, num_states, leaky_coefficient, flags, want_logp=True,
                 stream=None, grad_scale=1.0):
    global _launches
    lib = _cabi.lib()
    what = "nnet_output" if log_domain else "exp_nnet_output"
    _require_cuda(x, what)
    B, T, D = x.shape
    names = _GRAPH_NAMES_LOG if log_domain else _GRAPH_NAMES_PROB
    K = int(graph_tensors[0].size(1))
    S = int(num_states)
    shared = _is_shared(graph_tensors, B)
    if shared:
        flags |= GRAPH_SHARED
    dev = x.device
    g = _cached_upload(list(graph_tensors), dev, shared, names)
    ldev = _lengths_dev(lengths, dev, B)
    small, objf, ok, diag, logp = _small_outputs(B, dev)
    if not want_logp:
        logp = None
    grad = torch.empty_like(x)
    ws_bytes = lib.chain_b200_workspace_bytes(B, T, D, S, K, flags)
    tstream, cstream = _launch_stream(dev, stream)
    ws, ws_cached = _workspace(dev, ws_bytes, tstream, stream is not None)
    if log_domain:
        st = lib.chain_b200_forward_backward_log_domain(
            *[_ptr(t) for t in g], _ptr(x), _ptr(ldev), B, T, D, S, K, flags, float(grad_scale),
            _ptr(objf), _ptr(logp), _ptr(grad), _ptr(ok), _ptr(diag), _ptr(ws), ws_bytes, dev.index or 0, cstream)
        _cabi.check(st, "chain_b200_forward_backward_log_domain")
    else:
        st = lib.chain_b200_forward_backward(
            *[_ptr(t) for t in g], _ptr(x), _ptr(ldev), B, T, D, S, K, float(leaky_coefficient), flags, float(grad_scale),
            _ptr(objf), _ptr(logp), _ptr(grad), _ptr(ok), _ptr(diag), _ptr(ws), ws_bytes, dev.index or 0, cstream)
        _cabi.check(st, "chain_b200_forward_backward")
    _launches += lib.chain_b200_last_launch_count()
    if stream is not None:  # every buffer of the current stream's pool that the side-stream kernels touch
        _record_side_stream(tstream, [None if ws_cached else ws, x, ldev, small, grad] + list(g))
    return Result(objf, logp, grad, ok, diag)


# ------------------------------------------------------------------ shared denominator graph
class PreparedDenGraph(object):
    """A denominator graph packed once by chain_b200_den_prepare: host blob
    (launch geometry) + its device copy."""

    def __init__(self, blob_host, blob_dev, copies, S, K, D):
        self.blob_host, self.blob_dev, self.copies, self.S, self.K, self.D = blob_host, blob_dev, copies, S, K, D


_prepared_cache = OrderedDict()


def prepare_den_graph(graph_tensors, num_states, D, device, log_domain=False, batch_size=None, content_key=None):
    """Pack (once; cached) a single graph given its reference-layout tensors (row 0 of each is
    used).  `graph_tensors` order = _GRAPH_NAMES_PROB, or _GRAPH_NAMES_LOG for a log-domain graph.
    The cache is keyed by the tensors' storage (pointer + version), or by `content_key` when given."""
    lib = _cabi.lib()
    names = _GRAPH_NAMES_LOG if log_domain else _GRAPH_NAMES_PROB
    rows = [t[0] if t.dim() == (3 if "transition" in n and "probs" not in n else 2) else t
            for t, n in zip(graph_tensors, names)]
    # the record order is optimised for the group width the kernels will use for this batch size
    dev_index = torch.device(device).index or 0
    G = lib.chain_b200_den_group_width(int(num_states), int(D), int(batch_size), dev_index) if batch_size else 4
    G = G if G in (1, 2, 4) else 4
    ident = content_key if content_key is not None else tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in rows)
    key = (ident, int(D), str(device), bool(log_domain), G)
    hit = _prepared_cache.get(key)
    if hit is not None:
        _prepared_cache.move_to_end(key)
        return hit[1]
    host = [t.detach().cpu().contiguous() for t in rows]
    if log_domain:
        ft, fi, fp, bt, bi, bp, ini, fin = host
        lk = None
    else:
        ft, fi, fp, bt, bi, bp, lk, ini, fin = host
    S, K = int(num_states), int(ft.size(0))
    pflags = (_cabi.LOG_DOMAIN if log_domain else 0) | (G << 8)
    nbytes = lib.chain_b200_den_prepare_bytes(_ptr(ft), _ptr(fi), _ptr(bt), _ptr(bi), S, K, int(D), pflags)
    if nbytes < 0:
        _cabi.check(int(nbytes), "chain_b200_den_prepare_bytes")
    blob = torch.empty(int(nbytes), dtype=torch.uint8).pin_memory() if torch.cuda.is_available() else \
        torch.empty(int(nbytes), dtype=torch.uint8)
    st = lib.chain_b200_den_prepare(_ptr(ft), _ptr(fi), _ptr(fp), _ptr(bt), _ptr(bi), _ptr(bp), _ptr(lk), _ptr(ini),
                                    _ptr(fin), S, K, int(D), pflags, _ptr(blob), int(nbytes))
    _cabi.check(st, "chain_b200_den_prepare")
    copies = int(os.environ.get("CHAIN_B200_BLOB_COPIES", "1"))
    blob_dev = blob.to(device).repeat(copies) if torch.device(device).type == "cuda" else None
    prep = PreparedDenGraph(blob, blob_dev, copies, S, K, int(D))
    _prepared_cache[key] = (rows if content_key is None else None, prep)  # pointer keys: keep the sources alive
    while len(_prepared_cache) > 8:
        _prepared_cache.popitem(last=False)
    return prep


def _run_den_shared(prep, x, lengths, leaky_coefficient, flags, stream=None, grad_scale=1.0):
    global _launches
    lib = _cabi.lib()
    _require_cuda(x, "exp_nnet_output" if flags & X_IS_EXP else "nnet_output")
    B, T, D = x.shape
    if D != prep.D:
        raise ValueError("graph was prepared for {} pdfs, input has {}".format(prep.D, D))
    dev = x.device
    ldev = _lengths_dev(lengths, dev, B)
    small, objf, ok, diag, logp = _small_outputs(B, dev)
    grad = torch.empty_like(x)
    ws_bytes = lib.chain_b200_den_workspace_bytes(_ptr(prep.blob_host), B, T)
    tstream, cstream = _launch_stream(dev, stream)
    ws, ws_cached = _workspace(dev, ws_bytes, tstream, stream is not None)
    st = lib.chain_b200_den_forward_backward(
        _ptr(prep.blob_host), _ptr(prep.blob_dev), prep.copies, _ptr(x), _ptr(ldev), B, T, D, float(leaky_coefficient), flags,
        float(grad_scale), _ptr(objf), _ptr(logp), _ptr(grad), _ptr(ok), _ptr(diag), _ptr(ws), ws_bytes, dev.index or 0, cstream)
    _cabi.check(st, "chain_b200_den_forward_backward")
    _launches += lib.chain_b200_last_launch_count()
    if stream is not None:
        _record_side_stream(tstream, [None if ws_cached else ws, x, ldev, small, grad, prep.blob_dev])
    return Result(objf, logp, grad, ok, diag)


def _shared_fits(S, D):
    """Working set of the packed shared-graph kernels at G=1 (2 state rows + 3 frame rows, padded
    to 4) against the 227 KB of shared memory; larger graphs take the general kernel."""
    pad = lambda n: (n + 3) // 4 * 4  # noqa: E731
    return 4 * (2 * pad(S) + 3 * pad(D)) + 4096 <= 227 * 1024 and S <= 65535 and D <= 65535


# ------------------------------------------------------------------ public helpers
DEN_PATH = "auto"  # "auto": shared graphs use the packed shared-graph kernels; "generic": always the per-sequence kernel

_FALLBACK_STATUS = ("status -1)", "status -3)")  # ERR_INVALID_ARG / ERR_UNSUPPORTED from prepare or the packed launch


def _same_graph_everywhere(tensors, B):
    """A MATERIALISED batch (`[B,...]` contiguous, what the reference's `ChainGraphBatch.initialized_by_one`
    builds with `.repeat`, graph.py:101-120) whose rows all hold the same graph?  Checked on rows
    0, B//2 and B-1 of every tensor (all rows with CHAIN_B200_STRICT_SHARED_CHECK=1): such a batch is
    routed to the packed shared-graph kernels instead of the 8x slower per-sequence kernel."""
    if B < 2 or any(t is None or t.dim() < 2 or t.size(0) != B for t in tensors):
        return False
    if os.environ.get("CHAIN_B200_STRICT_SHARED_CHECK"):
        return all(bool((t == t[0:1]).all()) for t in tensors)
    return all(torch.equal(t[0], t[B // 2]) and torch.equal(t[0], t[B - 1]) for t in tensors)


def _content_key(rows):
    """Fingerprint of a graph's tensors (row 0 of a batch): materialised batches are new tensors every step
    (the reference re-creates them per forward, loss.py:99), so the packed-graph cache cannot key on pointers."""
    import hashlib
    h = hashlib.blake2b(digest_size=16)
    for t in rows:
        h.update(t.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def _shared_rows(graph_tensors, B):
    """Row-0 tensors when the batch is ONE graph (stride-0 views, a batch of one, or a materialised batch of
    identical rows), else None.  The second value is a content key for materialised batches."""
    if _is_shared(graph_tensors, B):
        return graph_tensors, None
    if _same_graph_everywhere(graph_tensors, B):
        rows = tuple(t[0:1] for t in graph_tensors)
        return rows, _content_key(rows)
    return None, None


def _try_packed(run, fallback):
    """Run the packed shared-graph path; graphs it cannot take (padded states in a batch of one,
    on-chip working set too large once rows are split) go to the general kernel."""
    try:
        return run()
    except ChainB200Error as e:
        if str(e).endswith(_FALLBACK_STATUS):
            return fallback()
        raise


def forward_backward_prob(graph_tensors, x, lengths, num_states, leaky_coefficient=1e-5, x_is_exp=True, clamp=False,
                          path=None, stream=None, grad_scale=1.0):
    """Probability-domain (denominator) forward-backward.  `graph_tensors` in
    _GRAPH_NAMES_PROB order, reference layout.  The posteriors are written multiplied by `grad_scale`."""
    flags = (X_IS_EXP if x_is_exp else 0) | (CLAMP_X if clamp else 0)
    path = path or DEN_PATH
    B = x.size(0)
    generic = lambda: _run_generic(False, graph_tensors, x, lengths, num_states, leaky_coefficient, flags,  # noqa: E731
                                   stream=stream, grad_scale=grad_scale)
    if path != "generic" and _shared_fits(int(num_states), int(x.size(2))):
        rows, ckey = _shared_rows(graph_tensors, B)
        if rows is not None:
            def packed():
                prep = prepare_den_graph(rows, num_states, x.size(2), x.device, batch_size=B, content_key=ckey)
                return _run_den_shared(prep, x, lengths, leaky_coefficient, flags, stream=stream, grad_scale=grad_scale)
            return _try_packed(packed, generic)
    return generic()


def forward_backward_log(graph_tensors, x, lengths, num_states, grad_as_log=True, clamp=False, path=None, stream=None,
                         grad_scale=1.0):
    """Log-domain (numerator) forward-backward.  `graph_tensors` in _GRAPH_NAMES_LOG order.
    A graph shared by the whole batch (stride-0 views) with >= 256 arcs runs on the packed
    shared-graph kernels, per-utterance graphs on the general kernel."""
    flags = (GRAD_AS_LOG if grad_as_log else 0) | (CLAMP_X if clamp else 0)
    path = path or DEN_PATH
    S, K = int(num_states), int(graph_tensors[0].size(1))
    B = x.size(0)
    generic = lambda: _run_generic(True, graph_tensors, x, lengths, num_states, 0.0, flags, stream=stream,  # noqa: E731
                                   grad_scale=grad_scale)
    if (path != "generic" and B > 1 and K >= 256 and S <= 32767 and x.size(2) <= 32767
            and _shared_fits(S, int(x.size(2)))):
        rows, ckey = _shared_rows(graph_tensors, B)
        if rows is not None:
            def packed():
                prep = prepare_den_graph(rows, num_states, x.size(2), x.device, log_domain=True, batch_size=B,
                                         content_key=ckey)
                return _run_den_shared(prep, x, lengths, 0.5, flags, stream=stream, grad_scale=grad_scale)
            return _try_packed(packed, generic)
    return generic()


def combine_loss(num, den, loss_scale, with_grad=True):
    """The tail of `ChainLoss.forward` (loss.py:102-104) in one launch: loss = (den.objf - num.objf) * loss_scale and
    den.grad += num.grad (both already scaled by the kernels).  Returns the 0-dim loss tensor."""
    global _launches
    lib = _cabi.lib()
    loss = torch.empty((), dtype=torch.float32, device=den.objf.device)
    cur = torch.cuda.current_stream(loss.device)
    st = lib.chain_b200_combine_loss(_ptr(num.objf), _ptr(den.objf), float(loss_scale), _ptr(den.grad),
                                     _ptr(num.grad) if with_grad else None, den.grad.numel(), _ptr(loss),
                                     loss.device.index or 0, ctypes.c_void_p(cur.cuda_stream))
    _cabi.check(st, "chain_b200_combine_loss")
    _launches += lib.chain_b200_last_launch_count()
    return loss


def scale_grad_(grad, upstream):
    """grad *= upstream (a 0-dim device tensor) in place; a no-op on the device when upstream == 1."""
    global _launches
    lib = _cabi.lib()
    if not upstream.is_cuda or upstream.dtype != torch.float32:
        upstream = upstream.to(device=grad.device, dtype=torch.float32)
    cur = torch.cuda.current_stream(grad.device)
    st = lib.chain_b200_scale_grad(_ptr(grad), grad.numel(), _ptr(upstream), grad.device.index or 0,
                                   ctypes.c_void_p(cur.cuda_stream))
    _cabi.check(st, "chain_b200_scale_grad")
    _launches += lib.chain_b200_last_launch_count()
    return grad
