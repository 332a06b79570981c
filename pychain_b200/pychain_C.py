"""Drop-in for the reference's native module `pychain_C`
(/root/reference/pytorch_binding/src/pychain.cc:131-135): the same three
callables with the same positional signatures and return values, implemented by
the sm_100a engine behind include/chain_b200.h.

    forward_backward(forward_transitions, forward_transition_indices, forward_transition_probs,
                     backward_transitions, backward_transition_indices, backward_transition_probs,
                     leaky_probs, initial_probs, final_probs, start_state,
                     exp_nnet_output, batch_sizes, sequence_lengths, num_states,
                     leaky_hmm_coefficient=1e-5) -> [objf, nnet_output_grad, ok]         (pychain.cc:26-79)
    forward_backward_log_domain(... no leaky_probs ..., nnet_output, batch_sizes,
                     sequence_lengths, num_states) -> [objf, log_grad, ok]               (pychain.cc:81-129)
    set_verbose_level(level)                                                           (pychain.cc:134)

Differences from the reference, all permitted by its contract:
  * `exp_nnet_output` / `nnet_output` must be a CUDA tensor (the reference also has
    a CPU path; this engine deliberately has none -- a CPU input raises).
  * graph tensors may be CPU (uploaded once and cached, instead of on every call)
    or CUDA, contiguous `[B,...]` or stride-0 shared views.
  * `start_state` and `batch_sizes` are accepted and ignored; the reference never
    reads `start_state` (SURVEY.md Appx C-12) and uses `batch_sizes` only to
    express "sequence b is alive at frame t", which `sequence_lengths` already says.
  * `ok` is returned as a CPU bool[1] tensor like the reference.  The reference fills it after
    `.item()` syncs; here it is a pinned host tensor filled by an ASYNCHRONOUS device-to-host copy on the
    current stream, so the call itself does not synchronise (the reference's only caller ignores `ok`,
    loss.py:44,62).  Its value is valid after any later synchronisation of that stream (`objf.item()`,
    `torch.cuda.synchronize()`); at verbose level >= 1 the call synchronises and reports like the reference.
"""
import torch

from . import _cabi, engine


_ok_ring = []      # pinned bool[1] buffers, reused round-robin (cudaHostAlloc per call would serialise)
_ok_next = 0
_ok_pending = []   # (event, buffer) of calls whose verdict has not been looked at yet


def poll_reports(wait=False):
    """Print the reference's failure line for every earlier call whose asynchronous self-check has completed
    and failed.  Called at the start of every forward_backward*; call it with wait=True (after a
    synchronisation point of your own) to flush."""
    import sys
    keep = []
    for ev, buf in _ok_pending:
        if wait:
            ev.synchronize()
        if ev.query():
            if not bool(buf[0]):
                print("Excessive error detected, will abandon this minibatch", file=sys.stderr)
        else:
            keep.append((ev, buf))
    _ok_pending[:] = keep


def _ok_host(ok_dev):
    """bool[1] CPU tensor that receives `ok` by an async copy on the current stream."""
    global _ok_next
    if not _ok_ring:
        _ok_ring.extend(torch.ones(1, dtype=torch.bool).pin_memory() for _ in range(64))
    if len(_ok_pending) >= len(_ok_ring) - 1:   # never hand out a buffer that is still pending
        poll_reports(wait=True)
    buf = _ok_ring[_ok_next % len(_ok_ring)]
    _ok_next += 1
    buf.copy_(ok_dev != 0, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(ok_dev.device))
    _ok_pending.append((ev, buf))
    return buf


def _per_frame_check(r, lengths, log_domain):
    """The reference re-runs its self-check on EVERY frame at verbose level >= 1 (chain-computation.cc:337,
    chain-log-domain-computation.cc:276-279): sum_n gamma(t,n) over the sequences alive at t must equal their number.
    Recomputed here from the gradient the kernels wrote (a debugging aid: it synchronises, like the reference's).
    Returns False when a frame is off by more than 5 %."""
    import sys
    g = r.grad.exp() if log_domain else r.grad
    rows = g.sum(-1)                                              # [B,T]
    T = rows.size(1)
    alive = (torch.arange(T, device=rows.device)[None, :] < lengths.to(rows.device)[:, None])
    sums = (rows * alive).sum(0).tolist()
    counts = alive.sum(0).tolist()
    ok = True
    for t in range(T - 1, 0, -1):                                 # frame 0 is reported by the caller
        s, n = sums[t], float(counts[t])
        if n and not abs(s - n) < s * 0.01:                       # ApproxEqual(s, n, 0.01), base.cc:6-10
            print("On time {}, log-prob-deriv sum {} != {}".format(t, s, n), file=sys.stderr)
            if abs(s - n) > 0.05 * n or s != s:
                print("Excessive error detected, will abandon this minibatch", file=sys.stderr)
                ok = False
    return ok


def _report(r, batch, log_domain, lengths=None):
    """`ok` for the caller + what the reference's self-check prints to stderr
    (chain-computation.cc:363-390, chain-log-domain-computation.cc:289-303) from the device-side
    diagnostics.  Quiet level: no synchronisation.  Verbose level >= 1: synchronise, report always
    (the reference re-checks every frame then; here frame 0 accumulates the whole backward recursion)."""
    import sys
    verbose = _cabi.lib().chain_b200_get_verbose_level() >= 1
    poll_reports()
    if not verbose:
        return _ok_host(r.ok)
    ok = bool(r.ok.item())
    if lengths is not None:
        ok = _per_frame_check(r, lengths, log_domain) and ok
    ab, gs = (float(v) for v in r.diag.tolist())
    if not log_domain:
        print("On time 0, alpha-beta product {} != {}".format(ab, batch), file=sys.stderr)
    print("On time 0, log-prob-deriv sum {} != {}".format(gs, batch), file=sys.stderr)
    if not ok:
        print("Excessive error detected, will abandon this minibatch", file=sys.stderr)
    return torch.full((1,), ok, dtype=torch.bool)


def _check_contiguous(t, name):
    if t is not None and not (t.is_contiguous() or t.stride(0) == 0):
        raise RuntimeError("{} must be contiguous".format(name))


def forward_backward(forward_transitions, forward_transition_indices, forward_transition_probs,
                     backward_transitions, backward_transition_indices, backward_transition_probs,
                     leaky_probs, initial_probs, final_probs, start_state,
                     exp_nnet_output, batch_sizes, sequence_lengths, num_states,
                     leaky_hmm_coefficient=1.0e-05):
    g = (forward_transitions, forward_transition_indices, forward_transition_probs,
         backward_transitions, backward_transition_indices, backward_transition_probs,
         leaky_probs, initial_probs, final_probs)
    for t, n in zip(g, engine._GRAPH_NAMES_PROB):
        _check_contiguous(t, n)
    r = engine.forward_backward_prob(g, exp_nnet_output, sequence_lengths, num_states, leaky_hmm_coefficient,
                                     x_is_exp=True, clamp=False)
    return [r.objf, r.grad, _report(r, exp_nnet_output.size(0), False, sequence_lengths)]


def forward_backward_log_domain(forward_transitions, forward_transition_indices, forward_transition_probs,
                                backward_transitions, backward_transition_indices, backward_transition_probs,
                                initial_probs, final_probs, start_state,
                                nnet_output, batch_sizes, sequence_lengths, num_states):
    g = (forward_transitions, forward_transition_indices, forward_transition_probs,
         backward_transitions, backward_transition_indices, backward_transition_probs,
         initial_probs, final_probs)
    for t, n in zip(g, engine._GRAPH_NAMES_LOG):
        _check_contiguous(t, n)
    r = engine.forward_backward_log(g, nnet_output, sequence_lengths, num_states, grad_as_log=True, clamp=False)
    return [r.objf, r.grad, _report(r, nnet_output.size(0), True, sequence_lengths)]


def set_verbose_level(level):
    _cabi.lib().chain_b200_set_verbose_level(int(level))
