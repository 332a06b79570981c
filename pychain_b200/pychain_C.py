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
  * `ok` is returned as a CPU bool[1] tensor like the reference (this synchronises;
    `pychain_b200.loss` uses the engine directly and does not).
"""
import torch

from . import _cabi, engine


def _report(r, batch, log_domain):
    """What the reference's self-check prints to stderr (chain-computation.cc:363-390,
    chain-log-domain-computation.cc:289-303), from the device-side diagnostics: on failure, and
    always at verbose level >= 1 (the reference re-checks every frame then; here frame 0 is the
    one that accumulates the whole backward recursion)."""
    import sys
    ok = bool(r.ok.item())
    verbose = _cabi.lib().chain_b200_get_verbose_level() >= 1
    if ok and not verbose:
        return ok
    ab, gs = (float(v) for v in r.diag.tolist())
    if not log_domain and (verbose or abs(ab - batch) >= ab * 0.01):
        print("On time 0, alpha-beta product {} != {}".format(ab, batch), file=sys.stderr)
    if verbose or abs(gs - batch) >= gs * 0.01:
        print("On time 0, log-prob-deriv sum {} != {}".format(gs, batch), file=sys.stderr)
    if not ok:
        print("Excessive error detected, will abandon this minibatch", file=sys.stderr)
    return ok


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
    ok = _report(r, exp_nnet_output.size(0), False)
    return [r.objf, r.grad, torch.full((1,), ok, dtype=torch.bool)]


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
    ok = _report(r, nnet_output.size(0), True)
    return [r.objf, r.grad, torch.full((1,), ok, dtype=torch.bool)]


def set_verbose_level(level):
    _cabi.lib().chain_b200_set_verbose_level(int(level))
