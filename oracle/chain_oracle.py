"""Python face of the oracle.  TEST INFRASTRUCTURE ONLY (see chain_oracle.c header).

Three independent checkers for the LF-MMI forward-backward hot path:

  * `den_c` / `num_c`        -- the fp32 C restatement (chain_oracle.c) via ctypes;
  * `den_f64` / `num_f64`    -- a float64 numpy restatement of SURVEY.md Appendix A
                                (small cases; independent of the C code's loop order);
  * `den_ref` / `num_ref`    -- the UNMODIFIED reference extension built by
                                oracle/build_ref.py (CPU or CUDA path, chosen by the
                                device of `x`), called exactly the way
                                /root/reference/pychain/loss.py:44-76 calls it.

Parity status: pinned -- `tests/test_oracle.py` checks den_c/num_c/den_f64/num_f64
against `den_ref`/`num_ref` outputs stored in tests/golden/ (and live against the
reference .so when it is present).
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(HERE, "chain_oracle.c")
_LIB = os.path.join(HERE, "_build", "libchain_oracle.so")
_lib = None


def build_c(force=False):
    if force or not os.path.isfile(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", _LIB, _SRC, "-lm"])
    return _LIB


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_c())
        _lib.oracle_den_forward_backward.restype = ctypes.c_int
        _lib.oracle_num_forward_backward_log_domain.restype = ctypes.c_int
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _graph_args(graphs, with_leaky):
    """Row-0 pointers + `shared` flag when the batch dim has stride 0, else the
    contiguous batch tensors."""
    names = ["forward_transitions", "forward_transition_indices", "forward_transition_probs",
             "backward_transitions", "backward_transition_indices", "backward_transition_probs"]
    if with_leaky:
        names.append("leaky_probs")
    names += ["initial_probs", "final_probs"]
    ts = [getattr(graphs, n) for n in names]
    shared = all(t.stride(0) == 0 for t in ts)
    keep = [t[0].contiguous() if shared else t.contiguous() for t in ts]
    for t in keep:
        assert t.device.type == "cpu"
    return keep, int(shared)


def den_c(graphs, exp_x, lengths, leaky_coefficient=1e-5):
    """fp32 C restatement of `pychain_C.forward_backward` (CPU path).
    Returns (objf float, logp [B], grad [B,T,D], ok bool, diag [2])."""
    lib = _load()
    exp_x = exp_x.detach().cpu().contiguous().float()
    lengths = lengths.detach().cpu().contiguous().long()
    B, T, D = exp_x.shape
    S = graphs.num_states
    K = graphs.forward_transitions.size(1)
    keep, shared = _graph_args(graphs, True)
    objf = ctypes.c_float()
    ok = ctypes.c_int()
    logp = torch.empty(B)
    grad = torch.empty(B, T, D)
    diag = torch.zeros(2)
    rc = lib.oracle_den_forward_backward(
        *[_p(t) for t in keep], ctypes.c_int(shared), _p(exp_x), _p(lengths),
        B, T, D, S, K, ctypes.c_float(leaky_coefficient),
        ctypes.byref(objf), _p(logp), _p(grad), ctypes.byref(ok), _p(diag))
    assert rc == 0
    return objf.value, logp, grad, bool(ok.value), diag


def num_c(graphs, x, lengths):
    """fp32 C restatement of `pychain_C.forward_backward_log_domain` (CPU path).
    Returns (objf, logp [B], log_grad [B,T,D], ok, diag [1])."""
    lib = _load()
    x = x.detach().cpu().contiguous().float()
    lengths = lengths.detach().cpu().contiguous().long()
    B, T, D = x.shape
    S = graphs.num_states
    K = graphs.forward_transitions.size(1)
    keep, shared = _graph_args(graphs, False)
    objf = ctypes.c_float()
    ok = ctypes.c_int()
    logp = torch.empty(B)
    grad = torch.empty(B, T, D)
    diag = torch.zeros(1)
    rc = lib.oracle_num_forward_backward_log_domain(
        *[_p(t) for t in keep], ctypes.c_int(shared), _p(x), _p(lengths),
        B, T, D, S, K, ctypes.byref(objf), _p(logp), _p(grad), ctypes.byref(ok), _p(diag))
    assert rc == 0
    return objf.value, logp, grad, bool(ok.value), diag


# ----------------------------------------------------------------------------- float64
def _row(graphs, name, b):
    t = getattr(graphs, name)
    return t[b].detach().cpu().numpy()


def den_f64(graphs, exp_x, lengths, leaky_coefficient=1e-5):
    """SURVEY.md Appendix A, denominator, in float64 (vectorised over arcs)."""
    exp_x = exp_x.detach().cpu().double().numpy()
    B, T, D = exp_x.shape
    S = graphs.num_states
    c = float(leaky_coefficient)
    grad = np.zeros((B, T, D))
    logp = np.zeros(B)
    for b in range(B):
        L = int(lengths[b])
        tr = _row(graphs, "forward_transitions", b).astype(np.int64)
        idx = _row(graphs, "forward_transition_indices", b)
        nk = int(idx[:, 1].max()) if idx.size else 0
        src, dst, pdf = tr[:nk, 0], tr[:nk, 1], tr[:nk, 2]
        p = _row(graphs, "forward_transition_probs", b)[:nk].astype(np.float64)
        lk = _row(graphs, "leaky_probs", b).astype(np.float64)
        ini = _row(graphs, "initial_probs", b).astype(np.float64)
        fin = _row(graphs, "final_probs", b).astype(np.float64)
        alpha = np.zeros((L + 1, S))
        tot = np.zeros(L + 1)
        a = ini.copy()
        tot[0] = a.sum()
        alpha[0] = a + c * tot[0] * lk
        for t in range(1, L + 1):
            a = np.bincount(dst, weights=alpha[t - 1][src] * p * exp_x[b, t - 1][pdf], minlength=S) / tot[t - 1]
            tot[t] = a.sum()
            alpha[t] = a + c * tot[t] * lk
        Z = (alpha[L] * fin).sum()
        logp[b] = np.log(tot[:L]).sum() + np.log(Z)
        bd = fin / Z
        beta = bd + c * (bd * lk).sum()
        for t in range(L - 1, -1, -1):
            v = p * beta[dst] * exp_x[b, t][pdf]
            bd = np.bincount(src, weights=v, minlength=S) / tot[t]
            grad[b, t] = np.bincount(pdf, weights=v * alpha[t][src] / tot[t], minlength=D)
            beta = bd + c * (bd * lk).sum()
    return float(logp.sum()), torch.from_numpy(logp), torch.from_numpy(grad)


def _lse(v):
    m = np.max(v) if v.size else -np.inf
    if not np.isfinite(m):
        return m
    return m + np.log(np.exp(v - m).sum())


def num_f64(graphs, x, lengths):
    """SURVEY.md Appendix A, numerator (log domain), in float64.  Returns
    (objf, logp [B], log_grad [B,T,D])."""
    x = x.detach().cpu().double().numpy()
    B, T, D = x.shape
    S = graphs.num_states
    lg = np.full((B, T, D), -np.inf)
    logp = np.zeros(B)
    with np.errstate(divide="ignore", invalid="ignore"):
        for b in range(B):
            L = int(lengths[b])
            tr = _row(graphs, "forward_transitions", b).astype(np.int64)
            idx = _row(graphs, "forward_transition_indices", b)
            nk = int(idx[:, 1].max()) if idx.size else 0
            src, dst, pdf = tr[:nk, 0], tr[:nk, 1], tr[:nk, 2]
            p = _row(graphs, "forward_transition_probs", b)[:nk].astype(np.float64)
            ini = _row(graphs, "initial_probs", b).astype(np.float64)
            fin = _row(graphs, "final_probs", b).astype(np.float64)
            alpha = np.full((L + 1, S), -np.inf)
            sc = np.zeros(L + 1)
            alpha[0] = ini
            for t in range(1, L + 1):
                term = alpha[t - 1][src] + p + x[b, t - 1][pdf]
                a = np.full(S, -np.inf)
                for j in np.unique(dst):
                    a[j] = _lse(term[dst == j])
                alpha[t] = a - sc[t - 1]
                sc[t] = _lse(alpha[t])
            lZ = _lse(alpha[L] + fin)
            scl = sc[:L]
            logp[b] = np.where(np.isfinite(scl), scl, 0.0).sum() + lZ
            beta = fin - lZ
            for t in range(L - 1, -1, -1):
                v = p + beta[dst] + x[b, t][pdf] - sc[t]
                nb = np.full(S, -np.inf)
                for i in np.unique(src):
                    nb[i] = _lse(v[src == i])
                occ = v + alpha[t][src]
                for n in np.unique(pdf):
                    lg[b, t, n] = _lse(occ[pdf == n])
                beta = nb
    return float(logp.sum()), torch.from_numpy(logp), torch.from_numpy(lg)


# --------------------------------------------------------------------------- reference
_ref_mod = None


def ref_available():
    from . import build_ref
    return build_ref.is_built()


def ref_module():
    global _ref_mod
    if _ref_mod is None:
        from . import build_ref
        _ref_mod = build_ref.load()
    return _ref_mod


def _batch_sizes(lengths, T):
    """What `pack_padded_sequence(...).batch_sizes` yields (loss.py:37-40)."""
    lengths = lengths.cpu().long()
    return torch.tensor([int((lengths > t).sum()) for t in range(int(lengths.max()))], dtype=torch.long)


def _contig_graph(graphs, names):
    return [getattr(graphs, n).contiguous() for n in names]


def den_ref(graphs, exp_x, lengths, leaky_coefficient=1e-5):
    """The reference `pychain_C.forward_backward` (pychain.cc:26-79), unmodified.
    CPU or CUDA path according to `exp_x.device`.  Returns (objf tensor, grad, ok)."""
    m = ref_module()
    names = ["forward_transitions", "forward_transition_indices", "forward_transition_probs",
             "backward_transitions", "backward_transition_indices", "backward_transition_probs",
             "leaky_probs", "initial_probs", "final_probs", "start_state"]
    lengths = lengths.cpu().long()
    bs = _batch_sizes(lengths, exp_x.size(1))
    objf, grad, ok = m.forward_backward(*_contig_graph(graphs, names), exp_x.contiguous(), bs, lengths,
                                        graphs.num_states, leaky_coefficient)
    return objf, grad, bool(ok.item())


def num_ref(graphs, x, lengths):
    """The reference `pychain_C.forward_backward_log_domain` (pychain.cc:81-129)."""
    m = ref_module()
    names = ["forward_transitions", "forward_transition_indices", "forward_transition_probs",
             "backward_transitions", "backward_transition_indices", "backward_transition_probs",
             "initial_probs", "final_probs", "start_state"]
    lengths = lengths.cpu().long()
    bs = _batch_sizes(lengths, x.size(1))
    objf, lgrad, ok = m.forward_backward_log_domain(*_contig_graph(graphs, names), x.contiguous(), bs, lengths,
                                                    graphs.num_states)
    return objf, lgrad, bool(ok.item())
