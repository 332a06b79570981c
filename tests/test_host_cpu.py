"""CPU-side checks: the C-ABI library loads and exports every symbol of
include/chain_b200.h, the graph containers mirror the reference's layout, the
denominator packer is a loss-less re-arrangement of the arcs, argument checking
and error behaviour.  No kernel is launched here."""
import ctypes
import os
import re
import struct

import numpy as np
import pytest
import torch

import pychain_b200
from pychain_b200 import ChainGraph, ChainGraphBatch, ChainFunction, Fst, synth, engine, _cabi
from pychain_b200 import pychain_C as new_C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "chain_b200.h")).read()
    declared = set(re.findall(r"\b(chain_b200_[a-z_]+)\s*\(", hdr))
    assert declared == set(_cabi.SIGNATURES), declared ^ set(_cabi.SIGNATURES)
    lib = _cabi.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.chain_b200_version() >= 100
    assert b"workspace" in lib.chain_b200_status_string(-2)
    lib.chain_b200_set_verbose_level(2)
    assert lib.chain_b200_get_verbose_level() == 2
    lib.chain_b200_set_verbose_level(0)
    assert lib.chain_b200_workspace_bytes(4, 10, 8, 5, 9, 0) > 4 * 11 * 5 * 4
    assert lib.chain_b200_workspace_bytes(0, 10, 8, 5, 9, 0) == 0


def test_pychain_surface_matches_reference_names():
    for n in ("ChainLoss", "ChainFunction", "ChainGraph", "ChainGraphBatch"):
        assert hasattr(pychain_b200, n)
    for n in ("forward_backward", "forward_backward_log_domain", "set_verbose_level"):
        assert callable(getattr(new_C, n))
    import inspect
    sig = list(inspect.signature(new_C.forward_backward).parameters)
    assert sig == ["forward_transitions", "forward_transition_indices", "forward_transition_probs",
                   "backward_transitions", "backward_transition_indices", "backward_transition_probs",
                   "leaky_probs", "initial_probs", "final_probs", "start_state", "exp_nnet_output",
                   "batch_sizes", "sequence_lengths", "num_states", "leaky_hmm_coefficient"]  # pychain.cc:26-41
    assert len(inspect.signature(new_C.forward_backward_log_domain).parameters) == 13  # pychain.cc:81-94


def test_no_cpu_fallback():
    g = ChainGraph(synth.random_den_fst(10, 20, 16, seed=1))
    gb = ChainGraphBatch(g, 2)
    x, lengths = synth.make_inputs(2, 5, 16, seed=2)
    with pytest.raises(_cabi.ChainB200Error, match="no CPU fallback"):
        ChainFunction.apply(x, lengths, gb, 1e-5)
    with pytest.raises(_cabi.ChainB200Error):
        new_C.forward_backward(gb.forward_transitions, gb.forward_transition_indices, gb.forward_transition_probs,
                               gb.backward_transitions, gb.backward_transition_indices, gb.backward_transition_probs,
                               gb.leaky_probs, gb.initial_probs, gb.final_probs, gb.start_state,
                               x.exp(), None, lengths, gb.num_states, 1e-5)


def test_chain_function_error_behaviour():
    g = ChainGraph(synth.random_den_fst(10, 20, 16, seed=1))
    x, _ = synth.make_inputs(3, 5, 16, seed=2)
    with pytest.raises(ValueError, match="does not equal to graph batch size"):  # loss.py:32-36
        ChainFunction.apply(x, torch.tensor([5, 5, 5]), ChainGraphBatch(g, 2), 1e-5)
    with pytest.raises(RuntimeError, match="sorted in decreasing order"):        # pack_padded_sequence, loss.py:37-39
        ChainFunction.apply(x, torch.tensor([3, 5, 4]), ChainGraphBatch(g, 3), 1e-5)
    with pytest.raises(RuntimeError, match="greater than 0"):
        ChainFunction.apply(x, torch.tensor([5, 3, 0]), ChainGraphBatch(g, 3), 1e-5)
    with pytest.raises(ValueError, match="batch size should be specified"):      # graph.py:76-77
        ChainGraphBatch(g)
    with pytest.raises(ValueError, match="max_num_transitions"):                 # graph.py:81-85
        ChainGraphBatch([g, g])


def test_fst_to_tensors_layout():
    """fstext.cc:19-117 semantics: forward CSR by source, backward CSR by destination (stable)."""
    #        src dst pdf  logw
    arcs = [(0, 1, 3, -0.5), (0, 2, 1, -1.0), (1, 2, 0, -0.2), (2, 0, 2, -0.7), (1, 1, 4, -0.9)]
    fst = Fst(3, 0, [a[0] for a in arcs], [a[1] for a in arcs], [a[2] for a in arcs], [a[3] for a in arcs],
              [-np.inf, -np.inf, -0.1])
    g = ChainGraph(fst)
    assert g.forward_transitions.tolist() == [[0, 1, 3], [0, 2, 1], [1, 2, 0], [1, 1, 4], [2, 0, 2]]
    assert g.forward_transition_indices.tolist() == [[0, 2], [2, 4], [4, 5]]
    assert g.backward_transitions.tolist() == [[2, 0, 2], [0, 1, 3], [1, 1, 4], [0, 2, 1], [1, 2, 0]]
    assert g.backward_transition_indices.tolist() == [[0, 1], [1, 3], [3, 5]]
    assert torch.allclose(g.forward_transition_probs, torch.tensor([-0.5, -1.0, -0.2, -0.9, -0.7]).exp())
    assert g.forward_transitions.dtype == torch.int32 and g.forward_transition_probs.dtype == torch.float32
    assert torch.allclose(g.final_probs, torch.tensor([0.0, 0.0, np.exp(-0.1)], dtype=torch.float32))
    assert g.initial_probs.tolist() == [1.0, 0.0, 0.0] and abs(g.leaky_probs.sum().item() - 1.0) < 1e-5
    gl = ChainGraph(fst, log_domain=True, final_mode="ones")
    assert gl.leaky_probs is None and gl.initial_probs[0] == 0 and torch.isneginf(gl.initial_probs[1:]).all()
    assert gl.final_probs.tolist() == [0.0, 0.0, 0.0]
    # text format round trip (fstprint): weights are costs, ilabel = pdf + 1
    txt = ["0 1 4 4 0.5", "0 2 2 2 1.0", "1 2 1 1 0.2", "2 0 3 3 0.7", "1 1 5 5 0.9", "2 0.1"]
    g2 = ChainGraph(Fst.from_text(txt))
    assert torch.equal(g2.forward_transitions, g.forward_transitions)
    assert torch.allclose(g2.forward_transition_probs, g.forward_transition_probs)


def test_leaky_probs_power_method():
    """fstext.cc:120-171: 100 iterations of the per-state-normalised HMM, averaged."""
    fst = synth.random_den_fst(12, 40, 5, seed=3)
    S = 12
    w = np.exp(fst.log_weight.astype(np.float64))
    nf = 1.0 / (np.exp(fst.final_log_weight.astype(np.float64)) + np.bincount(fst.src, weights=w, minlength=S))
    cur = np.zeros(S); cur[0] = 1.0; avg = np.zeros(S)
    for _ in range(100):
        avg += cur / 100
        nxt = np.zeros(S)
        for s, d, ww in zip(fst.src, fst.dst, w):
            nxt[d] += cur[s] * nf[s] * ww
        cur = nxt / nxt.sum()
    assert np.abs(fst.leaky_probs().numpy() - avg.astype(np.float32)).max() < 1e-7


def test_graph_batch_layouts_and_reorder():
    g = ChainGraph(synth.random_den_fst(10, 20, 16, seed=1))
    gb = ChainGraphBatch(g, 5)
    assert gb.forward_transitions.shape == (5, 20, 3) and gb.forward_transitions.stride(0) == 0
    assert gb.leaky_probs.shape == (5, 10) and gb.start_state.tolist() == [0] * 5
    gb.reorder(torch.tensor([4, 3, 2, 1, 0]))
    assert gb.forward_transitions.stride(0) == 0
    gm = ChainGraphBatch(g, 3).materialize()
    assert gm.forward_transitions.is_contiguous() and gm.shared_graph is None
    nb = synth.make_num_batch(4, 40, 16, seed=5, min_phones=3, max_phones=6)
    assert nb.log_domain and nb.leaky_probs is None
    assert torch.isneginf(nb.initial_probs).any() and nb.forward_transitions.shape[0] == 4
    first = nb.forward_transition_probs[0].clone()
    nb.reorder(torch.tensor([1, 0, 2, 3]))
    assert torch.equal(nb.forward_transition_probs[1], first)


def _parse_blob(blob):
    h = blob.numpy().tobytes()
    (magic, version, S, K, D, nw, S_pad, D_pad, max_sub, _, opt_g, C, C_pad, R_pad, embed, _r) = struct.unpack_from(
        "<II14i", h, 0)
    views, off = [], 64
    for i in range(5):
        n_rows, n_slices, n_split, n_sub, n_recs, o_info, o_rows, o_recs, o_split = struct.unpack_from("<4i5q", h, off)
        off += 56
        info = np.frombuffer(h, np.int64, nw * 3, o_info).reshape(nw, 3)
        # embedded format: row codes travel inside the record streams; compact format: a parallel stream
        rows = np.zeros(0, np.int32) if embed else np.frombuffer(h, np.int32, (n_slices + 3 * nw) * 32, o_rows)
        recs = np.frombuffer(h, np.uint32, n_recs * 2, o_recs).reshape(-1, 2)
        split = np.frombuffer(h, np.int32, n_split * 3, o_split).reshape(-1, 3)
        views.append(dict(n_rows=n_rows, n_slices=n_slices, n_sub=n_sub, info=info, rows=rows, recs=recs, split=split,
                          embed=embed))
    o_leaky, o_init, o_final, o_cd, o_cp, total = struct.unpack_from("<6q", h, off)
    pred = struct.unpack_from("<10d", h, off + 48)
    return dict(S=S, K=K, D=D, nw=nw, max_sub=max_sub, views=views, total=total, pred=pred, C=C, C_pad=C_pad,
                S_pad=S_pad, R_pad=R_pad, version=version, embed=embed,
                cls_dst=np.frombuffer(h, np.int32, C, o_cd), cls_pdf=np.frombuffer(h, np.int32, C, o_cp),
                leaky=np.frombuffer(h, np.float32, S, o_leaky), final=np.frombuffer(h, np.float32, S, o_final),
                init=np.frombuffer(h, np.float32, S, o_init))


def _decode_view(v, nw):
    """Walk the per-warp streams exactly like sell_pass does: returns {(row, a, b, w) multiset}.  Slots come in
    pairs, [pair][lane][2]; the sign bit of the second slot's weight word ends a slice (warp-uniform).  Embedded
    format: that slot is the lane's row code; compact format: it is still an arc and the row codes come from the
    parallel per-warp stream.  Weight 0 = idle slot."""
    out = []
    sub_owner = {}
    for row, first, cnt in v["split"]:
        for j in range(cnt):
            sub_owner[first + j] = row
    embed = v["embed"]
    slices_seen = 0
    for w in range(nw):
        rec0, npairs, third = v["info"][w]
        rp = third
        pending = [[] for _ in range(32)]
        seen_w = 0
        for j in range(npairs):
            blk = v["recs"][rec0 + j * 64: rec0 + (j + 1) * 64].reshape(32, 2, 2)   # [lane][slot][word]
            flags = (blk[:, 1, 1] >> 31) & 1
            assert flags.min() == flags.max(), "end-of-slice flag must be warp-uniform"
            for lane in range(32):
                for slot in range(2):
                    if slot == 1 and flags[0] and embed:
                        continue
                    ab, wb = int(blk[lane, slot, 0]), int(blk[lane, slot, 1])
                    assert not (wb >> 31) or (slot == 1 and not embed), "arc weights are >= 0"
                    wt = np.frombuffer(np.uint32(wb & 0x7fffffff).tobytes(), np.float32)[0]
                    if wt != 0.0:
                        pending[lane].append((ab & 0xffff, ab >> 16, wt))
            if flags[0]:
                slices_seen += 1
                seen_w += 1
                codes = blk[:, 1, 0].astype(np.int32) if embed else v["rows"][rp: rp + 32]
                rp += 32
                for lane in range(32):
                    code = int(np.int32(codes[lane]))
                    row = (code & 0xfffff) if code >= 0 else ((sub_owner[-2 - code] & 0xfffff) if code <= -2 else None)
                    assert row is not None or not pending[lane]
                    out += [(row,) + t for t in pending[lane]]
                    pending[lane] = []
        assert not any(pending) and (seen_w == third if embed else True)
    assert slices_seen == v["n_slices"]
    return sorted(out)


@pytest.mark.parametrize("variant,S,K,D", [("uniform", 70, 400, 24), ("skew", 90, 1500, 16)])
def test_den_packer_is_lossless(variant, S, K, D, monkeypatch):
    g = ChainGraph(synth.random_den_fst(S, K, D, seed=9, variant=variant))
    gb = ChainGraphBatch(g, 2)
    gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_PROB)
    engine._prepared_cache.clear()
    monkeypatch.setenv("CHAIN_B200_MAX_CLASSES", "0")   # the arc-per-record packing (classes: next test)
    prep = engine.prepare_den_graph(gt, g.num_states, D, "cpu")
    b = _parse_blob(prep.blob_host)
    assert (b["S"], b["K"], b["D"], b["nw"]) == (S, K, D, 16) and b["total"] == prep.blob_host.numel()
    assert b["C"] == 0 and b["R_pad"] == b["S_pad"]
    ft, fp = g.forward_transitions.numpy(), g.forward_transition_probs.numpy()
    want = [
        sorted((int(d), int(s), int(n), np.float32(p)) for (s, d, n), p in zip(ft, fp)),   # alpha: row=dst, a=src, b=pdf
        sorted((int(s), int(d), int(n), np.float32(p)) for (s, d, n), p in zip(ft, fp)),   # beta:  row=src, a=dst, b=pdf
        sorted((int(n), int(s), int(d), np.float32(p)) for (s, d, n), p in zip(ft, fp)),   # gamma: row=pdf, a=src, b=dst
    ]
    for i in range(3):
        got = _decode_view(b["views"][i], 16)
        assert len(got) == K
        assert [(r, a, bb) for r, a, bb, _ in got] == [(r, a, bb) for r, a, bb, _ in want[i]]
        assert np.allclose(sorted(w for *_, w in got), sorted(w for *_, w in want[i]))
    if variant == "skew":
        assert b["views"][0]["n_sub"] > 0 and b["max_sub"] >= b["views"][0]["n_sub"]
    assert np.array_equal(b["leaky"], g.leaky_probs.numpy())
    # the packer's bank-conflict model improves on the unarranged order and is >= the ideal
    assert all(b["pred"][i] >= b["pred"][5 + i] for i in range(3))
    os.environ["CHAIN_B200_PACK_EFFORT"] = "0"
    try:
        engine._prepared_cache.clear()
        base = _parse_blob(engine.prepare_den_graph(gt, g.num_states, D, "cpu").blob_host)
    finally:
        os.environ.pop("CHAIN_B200_PACK_EFFORT")
        engine._prepared_cache.clear()
    assert sum(b["pred"][:3]) < sum(base["pred"][:3])


def _blob_forward_backward(b, x, coeff):
    """The factorised recursion of den_shared_kernel, run in float64 on the DECODED views of a blob for one
    utterance (x = exp of the clamped nnet output, [T,D]): class sums through view 3, one arc per class in views 0
    and 2, the one-operand beta' arcs of view 4 added to the rows of view 1 that carry the extra flag."""
    S, D, S_pad, C = b["S"], b["D"], b["S_pad"], b["C"]
    T = x.shape[0]
    lk, ini, fin = (b[k].astype(np.float64) for k in ("leaky", "init", "final"))
    v = [_decode_view(b["views"][i], b["nw"]) for i in range(5)]
    flagged = set()
    v1 = b["views"][1]
    if b["embed"]:
        codes = [int(np.int32(c)) for c, wb in v1["recs"] if wb >> 31]
        for c, wb in v1["recs"]:   # the row-code slots of view 1 carry leaky(row)
            if wb >> 31 and np.int32(c) >= 0:
                lv = np.frombuffer(np.uint32(wb & 0x7fffffff).tobytes(), np.float32)[0]
                assert lv == b["leaky"][int(np.int32(c)) & 0xfffff]
    else:
        codes = [int(c) for c in v1["rows"]]
    for code in codes + [int(r) for r, _, _ in v1["split"]]:
        if code >= 0 and code & 0x40000000:
            flagged.add(code & 0xfffff)
    assert flagged == {r for r, _, _, _ in v[4]}
    R = b["R_pad"]

    def two(view, A, Bv, n_rows):
        out = np.zeros(n_rows)
        for row, a_, b_, w in view:
            out[row] += w * A[a_] * Bv[b_]
        return out

    def one(view, A, n_rows):
        out = np.zeros(n_rows)
        for row, a_, _, w in view:
            out[row] += w * A[a_]
        return out

    alpha = np.zeros((T + 1, R))
    tot = np.zeros(T + 1)
    alpha[0, :S] = ini
    tot[0] = ini.sum()
    alpha[0, :S] += coeff * tot[0] * lk
    for t in range(1, T + 1):
        xs = x[t - 1] / tot[t - 1]
        alpha[t - 1, S_pad:S_pad + max(C, 1)][:C] = one(v[3], alpha[t - 1], max(C, 1))[:C]
        a = two(v[0], alpha[t - 1], xs, S)
        tot[t] = a.sum()
        alpha[t, :S] = a + coeff * tot[t] * lk
    Z = (alpha[T, :S] * fin).sum()
    logp = np.log(tot[:T]).sum() + np.log(Z)
    beta = np.zeros(R)
    beta[:S] = fin / Z + coeff * (fin * lk).sum() / Z
    grad = np.zeros((T, D))
    for t in range(T - 1, -1, -1):
        xs = x[t] / tot[t]
        for c in range(C):
            beta[S_pad + c] = beta[b["cls_dst"][c]] * xs[b["cls_pdf"][c]]
        grad[t] = two(v[2], alpha[t], beta, D) * xs
        bp = two(v[1], beta, xs, S)
        extra = one(v[4], beta, S)
        for r in flagged:
            bp[r] += extra[r]
        beta[:S] = bp + coeff * (bp * lk).sum()
    return logp, grad


@pytest.mark.parametrize("variant,S,K,D,min_class", [("biphone", 150, 1400, 40, 4), ("biphone", 150, 1400, 40, 2),
                                                     ("uniform", 60, 900, 3, 2), ("skew", 80, 1500, 2, 3)])
@pytest.mark.parametrize("embed", [0, 1])
def test_den_packer_class_factorisation(variant, S, K, D, min_class, embed, monkeypatch):
    """(dst,pdf) classes: the packed views + class table reproduce the oracle's forward-backward exactly (up to
    float64 reordering) -- structured graphs (pdf a function of the destination), graphs where only some arcs
    share a class (few pdfs), and heavy rows split into sub-rows."""
    from oracle import chain_oracle as co
    g = ChainGraph(synth.random_den_fst(S, K, D, seed=19, variant=variant))
    gb = ChainGraphBatch(g, 1)
    gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_PROB)
    monkeypatch.setenv("CHAIN_B200_MIN_CLASS", str(min_class))
    monkeypatch.setenv("CHAIN_B200_EMBED_CODES", str(embed))    # both stream formats (den_blob.h)
    engine._prepared_cache.clear()
    b = _parse_blob(engine.prepare_den_graph(gt, g.num_states, D, "cpu").blob_host)
    engine._prepared_cache.clear()
    assert b["embed"] == embed
    assert b["version"] == 13 and b["C"] > 0 and b["R_pad"] == b["S_pad"] + b["C_pad"]
    # every arc is either a two-operand record of views 0/1/2 or a member of exactly one class (views 3 and 4)
    n_cls_arcs = len(_decode_view(b["views"][3], 16))
    assert n_cls_arcs == len(_decode_view(b["views"][4], 16)) >= min_class * b["C"]
    assert len(_decode_view(b["views"][1], 16)) + n_cls_arcs == g.num_transitions
    assert len(_decode_view(b["views"][0], 16)) == g.num_transitions - n_cls_arcs + b["C"]
    assert len(_decode_view(b["views"][2], 16)) == g.num_transitions - n_cls_arcs + b["C"]
    T = 9
    x, lengths = synth.make_inputs(1, T, D, seed=20)
    ex = x.clamp(-30, 30).exp()
    o_c, lp_c, g_c, _, _ = co.den_c(gb, ex, lengths, 1e-5)
    logp, grad = _blob_forward_backward(b, ex[0].double().numpy(), 1e-5)
    assert abs(logp - o_c) < 2e-6 * abs(o_c) + 5e-6   # the oracle is fp32
    assert np.abs(grad - g_c[0].numpy()).max() < 5e-6


def test_den_packer_rejects_bad_graphs():
    lib = _cabi.lib()
    g = ChainGraph(synth.random_den_fst(10, 30, 8, seed=1))
    ft, fi, bt, bi = (g.forward_transitions.clone(), g.forward_transition_indices.clone(),
                      g.backward_transitions.clone(), g.backward_transition_indices.clone())
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    assert lib.chain_b200_den_prepare_bytes(p(ft), p(fi), p(bt), p(bi), 10, 30, 8, 0) > 0
    assert lib.chain_b200_den_prepare_bytes(p(ft), p(fi), p(bt), p(bi), 10, 30, 4, 0) == _cabi.ERR_INVALID_ARG  # pdf out of range
    bad = ft.clone(); bad[3, 1] = 99
    assert lib.chain_b200_den_prepare_bytes(p(bad), p(fi), p(bt), p(bi), 10, 30, 8, 0) == _cabi.ERR_INVALID_ARG
    assert lib.chain_b200_den_prepare_bytes(p(ft), p(fi), p(bt), p(bi), 70000, 30, 8, 0) == _cabi.ERR_UNSUPPORTED
    assert lib.chain_b200_den_prepare_bytes(p(ft), p(fi), p(bt), p(bi), 40000, 30, 8, _cabi.LOG_DOMAIN) == _cabi.ERR_UNSUPPORTED
    assert lib.chain_b200_den_workspace_bytes(p(torch.zeros(512, dtype=torch.uint8)), 4, 10) == 0  # not a blob


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing in the shipped package may import, load or call it."""
    pkg = os.path.join(ROOT, "pychain_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h")):
                continue
            text = open(os.path.join(dirpath, f)).read()
            for line in text.splitlines():
                code = line.split("#")[0].split("//")[0]
                assert not re.search(r"\b(import|from)\s+oracle\b", code), (f, line)
                assert "chain_oracle" not in code and "oracle/_ref" not in code, (f, line)
    hdr = open(os.path.join(ROOT, "include", "chain_b200.h")).read()
    assert "NO CPU implementation" in hdr


def test_shared_batch_to_keeps_stride0_and_reorder_keeps_device():
    """`.to(device)` of a shared (stride-0) batch must not materialise B copies, and `reorder` must
    stay on that device."""
    g = ChainGraph(synth.random_den_fst(12, 40, 8, seed=3))
    gb = ChainGraphBatch(g, 5).to(torch.device("cpu"))
    assert gb.shared_graph is g
    for name in engine._GRAPH_NAMES_PROB:
        t = getattr(gb, name)
        assert t.size(0) == 5 and t.stride(0) == 0
    gb.reorder(torch.tensor([2, 0, 1]))
    assert gb.batch_size == 3 and gb.forward_transitions.stride(0) == 0 and gb.start_state.numel() == 3
    assert engine._is_shared(tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_PROB), 3)


def test_fst_from_text_formats():
    """acceptor-format text must not be misread as transducer text (and vice versa)."""
    tr = ["0 1 3 3 0.5", "1 1 2 2 0.25", "1 0.125"]
    f = Fst.from_text(tr)
    assert f.num_states() == 2 and f.num_arcs == 2 and list(f.pdf) == [2, 1]
    assert np.allclose(f.log_weight, [-0.5, -0.25]) and np.isclose(f.final_log_weight[1], -0.125)
    ac = ["0 1 3 0.5", "1 1 2 0.25", "1 0.125"]
    fa = Fst.from_text(ac, acceptor=True)
    assert list(fa.pdf) == [2, 1] and np.allclose(fa.log_weight, [-0.5, -0.25])
    with pytest.raises(ValueError):
        Fst.from_text(["0 1 3", "1"])                       # 3 fields: acceptor text without the flag
    with pytest.raises(ValueError):
        Fst.from_text(["0 1 3 0.5"], acceptor=False)        # olabel is not an integer
    with pytest.raises(ValueError):
        Fst.from_text(["0 1 3 3 0.5 7"])                    # too many fields


def test_materialised_batch_detection_host_side():
    """A materialised batch of identical rows is recognised (rows 0, B//2, B-1 of every tensor); one differing row
    at a sampled position, or a per-utterance batch, is not."""
    g = ChainGraph(synth.random_den_fst(12, 40, 8, seed=3))
    mat = ChainGraphBatch(g, 6).materialize()
    gt = tuple(getattr(mat, n) for n in engine._GRAPH_NAMES_PROB)
    rows, key = engine._shared_rows(gt, 6)
    assert rows is not None and key is not None and rows[0].shape[0] == 1
    mat.forward_transition_probs[3] *= 0.5
    assert engine._shared_rows(tuple(getattr(mat, n) for n in engine._GRAPH_NAMES_PROB), 6) == (None, None)
    nb = synth.make_num_batch(4, 30, 8, seed=1, min_phones=2, max_phones=4)
    assert engine._shared_rows(tuple(getattr(nb, n) for n in engine._GRAPH_NAMES_LOG), 4) == (None, None)


def test_headline_kernel_stream_banks_do_not_wait_for_younger_loads():
    """ptxas signals every record-stream load of a SELL pass on one scoreboard; if it schedules a register bank's
    first consumer AFTER the other bank's refill, every loop iteration waits a full L2 latency (measured: cfg4 5.3 ->
    5.6-6.1 ms).  Source semantics cannot express it, so the built SASS of the cfg4 instantiation is checked
    (tools/sass_scoreboards.py decodes the control words; profiles/r2_variants.md has the story)."""
    import shutil
    import subprocess
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    tool = os.path.join(ROOT, "tools", "sass_scoreboards.py")
    out = subprocess.run([sys.executable, tool, _cabi.LIB_PATH, "den_shared_kernelILi4ELb0ELb1E"],
                         capture_output=True, text=True, timeout=300).stdout
    assert "den_shared_kernelILi4ELb0ELb1E" in out, out[-500:]
    assert out.strip().splitlines()[-1] == "bad bank pairs: 0", out[-2000:]


def _edge_graphs():
    """Hand-made denominator graphs at the corners of the packer: (name, Fst, D)."""
    out = []
    # one state: a self loop and a final weight
    out.append(("one_state", Fst(1, 0, [0], [0], [0], [np.log(0.5)], [np.log(0.5)]), 1))
    # a left-to-right chain with self loops: every slice is one arc pair or less (thin slices)
    S = 70
    src = list(range(S)) + list(range(S - 1))
    dst = list(range(S)) + list(range(1, S))
    pdf = [i % 7 for i in range(S)] + [(3 * i) % 7 for i in range(S - 1)]
    fin = np.full(S, -np.inf, np.float32); fin[-1] = np.log(0.5)
    out.append(("chain", Fst(S, 0, src, dst, pdf, np.full(len(src), np.log(0.5), np.float32), fin), 7))
    # a star: 100 sources all enter state 0 on the same pdf (ONE class of 100 arcs: longer than a slice is wide),
    # state 0 fans back out on distinct pdfs
    S = 101
    src = list(range(1, S)) + [0] * (S - 1) + [0]
    dst = [0] * (S - 1) + list(range(1, S)) + [0]
    pdf = [2] * (S - 1) + [i % 5 for i in range(S - 1)] + [4]
    w = np.log(np.r_[np.full(S - 1, 0.9), np.full(S - 1, 0.009), 0.05]).astype(np.float32)
    fin = np.full(S, np.log(0.1), np.float32)
    out.append(("star", Fst(S, 0, src, dst, pdf, w, fin), 5))
    # a single pdf: every destination is one class; a few states have no entering arc but the start
    rng = np.random.default_rng(4)
    S, K = 40, 300
    src = np.r_[np.arange(S), rng.integers(0, S, K - S)]
    dst = np.r_[(np.arange(S) + 1) % S, rng.integers(0, S // 2, K - S)]
    out.append(("one_pdf", Fst(S, 0, src, dst, np.zeros(K, int), np.full(K, np.log(0.1), np.float32),
                                np.full(S, np.log(0.2), np.float32)), 1))
    return out


@pytest.mark.parametrize("embed", [0, 1])
@pytest.mark.parametrize("name", ["one_state", "chain", "star", "one_pdf"])
def test_den_packer_corner_graphs(name, embed, monkeypatch):
    """Corner graphs through the packer: every arc is packed exactly once (as a two-operand record or as a class
    member), and the factorised recursion run from the decoded blob reproduces the oracle."""
    from oracle import chain_oracle as co
    fst, D = next((f, d) for n, f, d in _edge_graphs() if n == name)
    g = ChainGraph(fst, initial_mode="leaky", final_mode="fst")
    gb = ChainGraphBatch(g, 1)
    gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_PROB)
    monkeypatch.setenv("CHAIN_B200_MIN_CLASS", "2")
    monkeypatch.setenv("CHAIN_B200_EMBED_CODES", str(embed))
    engine._prepared_cache.clear()
    b = _parse_blob(engine.prepare_den_graph(gt, g.num_states, D, "cpu").blob_host)
    engine._prepared_cache.clear()
    n_cls = len(_decode_view(b["views"][3], 16))
    assert n_cls == len(_decode_view(b["views"][4], 16))
    assert len(_decode_view(b["views"][1], 16)) + n_cls == g.num_transitions
    assert len(_decode_view(b["views"][0], 16)) == len(_decode_view(b["views"][2], 16)) == g.num_transitions - n_cls + b["C"]
    if name in ("star", "one_pdf"):
        assert b["C"] > 0
    T = 6
    x, lengths = synth.make_inputs(1, T, D, seed=21)
    ex = x.clamp(-30, 30).exp()
    o_c, _, g_c, _, _ = co.den_c(gb, ex, lengths, 1e-5)
    logp, grad = _blob_forward_backward(b, ex[0].double().numpy(), 1e-5)
    assert abs(logp - o_c) < 2e-6 * abs(o_c) + 5e-6
    assert np.abs(grad - g_c[0].numpy()).max() < 5e-6


@pytest.mark.parametrize("S,log_domain,ok", [(65535, False, True), (65536, False, False), (32767, True, True), (32768, True, False)])
def test_den_packer_index_limits_at_the_boundary(S, log_domain, ok):
    """16-bit state indices in the packed records (15-bit in the log domain, whose index word carries the slice flag):
    the largest graph that fits packs, one state more is CHAIN_B200_ERR_UNSUPPORTED (the engine then takes the
    general kernel) -- never a silently wrapped index."""
    lib = _cabi.lib()
    src = torch.arange(S, dtype=torch.int32)
    dst = (src + 1) % S
    ft = torch.stack([src, dst, src % 8], 1).contiguous()          # a ring: arc s -> s+1, pdf s mod 8
    fi = torch.stack([src, src + 1], 1).contiguous()
    order = torch.argsort(dst.long(), stable=True)
    bt, bi = ft[order].contiguous(), fi.clone()
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    flags = _cabi.LOG_DOMAIN if log_domain else 0
    got = lib.chain_b200_den_prepare_bytes(p(ft), p(fi), p(bt), p(bi), S, S, 8, flags)
    if not ok:
        assert got == _cabi.ERR_UNSUPPORTED
        return
    assert got > 0
    if not log_domain:      # the largest state index survives the packing: decode the blob of the alpha view
        g = ChainGraph(Fst(S, 0, src.numpy(), dst.numpy(), (src % 8).numpy(), np.full(S, np.log(0.5), np.float32),
                           np.full(S, np.log(0.5), np.float32)))
        gb = ChainGraphBatch(g, 1)
        gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_PROB)
        engine._prepared_cache.clear()
        try:
            b = _parse_blob(engine.prepare_den_graph(gt, S, 8, "cpu").blob_host)
        finally:
            engine._prepared_cache.clear()
        rows = _decode_view(b["views"][0], 16)
        assert len(rows) == S and (0, S - 1, (S - 1) % 8) in {(r, a, bb) for r, a, bb, _ in rows}
