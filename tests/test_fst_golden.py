"""`pychain_b200.graph.Fst` against the reference's own openfst_binding/src/fstext.cc.

The fixtures tests/golden/fst_*.npz / fst_*.fst were produced by the reference's UNMODIFIED fstext.cc (compiled
where it lies over oracle/fst_shim, see tests/golden/make_fst_golden.py): `fst_to_tensor` (fstext.cc:19-117) in both
domains, `set_leaky_probs` (fstext.cc:120-171), and the binary file as the reference module wrote it.  Where
oracle/_ref/simplefst.so is present the same comparison also runs live on fresh random graphs, and the reference's
unmodified pychain/graph.py (staged into oracle/_ref/) is run over it and compared field by field with
pychain_b200's ChainGraph / ChainGraphBatch.
"""
import glob
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pychain_b200.graph import Fst, ChainGraph, ChainGraphBatch  # noqa: E402
from oracle import build_ref  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "fst_*.npz")))
TENSOR_NAMES = ("forward_transitions", "forward_transition_probs", "forward_transition_indices",
                "backward_transitions", "backward_transition_probs", "backward_transition_indices", "final_probs")


def _case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    f = Fst(int(z["num_states"]), int(z["start_state"]), z["src"], z["dst"], z["pdf"], z["log_weight"],
            z["final_log_weight"])
    return z, f


def _same(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (what, a.shape, b.shape, a.dtype, b.dtype)
    if a.dtype.kind in "iu":
        assert np.array_equal(a, b), what
    else:  # the same float32 exp of the same float32 numbers: equal up to the last bit of libm vs torch
        np.testing.assert_allclose(a, b, rtol=2e-7, atol=0, err_msg=what)


def test_golden_cases_exist():
    assert len(CASES) >= 4


@pytest.mark.parametrize("name", CASES)
def test_to_tensors_matches_reference_fst_to_tensor(name):
    z, f = _case(name)
    for dom, flag in (("prob", False), ("log", True)):
        got = f.to_tensors(flag)
        for n, t in zip(TENSOR_NAMES, got):
            _same(t.numpy(), z[dom + "_" + n], name + ":" + dom + "_" + n)
    assert Fst.num_states(f) == int(z["num_states"]) and Fst.start_state(f) == int(z["start_state"])


@pytest.mark.parametrize("name", CASES)
def test_leaky_probs_matches_reference_set_leaky_probs(name):
    z, f = _case(name)
    got = f.leaky_probs().numpy()
    ref = z["leaky_probs"]
    assert got.dtype == ref.dtype == np.float32
    # float64 recursions in a different summation order, rounded to float32 at the end
    np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(Fst.set_leaky_probs(f).numpy(), ref, rtol=1e-6, atol=1e-12)


@pytest.mark.parametrize("name", CASES)
def test_binary_file_round_trip(name, tmp_path):
    """Reads the file the reference module wrote; writes the same bytes back."""
    z, f = _case(name)
    path = os.path.join(GOLDEN, name + ".fst")
    g = Fst.read(path)
    assert g.num_states() == f.num_states() and g.start_state() == f.start_state()
    for a in ("src", "dst", "pdf", "log_weight", "final_log_weight"):
        assert np.array_equal(getattr(g, a), getattr(f, a)), a
    out = str(tmp_path / "w.fst")
    assert g.write(out)
    assert open(out, "rb").read() == open(path, "rb").read()
    # an archive entry: the same FST behind some leading bytes (fstext.cc:7-16 seeks to the offset)
    ark = str(tmp_path / "a.ark")
    with open(ark, "wb") as fh:
        fh.write(b"utt-0001 ")
        fh.write(open(path, "rb").read())
    h = Fst.read_ark(ark, 9)
    assert np.array_equal(h.dst, f.dst) and np.array_equal(h.log_weight, f.log_weight)


def test_binary_reader_refuses_what_it_cannot_parse(tmp_path):
    good = open(os.path.join(GOLDEN, "fst_tiny.fst"), "rb").read()
    p = str(tmp_path / "x.fst")
    open(p, "wb").write(b"\x00" * 16)
    with pytest.raises(ValueError, match="magic"):
        Fst.read(p)
    open(p, "wb").write(good[:len(good) - 7])
    with pytest.raises(ValueError, match="truncated"):
        Fst.read(p)
    open(p, "wb").write(good.replace(b"standard", b"log\x00\x00\x00\x00\x00"))
    with pytest.raises(ValueError):
        Fst.read(p)
    flagged = bytearray(good)
    off = 4 + 4 + 6 + 4 + 8 + 4   # magic, "vector", "standard", version -> flags
    flagged[off] = 1
    open(p, "wb").write(bytes(flagged))
    with pytest.raises(ValueError, match="symbol"):
        Fst.read(p)


# ---------------------------------------------------------------------------------------------- live reference
def _ref_simplefst():
    if build_ref.build_simplefst() is None:
        pytest.skip("reference simplefst not built and /root/reference absent")
    return build_ref.load_simplefst()


def _random_fst(S, K, D, seed, start=0):
    sys.path.insert(0, GOLDEN)
    from make_fst_golden import random_fst
    return random_fst(S, K, D, 0.3, start, seed)


@pytest.mark.parametrize("S,K,D,seed", [(5, 9, 4, 1), (64, 700, 50, 2), (500, 3000, 200, 3)])
def test_live_against_reference_fstext(S, K, D, seed, tmp_path):
    ref = _ref_simplefst().StdVectorFst
    f = _random_fst(S, K, D, seed, start=S // 2)
    path = str(tmp_path / "g.fst")
    f.write(path)
    g = ref.read(path)
    assert ref.num_states(g) == S and ref.start_state(g) == S // 2
    for flag in (False, True):
        for n, a, b in zip(TENSOR_NAMES, f.to_tensors(flag), ref.fst_to_tensor(g, flag)):
            _same(a.numpy(), b.numpy(), n)
    np.testing.assert_allclose(f.leaky_probs().numpy(), ref.set_leaky_probs(g).numpy(), rtol=1e-6, atol=1e-12)


def test_reference_graph_py_over_reference_simplefst_equals_ours(tmp_path):
    """The reference's UNMODIFIED pychain/graph.py (ChainGraph :23-70, ChainGraphBatch :73-194) driven by the
    reference-compiled simplefst, against pychain_b200.graph on the same FSTs: every field of both containers."""
    mod = _ref_simplefst()
    staged = build_ref.stage_reference_graph()
    if staged is None:
        pytest.skip("reference pychain/graph.py not staged")
    saved = sys.modules.get("simplefst")
    sys.modules["simplefst"] = mod
    try:
        spec = importlib.util.spec_from_file_location("reference_pychain_graph", staged)
        rg = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(rg)
    finally:
        if saved is None:
            sys.modules.pop("simplefst", None)
        else:
            sys.modules["simplefst"] = saved

    fields = ChainGraph._GRAPH_FIELDS
    ours, theirs = [], []
    for i, (S, K, D) in enumerate([(6, 12, 5), (30, 140, 20), (17, 60, 9)]):
        f = _random_fst(S, K, D, 100 + i, start=i)
        path = str(tmp_path / "g{}.fst".format(i))
        f.write(path)
        g = mod.StdVectorFst.read(path)
        for kw in (dict(), dict(initial_mode="leaky", final_mode="ones"), dict(log_domain=True),
                   dict(log_domain=True, final_mode="ones")):
            a, b = ChainGraph(f, **kw), rg.ChainGraph(g, **kw)
            assert (a.num_states, a.num_transitions, a.start_state, a.is_empty, a.log_domain) == \
                   (b.num_states, b.num_transitions, b.start_state, b.is_empty, b.log_domain)
            for n in fields:
                x, y = getattr(a, n), getattr(b, n)
                assert (x is None) == (y is None), n
                if x is not None:
                    assert x.dtype == y.dtype and x.shape == y.shape, n
                    if x.dtype.is_floating_point:
                        torch.testing.assert_close(x, y, rtol=1e-6, atol=1e-12, equal_nan=True)
                    else:
                        assert torch.equal(x, y), n
        ours.append(ChainGraph(f, log_domain=True))
        theirs.append(rg.ChainGraph(g, log_domain=True))

    # a batch of different graphs (the numerator case) and a batch of one shared graph (the denominator case)
    def same_batch(a, b):
        assert (a.batch_size, a.num_states, a.log_domain) == (b.batch_size, b.num_states, b.log_domain)
        if hasattr(b, "num_transitions"):   # the reference sets it only for a list of graphs (graph.py:127)
            assert a.num_transitions == b.num_transitions
        for n in ChainGraphBatch._BATCH_FIELDS:
            x, y = getattr(a, n), getattr(b, n)
            assert (x is None) == (y is None), n
            if x is None:
                continue
            assert x.dtype == y.dtype and tuple(x.shape) == tuple(y.shape), (n, x.dtype, y.dtype, x.shape, y.shape)
            if x.dtype.is_floating_point:
                torch.testing.assert_close(x.contiguous(), y, rtol=1e-6, atol=1e-12, equal_nan=True)
            else:
                assert torch.equal(x.contiguous(), y), n

    for bad in (ChainGraphBatch, rg.ChainGraphBatch):   # both refuse a list without the padded sizes (graph.py:88-96)
        with pytest.raises(ValueError, match="max_num_transitions should be specified"):
            bad(ours if bad is ChainGraphBatch else theirs)
    same_batch(ChainGraphBatch(ours, max_num_transitions=140, max_num_states=30),
               rg.ChainGraphBatch(theirs, max_num_transitions=140, max_num_states=30))
    same_batch(ChainGraphBatch(ours, max_num_transitions=200, max_num_states=40),
               rg.ChainGraphBatch(theirs, max_num_transitions=200, max_num_states=40))
    den_o, den_r = ChainGraph(_random_fst(30, 140, 20, 101, 1)), None
    path = str(tmp_path / "den.fst")
    _random_fst(30, 140, 20, 101, 1).write(path)
    den_r = rg.ChainGraph(mod.StdVectorFst.read(path))
    same_batch(ChainGraphBatch(den_o, 5), rg.ChainGraphBatch(den_r, 5))
    a = ChainGraphBatch(ours, max_num_transitions=150, max_num_states=33)
    b = rg.ChainGraphBatch(theirs, max_num_transitions=150, max_num_states=33)
    order = torch.tensor([2, 0, 1])
    a.reorder(order), b.reorder(order)
    same_batch(a, b)


@pytest.mark.parametrize("name", CASES)
def test_reference_graph_py_over_our_simplefst_stand_in(name):
    """The other direction: the reference's unmodified pychain/graph.py driven by pychain_b200.simplefst (what a
    recipe gets after install_as_pychain() on a machine without OpenFst) reproduces the golden tensors."""
    staged = build_ref.stage_reference_graph()
    if staged is None:
        pytest.skip("reference pychain/graph.py not staged")
    import pychain_b200.simplefst as ours
    saved = sys.modules.get("simplefst")
    sys.modules["simplefst"] = ours
    try:
        spec = importlib.util.spec_from_file_location("reference_pychain_graph2", staged)
        rg = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(rg)
    finally:
        if saved is None:
            sys.modules.pop("simplefst", None)
        else:
            sys.modules["simplefst"] = saved
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    fst = ours.StdVectorFst.read(os.path.join(GOLDEN, name + ".fst"))
    for dom, flag in (("prob", False), ("log", True)):
        g = rg.ChainGraph(fst, log_domain=flag)
        assert g.num_states == int(z["num_states"]) and g.start_state == int(z["start_state"])
        for n in TENSOR_NAMES:
            _same(getattr(g, n).numpy(), z[dom + "_" + n], n)
        if not flag:
            np.testing.assert_allclose(g.leaky_probs.numpy(), z["leaky_probs"], rtol=1e-6, atol=1e-12)
