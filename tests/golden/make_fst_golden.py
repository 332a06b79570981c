#!/usr/bin/env python
"""Generate tests/golden/fst_*.npz + fst_*.fst from the reference's UNMODIFIED openfst_binding/src/fstext.cc.

oracle/build_ref.py compiles that file where it lies into oracle/_ref/simplefst.so, over oracle/fst_shim/ (a
stand-in for the OpenFst container classes: OpenFst itself is not in this image).  For every case below a random
FST is built, handed to the reference through a binary file, and the reference's `fst_to_tensor` (fstext.cc:19-117,
both domains) and `set_leaky_probs` (fstext.cc:120-171) outputs are stored, together with the file as written back
by the reference module's `write`.  tests/test_fst_golden.py checks `pychain_b200.graph.Fst` against them.

Run where /root/reference exists:   python tests/golden/make_fst_golden.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from pychain_b200.graph import Fst  # noqa: E402
from oracle import build_ref  # noqa: E402

CASES = {
    # name: (num_states, num_arcs, num_pdfs, final_fraction, start_state, seed)
    "fst_tiny": (7, 15, 5, 0.5, 0, 11),
    "fst_start3": (40, 260, 31, 0.2, 3, 12),
    "fst_mid": (300, 2400, 120, 0.05, 0, 13),
    "fst_allfinal": (25, 90, 9, 1.0, 24, 14),
}


def random_fst(S, K, D, final_fraction, start, seed):
    """Arcs in shuffled insertion order (so that per-state order is not sorted by anything), parallel arcs, self
    loops; every state has a final weight or at least one arc (fstext.cc:139 asserts a positive total)."""
    rng = np.random.default_rng(seed)
    src = rng.integers(0, S, size=K)
    dst = rng.integers(0, S, size=K)
    src[:S] = np.arange(S)                       # every state has an outgoing arc
    dst[:S] = (np.arange(S) + 1) % S             # ... along a ring, so that every state is reachable
    dst[S:S + S // 4] = src[S:S + S // 4]        # self loops
    if K > 2 * S:                                # parallel arcs between one pair of states
        src[-3:], dst[-3:] = src[0], dst[0]
    pdf = rng.integers(0, D, size=K)
    cost = rng.uniform(0.05, 3.0, size=K).astype(np.float32)
    perm = rng.permutation(K)
    fin = np.full(S, -np.inf, dtype=np.float32)
    is_final = rng.random(S) < final_fraction
    is_final[S - 1] = True
    fin[is_final] = -rng.uniform(0.0, 2.0, size=int(is_final.sum())).astype(np.float32)
    return Fst(S, start, src[perm], dst[perm], pdf[perm], -cost[perm], fin)


def main():
    if build_ref.build_simplefst() is None:
        raise SystemExit("reference simplefst unavailable (no /root/reference)")
    ref = build_ref.load_simplefst().StdVectorFst
    for name, (S, K, D, ff, start, seed) in CASES.items():
        f = random_fst(S, K, D, ff, start, seed)
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "in.fst")
            f.write(path)
            g = ref.read(path)
            assert g is not None, name
            out = {"num_states": np.int64(ref.num_states(g)), "start_state": np.int64(ref.start_state(g))}
            names = ("forward_transitions", "forward_transition_probs", "forward_transition_indices",
                     "backward_transitions", "backward_transition_probs", "backward_transition_indices", "final_probs")
            for dom, flag in (("prob", False), ("log", True)):
                for n, t in zip(names, ref.fst_to_tensor(g, flag)):
                    out[dom + "_" + n] = t.numpy()
            out["leaky_probs"] = ref.set_leaky_probs(g).numpy()
            back = os.path.join(HERE, name + ".fst")
            assert ref.write(g, back)
        # the case itself, so that the test does not depend on this generator
        out.update(src=f.src, dst=f.dst, pdf=f.pdf, log_weight=f.log_weight, final_log_weight=f.final_log_weight)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "S", S, "K", K, "leaky sum", float(out["leaky_probs"].sum()), os.path.getsize(back), "bytes")


if __name__ == "__main__":
    main()
