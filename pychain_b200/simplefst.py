"""Stand-in for the reference's `simplefst` module (openfst_binding/src/fstext.cc:174-184) without OpenFst.

`simplefst.StdVectorFst` is `pychain_b200.graph.Fst`: `read`, `read_ark`, `write`, `fst_to_tensor`,
`set_leaky_probs`, `num_states`, `start_state` have the reference's names, argument order and return layouts, so
recipe code (`simplefst.StdVectorFst.read(den_fst_path)`) and the reference's own `pychain/graph.py:44,48,61`
run over it unchanged.  `install_as_pychain()` registers it under the name `simplefst` when no real one exists.
"""
from .graph import Fst as StdVectorFst  # noqa: F401

__all__ = ["StdVectorFst"]
