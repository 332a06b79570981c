"""pychain_b200 -- a B200-native (sm_100a) LF-MMI chain-loss engine behind pychain's API.

    from pychain_b200 import ChainLoss, ChainFunction, ChainGraph, ChainGraphBatch, Fst

mirrors `from pychain import *` (/root/reference/pychain/__init__.py:1-2).
`pychain_b200.pychain_C` mirrors the reference's native module.  Call
`install_as_pychain()` to register both under the reference's module names.
"""
from .graph import *  # noqa: F401,F403
from .loss import *  # noqa: F401,F403


def install_as_pychain():
    """Make `import pychain`, `import pychain.loss`, `import pychain.graph` and
    `import pychain_C` resolve to this package (drop-in for existing recipes); `import simplefst` resolves to
    the OpenFst-free stand-in unless the reference's real binding is installed."""
    import sys
    from . import graph, loss, pychain_C, simplefst
    me = sys.modules[__name__]
    sys.modules.setdefault("pychain", me)
    sys.modules.setdefault("pychain.loss", loss)
    sys.modules.setdefault("pychain.graph", graph)
    sys.modules.setdefault("pychain_C", pychain_C)
    if graph.simplefst is None:
        sys.modules.setdefault("simplefst", simplefst)
