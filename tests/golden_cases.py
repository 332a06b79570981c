"""Loads tests/golden/*.npz and rebuilds the seeded inputs of each case."""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)


def load(name):
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    return {k: z[k] for k in z.files}
