"""Minimal space objects. The runners only read `.shape`, `.n` and `__class__.__name__`
(reference onpolicy/runner/shared/base_runner.py:94-143, graph_mpe_runner.py:26), so `gym` is not needed."""
from __future__ import annotations

import numpy as np


class Box:
    def __init__(self, low, high, shape, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def __repr__(self):
        return f"Box{self.shape}"


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.int64

    def __repr__(self):
        return f"Discrete({self.n})"
