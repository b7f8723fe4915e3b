import numpy as np


class Space(object):
    pass


class Discrete(Space):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.int64


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low = low
        self.high = high
        self.shape = tuple(shape) if shape is not None else np.shape(low)
        self.dtype = dtype


class Tuple(Space):
    def __init__(self, spaces):
        self.spaces = tuple(spaces)


class MultiDiscrete(Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec)
