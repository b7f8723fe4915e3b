"""`jax.numpy` -> numpy. The reference only uses array/where/cos/sin/eye/zeros/ones."""
from numpy import *  # noqa: F401,F403
import numpy as _np

ndarray = _np.ndarray
