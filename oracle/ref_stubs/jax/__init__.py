"""Stand-in for jax: `jax.numpy` is numpy (float64), see ref_stubs/README.md."""
from . import numpy  # noqa: F401
