"""Functional numpy restatement of the slice of `hj_reachability==0.5.0` that the
reference's hot path calls. TEST INFRASTRUCTURE ONLY.

Call sites in the reference that this stands in for:
  * `Grid.interpolate`  - multiagent/safety_filter.py:195,245,348,418,
                          multiagent/core.py:463,
                          multiagent/custom_scenarios/navigation_graph_safe.py:751
  * `Grid.grad_values`  - multiagent/safety_filter.py:167
  * `sets.Box.extreme_point` - multiagent/safety_filter.py:70-78,250,423

DECLARED SEMANTICS (parity unpinned against the real library):
  interpolate(values, x):
    position = (x - lo) / spacing                        (float64)
    i_lo = floor(position); i_hi = i_lo + 1
    w_hi = position - i_lo;  w_lo = 1 - w_hi             (from the UNclipped i_lo)
    periodic dims: index mod shape (python sign);
    other dims:    index clipped to [0, shape-1]   (jax 0.4.x `x[idx]` gathers clamp
                   out-of-bounds indices, so a state outside the box reads the edge
                   node; nothing is NaN unless the state itself is NaN)
    result = sum over the 2^d corners, binary counting with dim 0 slowest,
             weight = ((w0*w1)*w2)..., term = weight * value, sequential adds.
    Grid data are stored float32 (JAX default dtype); arithmetic is float64.
  grad_values(values): central difference in the interior, one-sided first order at
    the two ends of a non-periodic dim, wrap-around on periodic dims
    (== numpy.gradient(edge_order=1) semantics), output float32, last axis = dim.
  Box.extreme_point(d) = where(d < 0, lo, hi) with lo/hi held in float32 (JAX default),
    so a bang-bang control is a float32 number.
"""
import numpy as np

# Second declared mode (oracle/ref_harness.make_env(interp_float32=True)): the same formula with state, domain_lo and
# spacing rounded to float32 and every operation - position, floor, weights, weight products, corner sum - in float32,
# which is what jax computes without jax_enable_x64 (the reference never enables it). Grid data are float32 either way.
FLOAT32_INTERPOLATION = False


class _Box(object):
    def __init__(self, lo, hi):
        self.lo = np.asarray(lo, dtype=np.float32)
        self.hi = np.asarray(hi, dtype=np.float32)

    def extreme_point(self, direction):
        direction = np.asarray(direction)
        return np.where(direction < 0, self.lo, self.hi)

    @property
    def ndim(self):
        return self.lo.shape[-1]


class sets(object):
    Box = _Box


class Grid(object):
    def __init__(self, domain_lo, domain_hi, shape, periodic_dims=()):
        self.lo = np.asarray(domain_lo, dtype=np.float64)
        self.hi = np.asarray(domain_hi, dtype=np.float64)
        self.shape = tuple(int(s) for s in shape)
        self.ndim = len(self.shape)
        per = np.zeros(self.ndim, dtype=bool)
        for d in (periodic_dims if periodic_dims is not None else ()):
            per[int(d)] = True
        self.periodic = per
        n = np.asarray(self.shape, dtype=np.float64)
        # periodic: linspace(lo, hi, n, endpoint=False); otherwise endpoint=True
        self.spacings = np.where(per, (self.hi - self.lo) / n, (self.hi - self.lo) / (n - 1.0))
        self.coordinate_vectors = [
            self.lo[d] + self.spacings[d] * np.arange(self.shape[d]) for d in range(self.ndim)]

    @classmethod
    def from_lattice_parameters_and_boundary_conditions(cls, domain, shape, periodic_dims=None, **_):
        return cls(domain.lo, domain.hi, shape, periodic_dims)

    def interpolate(self, values, state):
        values = np.asarray(values)
        if FLOAT32_INTERPOLATION:
            return self._interpolate_f32(values, state)
        state = np.asarray(state, dtype=np.float64)
        assert state.shape == (self.ndim,)
        position = (state - self.lo) / self.spacings
        if np.any(np.isnan(position)):
            return np.full(values.shape[self.ndim:], np.nan)
        position = np.clip(position, -1.0e9, 1.0e9)
        i_lo = np.floor(position)
        w_hi = position - i_lo
        w_lo = 1.0 - w_hi
        i_lo = i_lo.astype(np.int64)
        i_hi = i_lo + 1
        shape = np.asarray(self.shape, dtype=np.int64)
        idx = []
        for ind in (i_lo, i_hi):
            idx.append(np.where(self.periodic, np.mod(ind, shape), np.clip(ind, 0, shape - 1)))
        w = (w_lo, w_hi)
        out = np.zeros(values.shape[self.ndim:], dtype=np.float64)
        for corner in range(1 << self.ndim):
            weight = None
            index = []
            for d in range(self.ndim):
                bit = (corner >> (self.ndim - 1 - d)) & 1
                wd = w[bit][d]
                weight = wd if weight is None else weight * wd
                index.append(int(idx[bit][d]))
            out = out + weight * np.asarray(values[tuple(index)], dtype=np.float64)
        return out if out.shape else np.float64(out)

    def _interpolate_f32(self, values, state):
        f32 = np.float32
        state = np.asarray(state, dtype=np.float64).astype(f32)
        assert state.shape == (self.ndim,)
        position = (state - self.lo.astype(f32)) / self.spacings.astype(f32)
        assert position.dtype == f32
        if np.any(np.isnan(position)):
            return np.full(values.shape[self.ndim:], np.nan)
        position = np.minimum(np.maximum(position, f32(-1.0e9)), f32(1.0e9))
        i_lo = np.floor(position)
        w_hi = position - i_lo
        w_lo = f32(1.0) - w_hi
        i_lo = i_lo.astype(np.int64)
        i_hi = i_lo + 1
        shape = np.asarray(self.shape, dtype=np.int64)
        idx = [np.where(self.periodic, np.mod(ind, shape), np.clip(ind, 0, shape - 1)) for ind in (i_lo, i_hi)]
        w = (w_lo, w_hi)
        out = np.zeros(values.shape[self.ndim:], dtype=f32)
        for corner in range(1 << self.ndim):
            weight = None
            index = []
            for d in range(self.ndim):
                bit = (corner >> (self.ndim - 1 - d)) & 1
                wd = w[bit][d]
                weight = wd if weight is None else f32(weight * wd)
                index.append(int(idx[bit][d]))
            out = (out + f32(weight) * np.asarray(values[tuple(index)], dtype=f32)).astype(f32)
        return out if out.shape else f32(out)

    def grad_values(self, values, upwind_scheme=None):
        values = np.asarray(values, dtype=np.float64)
        grads = []
        for d in range(self.ndim):
            h = self.spacings[d]
            if self.periodic[d]:
                g = (np.roll(values, -1, axis=d) - np.roll(values, 1, axis=d)) / (2.0 * h)
            else:
                g = np.gradient(values, h, axis=d, edge_order=1)
            grads.append(g)
        return np.stack(grads, axis=-1).astype(np.float32)
