"""include/lsm_math.h: the float64 sin / cos / atan2 shared by the CUDA kernels, the host side of liblsm_b200.so and
the C oracle (numpy's libm in the reference: multiagent/core.py:105-131,179-181, safety_filter.py:277-284,
navigation_graph_safe.py:606-656, utils.py:79-349).

CPU: accuracy against numpy (sin / cos <= 1 ulp apart, atan2 <= 2 ulp) and bit-identity of the two host
builds (gcc -ffp-contract=off in the oracle, nvcc's host compiler in the product library).
GPU: the device evaluation is BIT-IDENTICAL to the host evaluation - which is what lets the parity tests demand
bit-exact discrete outputs and float64 states from the CUDA path."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'oracle'))


def _inputs(n=400_000, seed=0):
    rng = np.random.default_rng(seed)
    x = np.concatenate([rng.uniform(-8, 8, n), rng.uniform(-200, 200, n), rng.uniform(-1.5e6, 1.5e6, n // 4),
                        rng.uniform(-1e-3, 1e-3, n // 4) * 10.0 ** rng.integers(-20, 0, n // 4),
                        np.array([0.0, -0.0, np.pi, -np.pi, np.pi / 2, np.pi / 4, 1e-300, 5e-324, 0.7853981633974483,
                                  0.7853981633974484, 2.356194490192345, np.inf, -np.inf, np.nan, 1e7, 1e300])])
    y = np.concatenate([rng.uniform(-5, 5, n), rng.uniform(-1e-3, 1e-3, n), rng.uniform(-5, 5, n // 2),
                        np.array([0.0, -0.0, 0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.inf, 1.0, np.nan, 1e-310, 1e300, 3.0])])
    z = np.concatenate([rng.uniform(-5, 5, n), rng.uniform(-5, 5, n), rng.uniform(-1e-9, 1e-9, n // 2),
                        np.array([0.0, 0.0, -0.0, -0.0, 0.0, -0.0, np.inf, np.inf, -np.inf, np.inf, 1.0, 1e300, 1e-310, 1.0])])
    return x, y, z


def _ulps(got, want):
    fin = np.isfinite(want)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    sp = np.spacing(np.abs(want[fin]))
    return np.abs(got[fin] - want[fin]) / sp


def _product_eval(op, a, b=None):
    from layered_safe_marl_b200 import _lib
    lib = _lib.load()
    a = np.ascontiguousarray(a, dtype=np.float64)
    out = np.empty_like(a)
    bp = None if b is None else np.ascontiguousarray(b, dtype=np.float64).ctypes.data_as(C.c_void_p)
    _lib.check(lib.lsm_math_eval(op, a.ctypes.data_as(C.c_void_p), bp, out.ctypes.data_as(C.c_void_p), a.size), 'lsm_math_eval')
    return out


def test_accuracy_against_libm():
    import oracle_env as O
    x, y, z = _inputs()
    with np.errstate(invalid='ignore'):
        small = np.abs(x) < 1.6e6                   # the Cody-Waite range; beyond it the header is deterministic, not accurate
        xs = x[small | ~np.isfinite(x)]
        assert _ulps(O.math_eval(0, xs), np.sin(xs)).max() <= 1.0
        assert _ulps(O.math_eval(1, xs), np.cos(xs)).max() <= 1.0
        got, want = O.math_eval(2, y, z), np.arctan2(y, z)
    assert _ulps(got, want).max() <= 2.0     # one rounding of y / x on top of atan's < 1 ulp
    assert np.array_equal(np.signbit(got[want == 0]), np.signbit(want[want == 0]))
    # the half-plane conventions the reference relies on: atan2(0, 0) = 0 (stopped double integrator, core.py:179-181)
    assert O.math_eval(2, np.array([0.0]), np.array([0.0]))[0] == 0.0
    # out-of-contract arguments stay finite and inside [-1, 1]
    big = O.math_eval(0, np.array([1e7, -3e9, 1e300]))
    assert np.all(np.abs(big) <= 1.0)


def test_product_host_build_is_bit_identical_to_the_oracle_build():
    import oracle_env as O
    x, y, z = _inputs(seed=1)
    for op in (0, 1):
        assert np.array_equal(_product_eval(op, x), O.math_eval(op, x), equal_nan=True)
    assert np.array_equal(_product_eval(2, y, z), O.math_eval(2, y, z), equal_nan=True)
    # the one-reduction sincos returns exactly sin and cos
    assert np.array_equal(_product_eval(3, x), O.math_eval(0, x), equal_nan=True)
    assert np.array_equal(_product_eval(4, x), O.math_eval(1, x), equal_nan=True)


@pytest.mark.gpu
def test_device_evaluation_is_bit_identical_to_the_host():
    import torch
    from layered_safe_marl_b200 import _lib
    lib = _lib.load()
    x, y, z = _inputs(n=1_000_000, seed=2)
    dev = torch.device('cuda:0')
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def on_device(op, a, b=None):
        ta = torch.from_numpy(a).to(dev)
        tb = torch.from_numpy(b).to(dev) if b is not None else None
        out = torch.empty_like(ta)
        _lib.check(lib.lsm_math_eval_device(op, C.c_void_p(ta.data_ptr()), C.c_void_p(tb.data_ptr()) if tb is not None else None,
                                            C.c_void_p(out.data_ptr()), ta.numel(), stream), 'lsm_math_eval_device')
        torch.cuda.synchronize()
        return out.cpu().numpy()

    for op in (0, 1, 3, 4):
        got, want = on_device(op, x), _product_eval(op, x)
        assert np.array_equal(got.view(np.int64)[~np.isnan(want)], want.view(np.int64)[~np.isnan(want)]), f"op {op}"
        assert np.array_equal(np.isnan(got), np.isnan(want))
    got, want = on_device(2, y, z), _product_eval(2, y, z)
    assert np.array_equal(got.view(np.int64)[~np.isnan(want)], want.view(np.int64)[~np.isnan(want)])
    assert np.array_equal(np.isnan(got), np.isnan(want))
