"""Closed-form stand-in for the one cvxpy program the reference solves
(multiagent/safety_filter.py:286-308 and :364-376). TEST INFRASTRUCTURE ONLY.

    minimise (u - r)^T P (u - r)   s.t.  a.u + b >= 0        (P diagonal, > 0)

DECLARED SEMANTICS (replaces cvxpy 1.4.1 + OSQP, whose ~1e-5 iterate noise is not
reproducible): with s = a.r + b,
    s >= 0                  -> u = r
    a^T P^-1 a == 0         -> infeasible -> u.value is None (caller falls back to r)
    otherwise               -> u = r - (s / a^T P^-1 a) * P^-1 a
"""
import numpy as np


class _Affine(object):
    """A @ u + b with u the single Variable of the problem."""
    __array_ufunc__ = None  # make numpy defer `ndarray @ _Affine`, `ndarray + _Affine`

    def __init__(self, var, A, b):
        self.var = var
        self.A = np.asarray(A, dtype=np.float64)
        self.b = np.asarray(b, dtype=np.float64)

    def __add__(self, other):
        if isinstance(other, _Affine):
            return _Affine(self.var, self.A + other.A, self.b + other.b)
        return _Affine(self.var, self.A, self.b + np.asarray(other, dtype=np.float64))

    __radd__ = __add__

    def __sub__(self, other):
        if isinstance(other, _Affine):
            return _Affine(self.var, self.A - other.A, self.b - other.b)
        return _Affine(self.var, self.A, self.b - np.asarray(other, dtype=np.float64))

    def __rsub__(self, other):
        return _Affine(self.var, -self.A, np.asarray(other, dtype=np.float64) - self.b)

    def __neg__(self):
        return _Affine(self.var, -self.A, -self.b)

    def __rmatmul__(self, M):
        # explicit sequential products/sums (no BLAS, no FMA) so the operation order is pinned
        M = np.asarray(M, dtype=np.float64)
        A = np.atleast_2d(self.A)
        b = np.atleast_1d(self.b)
        if M.ndim == 2:
            rows = [self.__rmatmul__(M[i]) for i in range(M.shape[0])]
            return _Affine(self.var, np.stack([r.A for r in rows]), np.stack([r.b for r in rows]))
        n, m = A.shape
        assert M.shape == (n,)
        outA = np.zeros(m)
        for j in range(m):
            acc = 0.0
            for k in range(n):
                acc = acc + M[k] * A[k, j]
            outA[j] = acc
        acc = 0.0
        for k in range(n):
            acc = acc + M[k] * b[k]
        return _Affine(self.var, outA, np.float64(acc))

    def __mul__(self, c):
        return _Affine(self.var, self.A * c, self.b * c)

    __rmul__ = __mul__

    def __ge__(self, other):
        return _Constraint(self - other)


class Variable(_Affine):
    def __init__(self, n):
        self.n = int(n)
        self.value = None
        _Affine.__init__(self, self, np.eye(self.n), np.zeros(self.n))


class _Constraint(object):
    def __init__(self, expr):  # expr >= 0
        self.expr = expr


class _QuadForm(object):
    def __init__(self, expr, P):
        self.expr = expr
        self.P = np.asarray(P, dtype=np.float64)


def quad_form(expr, P):
    return _QuadForm(expr, P)


class Minimize(object):
    def __init__(self, q):
        self.q = q


class Problem(object):
    def __init__(self, objective, constraints):
        self.objective = objective
        self.constraints = constraints

    def solve(self, *args, **kwargs):
        q = self.objective.q
        var = q.expr.var
        # objective expression is (I u - r)
        assert np.array_equal(q.expr.A, np.eye(var.n)), "objective must be quad_form(u - r, P)"
        r = -q.expr.b
        P = q.P
        assert np.array_equal(P, np.diag(np.diag(P))), "P must be diagonal"
        p_inv = 1.0 / np.diag(P)
        assert len(self.constraints) == 1
        c = self.constraints[0].expr
        a = np.asarray(c.A, dtype=np.float64).reshape(-1)
        b = float(np.asarray(c.b).reshape(-1)[0]) if np.ndim(c.b) else float(c.b)
        s = 0.0
        for k in range(var.n):  # sequential dot, the order the kernels use
            s = s + a[k] * r[k]
        s = s + b
        if s >= 0.0:
            var.value = r.copy()
            return 0.0
        denom = 0.0
        for k in range(var.n):
            denom = denom + (a[k] * p_inv[k]) * a[k]
        if denom == 0.0:
            var.value = None
            return np.inf
        lam = s / denom
        var.value = np.array([r[k] - lam * (p_inv[k] * a[k]) for k in range(var.n)])
        return 0.0
