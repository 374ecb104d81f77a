"""Cubic B-spline monomial tables for the spline-Jastrow basis.

The reference stores piece ``p`` of basis spline ``s`` as monomial coefficients in ABSOLUTE ``r``:
``B_s(r) = w[s][p][0] + w[s][p][1] r + w[s][p][2] r^2 + w[s][p][3] r^3`` on ``(t_{s+p}, t_{s+p+1}]``
(``SplineFactory::GetWeights3``, src/SplineFactory.cpp:54-101, closed-form expressions).  This module
builds the same table independently from the Cox-de Boor recursion carried out on polynomial
coefficients in extended precision.  The two tables agree to rounding of the individual coefficients
(~1e-16 relative), NOT bit for bit -- and because the absolute-``r`` monomial form is ill-conditioned
(coefficients reach 1e6 while |B| <= 2/3) a 1-ulp coefficient difference moves B by up to ~1e-10.
For 1e-10 parity with the reference the C-ABI therefore takes the table from the caller: the
reference driver passes its own ``SplineFactory`` output, the parity tests pass the table stored
in ``tests/golden``.  Synthetic workloads (bench, statistics tests) use this builder.
"""
import numpy as np

__all__ = ["bspline_monomial_weights", "extend_knots_mirrored", "uniform_knots"]


def uniform_knots(n_param, half_length):
    """Knots ``(i * L/2) / (P - 1)`` for ``i = -3 .. P+2`` (BosonsBulk::InitSystem, BosonsBulk.cpp:61-67)."""
    i = np.arange(-3, n_param + 3, dtype=np.float64)
    return (i * half_length) / (float(n_param) - 1.0)


def extend_knots_mirrored(grid):
    """``SetNodes``: pad a NURBS grid by three mirrored knots on each side (BosonsBulk.cpp:35-44)."""
    n = [float(x) for x in grid]
    nodes = list(n)
    for i in range(3):
        nodes.insert(0, -n[i + 1])
        nodes.append(2.0 * n[-1] - n[len(n) - 2 - i])
    return np.array(nodes, dtype=np.float64)


def bspline_monomial_weights(knots):
    """Return ``w[K][4][4]`` (spline, piece, monomial power) for cubic B-splines on ``knots``.

    Cox-de Boor: ``B_{i,0} = 1`` on ``[t_i, t_{i+1})`` and
    ``B_{i,d}(r) = (r - t_i)/(t_{i+d} - t_i) B_{i,d-1}(r) + (t_{i+d+1} - r)/(t_{i+d+1} - t_{i+1}) B_{i+1,d-1}(r)``,
    evaluated symbolically: every ``B_{i,d}`` restricted to knot interval ``j`` is a polynomial whose
    coefficients are carried in ``np.longdouble``.
    """
    t = np.asarray(knots, dtype=np.longdouble)
    nk = len(t)
    # polys[d][i][j] = coefficients (power 0..3) of B_{i,d} on interval [t_j, t_{j+1}), j in i..i+d
    prev = [{i: np.array([1, 0, 0, 0], dtype=np.longdouble)} for i in range(nk - 1)]
    for d in range(1, 4):
        cur = []
        for i in range(nk - d - 1):
            pieces = {}
            den_a = t[i + d] - t[i]
            den_b = t[i + d + 1] - t[i + 1]
            for j in range(i, i + d + 1):
                acc = np.zeros(4, dtype=np.longdouble)
                if j in prev[i] and den_a != 0:
                    q = prev[i][j]
                    # (r - t_i)/den_a * q(r)
                    acc[1:] += q[:3] / den_a
                    acc += -t[i] / den_a * q
                if j in prev[i + 1] and den_b != 0:
                    q = prev[i + 1][j]
                    # (t_{i+d+1} - r)/den_b * q(r)
                    acc[1:] -= q[:3] / den_b
                    acc += t[i + d + 1] / den_b * q
                pieces[j] = acc
            cur.append(pieces)
        prev = cur
    K = nk - 4
    w = np.zeros((K, 4, 4), dtype=np.float64)
    for s in range(K):
        for p in range(4):
            w[s, p, :] = prev[s][s + p].astype(np.float64)
    return w
