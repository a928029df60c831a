"""Gauss-Jacobi quadrature on reference simplices, restating ``basix.make_quadrature(cell, degree,
QuadratureType.gauss_jacobi)`` as the reference calls it (R/src/spaces/product_space.py:31-50).

basix (un-vendored, unpinned; R/requirements.txt:2) builds the rule from ``m = (degree + 2) // 2`` Gauss-Jacobi
points per direction with Jacobi weights (1-x)^a, a = 0, 1, 2, in collapsed coordinates, points ordered by
nested loops over the 1-D rules (outermost = first coordinate).  Point order only affects round-off.
"""
import numpy as np


def gauss_jacobi_1d(alpha, m):
    """Nodes (ascending) and weights of the m-point Gauss rule for the weight (1-x)^alpha on [-1, 1]."""
    # Golub-Welsch on the Jacobi(alpha, 0) three-term recurrence, then Newton polish on P_m.
    k = np.arange(m, dtype=np.float64)
    a, b = float(alpha), 0.0
    with np.errstate(divide="ignore", invalid="ignore"):
        diag = (b * b - a * a) / ((2 * k + a + b) * (2 * k + a + b + 2))
    if a + b == 0.0:
        diag[0] = 0.0
    else:
        diag[0] = (b - a) / (a + b + 2)
    k1 = np.arange(1, m, dtype=np.float64)
    off = (
        2.0
        / (2 * k1 + a + b)
        * np.sqrt(k1 * (k1 + a) * (k1 + b) * (k1 + a + b) / ((2 * k1 + a + b - 1) * (2 * k1 + a + b + 1)))
    )
    if m > 1 and a + b == 0.0:
        off[0] = 2.0 / (2 + a + b) * np.sqrt((1 + a) * (1 + b) / (3 + a + b))
    T = np.diag(diag) + np.diag(off, 1) + np.diag(off, -1)
    x = np.linalg.eigvalsh(T)

    def jacobi(n, x):  # P_n^{(a,0)}(x) and derivative
        p0 = np.ones_like(x)
        if n == 0:
            return p0, np.zeros_like(x)
        p1 = 0.5 * (a + (a + 2) * x)
        for j in range(1, n):
            c1 = 2 * (j + 1) * (j + a + 1) * (2 * j + a)
            c2 = (2 * j + a + 1) * a * a
            c3 = (2 * j + a) * (2 * j + a + 1) * (2 * j + a + 2)
            c4 = 2 * (j + a) * j * (2 * j + a + 2)
            p0, p1 = p1, ((c2 + c3 * x) * p1 - c4 * p0) / c1
        # derivative via (2n+a)(1-x^2) P_n' = n(a - (2n+a) x) P_n + 2 n (n+a) P_{n-1}   (b = 0)
        dp = (n * (a - (2 * n + a) * x) * p1 + 2 * n * (n + a) * p0) / ((2 * n + a) * (1 - x * x))
        return p1, dp

    for _ in range(3):
        p, dp = jacobi(m, x)
        x = x - p / dp
    p, dp = jacobi(m, x)
    # w_i = Gamma-factor / ((1-x^2) P_m'(x)^2); for b = 0: 2^(a+1) (m+a)!/(m! ...) simplifies to 2^(a+1)
    w = 2.0 ** (a + 1) / ((1 - x * x) * dp * dp)
    return x, w


def make_quadrature(dim, degree):
    """``(points[nq, dim], weights[nq])`` on the reference simplex of dimension ``dim`` (1, 2 or 3)."""
    m = (int(degree) + 2) // 2
    px, wx = gauss_jacobi_1d(0.0, m)
    if dim == 1:
        return (0.5 * (1.0 + px))[:, None].copy(), 0.5 * wx
    py, wy = gauss_jacobi_1d(1.0, m)
    if dim == 2:
        i, j = np.meshgrid(np.arange(m), np.arange(m), indexing="ij")
        i, j = i.ravel(), j.ravel()
        pts = np.stack([0.25 * (1.0 + px[i]) * (1.0 - py[j]), 0.5 * (1.0 + py[j])], axis=1)
        return pts, wx[i] * wy[j] * 0.125
    pz, wz = gauss_jacobi_1d(2.0, m)
    if dim == 3:
        i, j, k = np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing="ij")
        i, j, k = i.ravel(), j.ravel(), k.ravel()
        pts = np.stack(
            [
                0.125 * (1.0 + px[i]) * (1.0 - py[j]) * (1.0 - pz[k]),
                0.25 * (1.0 + py[j]) * (1.0 - pz[k]),
                0.5 * (1.0 + pz[k]),
            ],
            axis=1,
        )
        return pts, wx[i] * wy[j] * wz[k] * 0.125 * 0.125
    raise ValueError("make_quadrature: dimension %r not supported" % dim)
