"""CPU tests: quadrature and reference tables of the product vs independent statements (scipy, closed forms,
exact rationals)."""
from fractions import Fraction
from math import factorial

import numpy as np
import pytest
from scipy.special import roots_jacobi

from nfempy_b200.basis import create_element, make_quadrature
from nfempy_b200.basis.quadrature import gauss_jacobi_1d
from nfempy_b200.weak_forms.h1_common import reference_stiffness_tensor
from oracle import ref_basis


@pytest.mark.parametrize("alpha", [0.0, 1.0, 2.0])
@pytest.mark.parametrize("m", [1, 2, 3, 4, 5, 8])
def test_gauss_jacobi_matches_scipy(alpha, m):
    x, w = gauss_jacobi_1d(alpha, m)
    xs, ws = roots_jacobi(m, alpha, 0.0)
    assert np.allclose(x, xs, rtol=0, atol=1e-14) and np.allclose(w, ws, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("degree", [1, 3, 5, 7])
def test_quadrature_exact_on_monomials_and_matches_oracle(dim, degree):
    pts, wts = make_quadrature(dim, degree)
    m = (degree + 2) // 2
    assert pts.shape == (m**dim, dim)
    rp, rw = ref_basis.gauss_jacobi(dim, degree)
    assert np.allclose(pts, rp, atol=1e-14) and np.allclose(wts, rw, atol=1e-15)
    # int x^a y^b z^c over the reference simplex = a! b! c! / (a + b + c + dim)!
    for total in range(degree + 1):
        expo = [total] + [0] * (dim - 1) if dim == 1 else [total // 2, total - total // 2] + [0] * (dim - 2)
        exact = np.prod([factorial(e) for e in expo]) / factorial(sum(expo) + dim)
        val = (wts * np.prod(pts ** np.array(expo), axis=1)).sum()
        assert abs(val - exact) <= 1e-15 + 1e-13 * abs(exact)


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("k", [0, 1, 2])
def test_lagrange_tables_match_closed_forms(dim, k):
    el = create_element("Lagrange", dim, k, k == 0)
    pts, _ = make_quadrature(dim, 5)
    assert np.abs(el.tabulate(1, pts) - ref_basis.tabulate("Lagrange", dim, k, pts)).max() <= 2e-15
    assert el.num_entity_dofs == ref_basis.num_entity_dofs("Lagrange", dim, k, k == 0)
    assert el.dof_transformations_are_identity


def test_lagrange_p3_is_nodal_with_gll_edge_points():
    el = create_element("Lagrange", 3, 3)
    assert el.n_phi == 20 and not el.dof_transformations_are_identity
    vals = el.tabulate(0, el.nodes)[0, :, :, 0]
    assert np.abs(vals - np.eye(20)).max() < 1e-13
    g = (1 - 1 / np.sqrt(5)) / 2
    edge0 = el.nodes[el.entity_dofs[1][0]]  # edge (2,3): from vertex 2 to vertex 3
    assert np.allclose(edge0, [[0, 1 - g, g], [0, g, 1 - g]])


def test_p1_p2_reference_stiffness_in_exact_rationals():
    """Known answers: int grad(lambda_a).grad(lambda_b) on the reference tet, and the P2 vertex-vertex entries
    int (4 l_a - 1)(4 l_b - 1) = 1/10 (a == b) and -1/30 (a != b) times grad l_a . grad l_b."""
    dl = np.array([[-1, -1, -1], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
    pts, wts = make_quadrature(3, 5)
    m1 = reference_stiffness_tensor(create_element("Lagrange", 3, 1), pts, wts)
    exact = np.einsum("ak,bl->abkl", dl, dl) / 6.0
    assert np.abs(m1 - exact).max() < 1e-15
    m2 = reference_stiffness_tensor(create_element("Lagrange", 3, 2), pts, wts)
    for a in range(4):
        for b in range(4):
            c = Fraction(1, 10) if a == b else Fraction(-1, 30)
            assert np.abs(m2[a, b] - float(c) * np.outer(dl[a], dl[b])).max() < 1e-15
    # rows sum to zero (constants have zero gradient) and M[a,b,k,l] = M[b,a,l,k]
    assert np.abs(m2.sum(axis=1)).max() < 1e-15
    assert np.abs(m2 - m2.transpose(1, 0, 3, 2)).max() < 1e-16
