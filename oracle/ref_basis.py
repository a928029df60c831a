"""ORACLE (test infrastructure only -- never imported by the product path).

Stand-in for the basix calls the reference makes, written from the elements' textbook definitions and kept
independent of ``nfempy_b200.basis`` (closed forms / scipy instead of the product's dual-matrix tables):

* ``gauss_jacobi``             -- ``basix.make_quadrature(cell, degree, gauss_jacobi)`` (R/src/spaces/product_space.py:38-40)
  via ``scipy.special.roots_jacobi``;
* ``tabulate``                 -- ``element.tabulate(1, pts)`` -> ``[1+dim, nq, nphi, vs]`` (R/src/basis/finite_element.py:162);
* ``evaluate_linear_shapes`` / ``evaluate_mapping`` -- R/src/geometry/mapping.py:9-55, including the per-point QR;
* ``push_forward``             -- identity for Lagrange, contravariant Piola for RT/BDM (R/src/basis/finite_element.py:163-170).

PARITY UNPINNED against real basix: basix is an un-vendored, unpinned dependency (R/requirements.txt:2) that cannot
be installed here, and the reference's tests hold no golden vectors for this path.  Lagrange
P1/P2 nodal bases are unique, so those tables are determined; H(div) tables follow basix's documented construction.
Everything the reference does ABOVE these calls is pinned: tools/make_reference_pins.py runs the reference's own source with
this module standing in for basix (tools/ref_shims/basix) and tests/test_cpu_reference_pins.py compares the oracle with what it
produced (tests/golden/refcode_*.npz); what stays unpinned is exactly this module's and ref_hdiv.py's element tables.
"""
import numpy as np
from scipy.special import roots_jacobi

SUB = {
    1: {0: [[0], [1]], 1: [[0, 1]]},
    2: {0: [[0], [1], [2]], 1: [[1, 2], [0, 2], [0, 1]], 2: [[0, 1, 2]]},
    3: {
        0: [[0], [1], [2], [3]],
        1: [[2, 3], [1, 3], [1, 2], [0, 3], [0, 2], [0, 1]],
        2: [[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]],
        3: [[0, 1, 2, 3]],
    },
}


def gauss_jacobi(dim, degree):
    m = (degree + 2) // 2
    px, wx = roots_jacobi(m, 0.0, 0.0)
    if dim == 1:
        return np.array([[0.5 * (1 + x)] for x in px]), 0.5 * wx
    py, wy = roots_jacobi(m, 1.0, 0.0)
    if dim == 2:
        pts, wts = [], []
        for i in range(m):
            for j in range(m):
                pts.append([0.25 * (1 + px[i]) * (1 - py[j]), 0.5 * (1 + py[j])])
                wts.append(wx[i] * wy[j] * 0.125)
        return np.array(pts), np.array(wts)
    pz, wz = roots_jacobi(m, 2.0, 0.0)
    pts, wts = [], []
    for i in range(m):
        for j in range(m):
            for k in range(m):
                pts.append(
                    [
                        0.125 * (1 + px[i]) * (1 - py[j]) * (1 - pz[k]),
                        0.25 * (1 + py[j]) * (1 - pz[k]),
                        0.5 * (1 + pz[k]),
                    ]
                )
                wts.append(wx[i] * wy[j] * wz[k] * 0.125 * 0.125)
    return np.array(pts), np.array(wts)


def num_entity_dofs(family, dim, k, discontinuous):
    if dim == 0:
        return [[1]]
    if family == "Lagrange":
        if k == 0:
            per = [0] * dim + [1]
        else:
            per = [1, k - 1, (k - 1) * (k - 2) // 2, (k - 1) * (k - 2) * (k - 3) // 6][: dim + 1]
    elif family == "RT":
        facet = k if dim == 2 else k * (k + 1) // 2
        interior = dim * (k * (k - 1) // 2 if dim == 2 else k * (k - 1) * (k + 1) // 6)
        per = [0] * (dim - 1) + [facet, interior]
    elif family == "BDM":
        facet = k + 1 if dim == 2 else (k + 1) * (k + 2) // 2
        interior = (k - 1) * (k + 1) if dim == 2 else (k - 1) * (k + 1) * (k + 2) // 2
        per = [0] * (dim - 1) + [facet, interior]
    elif family == "N1E":
        if k != 1:
            raise NotImplementedError("oracle N1E degree >= 2")
        per = [0, 1, 0, 0][: dim + 1]  # one tangential moment per edge
    elif family == "N2E":
        if k != 1:
            raise NotImplementedError("oracle N2E degree >= 2")
        per = [0, 2, 0, 0][: dim + 1]  # two tangential moments per edge, nothing else at degree 1
    else:
        raise NotImplementedError(family)
    out = [[per[d]] * len(SUB[dim][d]) for d in range(dim + 1)]
    if discontinuous:
        total = sum(sum(c) for c in out)
        out = [[0] * len(SUB[dim][d]) for d in range(dim + 1)]
        out[dim][0] = total
    return out


def _barycentric(dim, pts):
    lam = np.concatenate([1.0 - pts.sum(axis=1, keepdims=True), pts], axis=1)  # [nq, dim+1]
    dlam = np.concatenate([-np.ones((1, dim)), np.eye(dim)], axis=0)  # [dim+1, dim]
    return lam, dlam


def tabulate_lagrange(dim, k, pts):
    """Closed-form P0/P1/P2 (numerical Vandermonde for P3) values and reference gradients, basix dof order (vertices,
    edges, faces)."""
    pts = np.asarray(pts, dtype=float).reshape(-1, dim)
    nq = len(pts)
    lam, dlam = _barycentric(dim, pts)
    if k == 0:
        tab = np.zeros((1 + dim, nq, 1, 1))
        tab[0] = 1.0
        return tab
    if k == 1:
        tab = np.zeros((1 + dim, nq, dim + 1, 1))
        tab[0, :, :, 0] = lam
        for d in range(dim):
            tab[1 + d, :, :, 0] = dlam[:, d]
        return tab
    if k == 2:
        edges = SUB[dim][1]
        nphi = dim + 1 + len(edges)
        tab = np.zeros((1 + dim, nq, nphi, 1))
        for v in range(dim + 1):
            tab[0, :, v, 0] = lam[:, v] * (2 * lam[:, v] - 1)
            for d in range(dim):
                tab[1 + d, :, v, 0] = (4 * lam[:, v] - 1) * dlam[v, d]
        for e, (a, b) in enumerate(edges):
            tab[0, :, dim + 1 + e, 0] = 4 * lam[:, a] * lam[:, b]
            for d in range(dim):
                tab[1 + d, :, dim + 1 + e, 0] = 4 * (lam[:, a] * dlam[b, d] + lam[:, b] * dlam[a, d])
        return tab
    if k == 3:
        return _tabulate_lagrange_p3(dim, pts)
    raise NotImplementedError("oracle Lagrange degree %d" % k)


def _tabulate_lagrange_p3(dim, pts):
    """Cubic nodal basis for the ``gll_centroid`` nodes (vertices, two Gauss-Lobatto points per edge listed from the edge's
    lower to its higher local vertex, face centroids), by a numerical Vandermonde solve in the monomial basis."""
    verts = np.concatenate([np.zeros((1, dim)), np.eye(dim)], axis=0)
    g = (1.0 - 1.0 / np.sqrt(5.0)) / 2.0
    nodes = [v for v in verts]
    for a, b in SUB[dim][1]:
        nodes += [verts[a] + t * (verts[b] - verts[a]) for t in (g, 1.0 - g)]
    if dim >= 2:
        for lv in SUB[dim][2]:
            nodes.append(verts[list(lv)].mean(axis=0))
    nodes = np.array(nodes)
    exps = [e for total in range(4) for e in np.ndindex(*([total + 1] * dim)) if sum(e) == total]

    def mono(x):
        val = np.stack([np.prod(x ** np.array(e), axis=1) for e in exps], axis=1)
        grad = np.zeros((dim, len(x), len(exps)))
        for m, e in enumerate(exps):
            for d in range(dim):
                if e[d] > 0:
                    ee = np.array(e)
                    ee[d] -= 1
                    grad[d, :, m] = e[d] * np.prod(x**ee, axis=1)
        return val, grad

    C = np.linalg.inv(mono(nodes)[0])  # column i: monomial coefficients of the nodal function of node i
    val, grad = mono(np.asarray(pts, dtype=float).reshape(-1, dim))
    tab = np.zeros((1 + dim, len(val), len(nodes), 1))
    tab[0, :, :, 0] = val @ C
    for d in range(dim):
        tab[1 + d, :, :, 0] = grad[d] @ C
    return tab


def tabulate(family, dim, k, pts, discontinuous=False):
    if dim == 0:
        return np.ones((1, 1, 1, 1))
    if family == "Lagrange":
        return tabulate_lagrange(dim, k, pts)
    from . import ref_hdiv

    return ref_hdiv.tabulate(family, dim, k, pts)


def evaluate_linear_shapes(dim, points):
    return tabulate_lagrange(dim, 1, points)


def _det_and_pseudo_inverse(grad_xmap):  # R/src/geometry/mapping.py:23-36
    q, r = np.linalg.qr(grad_xmap)
    sign = np.diag(np.sign(np.diag(r)), 0)
    q = np.dot(q, sign)
    r = np.dot(sign, r)
    det = np.linalg.det(r)
    inv = np.dot(np.linalg.inv(r), q.T)
    return det, inv


def evaluate_mapping(dim, phi, cell_points):  # R/src/geometry/mapping.py:39-55
    if dim == 0:
        one = np.array([1.0])
        return cell_points, one, one, one
    x = phi[0, :, :, 0] @ cell_points
    jac = np.rollaxis(phi[list(range(1, dim + 1)), :, :, 0], 1) @ cell_points
    jac = np.transpose(jac, (0, 2, 1))
    det, inv = zip(*[_det_and_pseudo_inverse(j) for j in jac])
    return x, jac, np.array(det), np.array(inv)


def push_forward(family, phi0, jac, det, inv):
    """basix ``push_forward`` on the value slab ``[nq, nphi, vs]``: identity (Lagrange) or contravariant Piola
    ``(1/det) J phi`` (RT/BDM) with the reference's positive ``det`` and the 3 x dim Jacobian."""
    if family == "Lagrange":
        return phi0
    return np.einsum("q,qij,qpj->qpi", 1.0 / det, jac, phi0)
