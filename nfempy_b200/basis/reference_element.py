"""Reference-cell finite elements on simplices: the part of basix the reference touches.

The reference builds ``basix.create_element(family, celltype, degree, lagrange_variant=gll_centroid,
discontinuous)`` (R/src/basis/finite_element.py:85-91, R/src/spaces/dof_map.py:16-22) and reads
``tabulate(1, pts)``, ``entity_dofs``, ``num_entity_dofs``, ``dof_transformations_are_identity`` and
``entity_transformations()`` (R/src/basis/finite_element.py:98-108, 156-165).  basix is not installable here,
so this module restates those elements from their published definition (Ciarlet triple: polynomial space +
dual functionals, basis = dual basis).  Sub-entity numbering follows the convention the reference hard-codes
(R/src/mesh/mesh.py:187-203, R/src/basis/permute_and_transform.py:9-11, 39-61).

Every element is stored as coefficients over monomials ``x^a y^b z^c``: ``phi[i, c] = sum_m coeffs[i, c, m] *
mono_m``.  Coefficients come from a dual-matrix inverse done in 60-digit arithmetic, so tables are correct to
the last bit or two and independent of how the oracle computes the same functions.
"""
import functools
import itertools

import mpmath
import numpy as np

from .quadrature import make_quadrature

mpmath.mp.dps = 60

REF_VERTICES = {
    1: np.array([[0.0], [1.0]]),
    2: np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]),
    3: np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]),
}
# sub-entity -> local vertices
SUB_ENTITIES = {
    1: {0: [[0], [1]], 1: [[0, 1]]},
    2: {0: [[0], [1], [2]], 1: [[1, 2], [0, 2], [0, 1]], 2: [[0, 1, 2]]},
    3: {
        0: [[0], [1], [2], [3]],
        1: [[2, 3], [1, 3], [1, 2], [0, 3], [0, 2], [0, 1]],
        2: [[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]],
        3: [[0, 1, 2, 3]],
    },
}


def monomial_exponents(dim, degree):
    """All exponent tuples of total degree <= ``degree``, graded then lexicographic."""
    out = []
    for total in range(degree + 1):
        for e in itertools.product(range(total + 1), repeat=dim):
            if sum(e) == total:
                out.append(e)
    return np.array(out, dtype=np.int64).reshape(-1, dim)


def tabulate_monomials(exps, points, nderiv=1):
    """``[1+dim (or 1), nq, n_mono]``: monomial values and first derivatives at ``points``."""
    points = np.asarray(points, dtype=np.float64)
    nq, dim = points.shape
    nm = len(exps)
    out = np.zeros((1 + (dim if nderiv >= 1 else 0), nq, nm))
    pw = np.ones((dim, int(exps.max(initial=0)) + 1, nq))
    for d in range(dim):
        for p in range(1, pw.shape[1]):
            pw[d, p] = pw[d, p - 1] * points[:, d]
    for m, e in enumerate(exps):
        v = np.ones(nq)
        for d in range(dim):
            v = v * pw[d, e[d]]
        out[0, :, m] = v
        if nderiv >= 1:
            for dd in range(dim):
                if e[dd] == 0:
                    continue
                v = np.full(nq, float(e[dd]))
                for d in range(dim):
                    v = v * pw[d, e[d] - (1 if d == dd else 0)]
                out[1 + dd, :, m] = v
    return out


def _gll_unit(n):
    """n+1 Gauss-Lobatto-Legendre points on [0, 1] (n <= 3 in closed form)."""
    if n == 1:
        return [mpmath.mpf(0), mpmath.mpf(1)]
    if n == 2:
        return [mpmath.mpf(0), mpmath.mpf(1) / 2, mpmath.mpf(1)]
    if n == 3:
        g = (1 - 1 / mpmath.sqrt(5)) / 2
        return [mpmath.mpf(0), g, 1 - g, mpmath.mpf(1)]
    raise NotImplementedError("Lagrange degree > 3 (gll_centroid interior placement) is not restated")


def _lagrange_nodes(dim, degree):
    """Nodes (mp numbers) grouped by sub-entity, in basix order; interior points per entity listed along the
    entity's own reference direction (lower local vertex -> higher local vertex)."""
    verts = [[mpmath.mpf(v) for v in row] for row in REF_VERTICES[dim].tolist()]
    nodes, entity_dofs = [], [[[] for _ in SUB_ENTITIES[dim][d]] for d in range(dim + 1)]
    if degree == 0:
        c = [sum(v[i] for v in verts) / (dim + 1) for i in range(dim)]
        entity_dofs[dim][0].append(0)
        return [c], entity_dofs
    g = _gll_unit(degree)
    for d in range(dim + 1):
        for e, lv in enumerate(SUB_ENTITIES[dim][d]):
            if d == 0:
                pts = [verts[lv[0]]]
            elif d == 1:
                a, b = verts[lv[0]], verts[lv[1]]
                pts = [[a[i] + t * (b[i] - a[i]) for i in range(dim)] for t in g[1:-1]]
            elif d == 2:
                if degree < 3:
                    pts = []
                elif degree == 3:
                    pts = [[sum(verts[v][i] for v in lv) / 3 for i in range(dim)]]
            else:
                if degree < 4:
                    pts = []
            for p in pts:
                entity_dofs[d][e].append(len(nodes))
                nodes.append(p)
    return nodes, entity_dofs


def _mp_inverse_to_float(rows):
    inv = mpmath.matrix(rows) ** -1
    return np.array([[float(inv[i, j]) for j in range(inv.cols)] for i in range(inv.rows)])


class ReferenceElement:
    """Attribute surface of a basix element as far as the reference reads it."""

    def __init__(self, family, dim, degree, discontinuous=False):
        self.family = family
        self.dim = dim
        self.degree = degree
        self.discontinuous = discontinuous
        if dim == 0:
            self.value_size = 1
            self.exps = np.zeros((1, 0), dtype=np.int64)
            self.coeffs = np.ones((1, 1, 1))
            self._entity_dofs = [[[0]]]
            self._interval_T = self._triangle_T = None
        elif family == "Lagrange":
            self._build_lagrange()
        elif family in ("RT", "BDM"):
            from .hdiv import build_hdiv

            build_hdiv(self)
        elif family in ("N1E", "N2E"):
            from .hcurl import build_hcurl

            build_hcurl(self)
        else:
            raise NotImplementedError("ReferenceElement: family %r is not restated" % family)
        self.n_phi = self.coeffs.shape[0]
        if discontinuous:
            every = sorted(i for per_d in self._entity_dofs for e in per_d for i in e)
            self._entity_dofs = [[[] for _ in per_d] for per_d in self._entity_dofs]
            self._entity_dofs[dim][0] = every

    # -- construction -----------------------------------------------------------------------------
    def _build_lagrange(self):
        dim, k = self.dim, self.degree
        if k == 0 and not self.discontinuous:
            raise ValueError("Lagrange degree 0 must be discontinuous")
        self.value_size = 1
        self.exps = monomial_exponents(dim, k)
        nodes, self._entity_dofs = _lagrange_nodes(dim, k)
        vand = [[mpmath.fprod(x[d] ** int(e[d]) for d in range(dim)) for e in self.exps] for x in nodes]
        # phi_i = sum_m C[m, i] mono_m with V C = I
        self.coeffs = _mp_inverse_to_float(vand).T[:, None, :].copy()
        self.nodes = np.array([[float(v) for v in x] for x in nodes])
        self._interval_T = self._triangle_T = None

    # -- basix-like API ---------------------------------------------------------------------------
    @property
    def entity_dofs(self):
        return self._entity_dofs

    @property
    def num_entity_dofs(self):
        return [[len(e) for e in per_d] for per_d in self._entity_dofs]

    @property
    def dof_transformations_are_identity(self):
        if self.discontinuous or self.dim < 2:
            return True
        if self.family == "Lagrange":
            return self.degree <= 2
        return False

    def entity_transformations(self):
        from .transformations import entity_transformations

        return entity_transformations(self)

    def tabulate(self, nderiv, points):
        """``[1 + dim, nq, n_phi, value_size]`` for ``nderiv == 1`` (``[1, ...]`` for 0)."""
        points = np.asarray(points, dtype=np.float64).reshape(-1, max(self.dim, 1))
        if self.dim == 0:
            return np.ones((1, 1, 1, 1))
        mono = tabulate_monomials(self.exps, points, nderiv)
        return np.einsum("dqm,icm->dqic", mono, self.coeffs)


@functools.lru_cache(maxsize=None)
def create_element(family, dim, degree, discontinuous=False):
    return ReferenceElement(family, dim, degree, discontinuous)
