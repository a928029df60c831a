"""Stand-in for fenics-basix (absent here) behind the call surface the reference uses (SURVEY.md section 8c lists the call
sites).  Every answer but the N1E table (``_whitney``, textbook) comes from ``oracle/ref_basis.py`` / ``oracle/ref_hdiv.py``: the pins made with this stand-in cover the
reference's code ABOVE the basix boundary, not basix's own conventions (see tools/ref_shims/README.md)."""
import enum

import numpy as np

from oracle import ref_basis, ref_hdiv

from . import cell  # noqa: F401  (basix.cell.sub_entity_connectivity)


class CellType(enum.Enum):
    point = 0
    interval = 1
    triangle = 2
    tetrahedron = 3


class ElementFamily(enum.Enum):
    P = "Lagrange"
    RT = "RT"
    BDM = "BDM"
    N1E = "N1E"


class LagrangeVariant(enum.Enum):
    gll_centroid = 0


class QuadratureType(enum.Enum):
    gauss_jacobi = 0


def geometry(celltype):
    dim = celltype.value
    return np.concatenate([np.zeros((1, dim)), np.eye(dim)], axis=0)


def make_quadrature(celltype, degree, rule=QuadratureType.gauss_jacobi):
    return ref_basis.gauss_jacobi(celltype.value, int(degree))


class _Element:
    def __init__(self, family, dim, degree, discontinuous):
        self.family_name, self.dim, self.degree, self.discontinuous = family.value, dim, int(degree), bool(discontinuous)
        self.num_entity_dofs = ref_basis.num_entity_dofs(self.family_name, dim, self.degree, self.discontinuous)
        if self.family_name == "N1E" and self.degree != 1:
            raise NotImplementedError("basix stand-in: N1E beyond degree 1")
        if self.family_name in ("Lagrange", "N1E") or self.discontinuous or dim == 0:
            # dofs are numbered entity by entity (vertices, edges, faces, interior)
            run, out = 0, []
            for per in self.num_entity_dofs:
                out.append([])
                for n in per:
                    out[-1].append(list(range(run, run + n)))
                    run += n
            self.entity_dofs = out
        else:
            self.entity_dofs = [[list(e) for e in per] for per in ref_hdiv.entity_dofs(self.family_name, dim, self.degree)]
        if self.discontinuous or dim < 2:
            self.dof_transformations_are_identity = True
        elif self.family_name == "Lagrange":
            self.dof_transformations_are_identity = self.degree <= 2
        else:
            self.dof_transformations_are_identity = False

    def entity_transformations(self):
        if self.dim < 2:
            return {}
        if self.family_name == "Lagrange":
            n_edge = max(self.degree - 1, 0)
            out = {"interval": [np.fliplr(np.eye(n_edge))]}
            if self.dim == 3:
                n_face = max((self.degree - 1) * (self.degree - 2) // 2, 0)
                if n_face > 1:
                    raise NotImplementedError("basix stand-in: Lagrange face transformations beyond degree 3")
                out["triangle"] = [np.eye(n_face), np.eye(n_face)]
            return out
        if self.family_name == "N1E":  # one tangential moment per edge: reversing the edge flips its sign; no face dofs
            out = {"interval": [-np.eye(1)]}
            if self.dim == 3:
                out["triangle"] = [np.zeros((0, 0)), np.zeros((0, 0))]
            return out
        return ref_hdiv.entity_transformations(self.family_name, self.dim, self.degree)

    def tabulate(self, n_derivatives, points):
        assert n_derivatives == 1
        points = np.asarray(points, dtype=float).reshape(-1, self.dim)
        if self.family_name == "N1E":
            return _whitney(self.dim, points)
        return np.array(ref_basis.tabulate(self.family_name, self.dim, self.degree, points, self.discontinuous), dtype=float)

    def push_forward(self, values, jac, det_jac, inv_jac):
        if self.family_name == "N1E":  # covariant map: u = K^T u_ref with the (pseudo-)inverse Jacobian K[dim, 3]
            return np.einsum("qdi,qjd->qji", inv_jac, values)
        return ref_basis.push_forward(self.family_name, values, jac, det_jac, inv_jac)


def create_element(family, celltype, degree, lagrange_variant=None, discontinuous=False, **_):
    return _Element(family, celltype.value, degree, discontinuous)


def _whitney(dim, points):
    """Lowest-order Nedelec functions of the first kind on the reference simplex, written from their textbook definition
    (NOT from the oracle, which builds them in physical coordinates, nor from the product's tables): edge (a, b), a < b, in
    basix's edge order carries ``lambda_a grad lambda_b - lambda_b grad lambda_a``, whose tangential moment along
    ``v_b - v_a`` is one.  ``[1 + dim, nq, n_edges, dim]``."""
    from oracle.ref_basis import SUB

    grad = np.concatenate([-np.ones((1, dim)), np.eye(dim)], axis=0)  # grad lambda_i
    lam = np.concatenate([1.0 - points.sum(axis=1, keepdims=True), points], axis=1)  # [nq, dim + 1]
    edges = SUB[dim][1]
    tab = np.zeros((1 + dim, len(points), len(edges), dim))
    for j, (a, b) in enumerate(edges):
        tab[0, :, j, :] = lam[:, a, None] * grad[b] - lam[:, b, None] * grad[a]
        for d in range(dim):
            tab[1 + d, :, j, :] = grad[a, d] * grad[b] - grad[b, d] * grad[a]
    return tab
