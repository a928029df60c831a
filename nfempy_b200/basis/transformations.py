"""``entity_transformations()`` of the restated elements: the matrices the reference applies to the basis
functions of a mis-oriented edge / face (R/src/basis/permute_and_transform.py:22-35, 110-140:
``phi[dofs] = T @ phi[dofs]`` with ``T = transformations["interval"][0]`` or
``reflect^n_ref . rotate^n_rot`` from ``rotate, reflect = transformations["triangle"]``).

Derivation: a sub-entity's dofs are moments against (or evaluations at) a nodal Lagrange family on the reference
sub-entity.  Re-parametrising the sub-entity by a symmetry ``g`` permutes that nodal family
(``q_i o g = q_{pi(i)}``) and, for normal moments of H(div) elements, flips the sign when ``g`` reverses the
orientation.  The matrices below are exactly those signed permutations:

* interval reflection   ``s -> 1 - s``;
* triangle rotation     vertices ``(0, 1, 2) -> (1, 2, 0)``;
* triangle reflection   vertices ``(0, 1, 2) -> (0, 2, 1)``.

Whether the "rotation" is this cyclic shift or its inverse is a convention of basix that cannot be read off the
reference; it is fixed here as the one under which the reference's own orientation table
(R/src/basis/permute_and_transform.py:83-100) yields an H(div)-conforming global basis -- checked by
tests/test_cpu_hdiv.py::test_global_basis_is_hdiv_conforming.
"""
import numpy as np

from .reference_element import REF_VERTICES, SUB_ENTITIES

ROTATION_IS_FORWARD_SHIFT = True


def _lagrange_nodes(dim, degree):
    """Nodes of the nodal family of degree <= 3 on the reference simplex, basix order."""
    verts = REF_VERTICES[dim] if dim > 0 else np.zeros((1, 0))
    if degree == 0:
        return verts.mean(axis=0, keepdims=True)
    nodes = [v for v in verts]
    if degree == 2:
        for a, b in SUB_ENTITIES[dim][1]:
            nodes.append(0.5 * (verts[a] + verts[b]))
    elif degree == 3:  # the equispaced cubic lattice of basis/hdiv.py (BDM_3 facet moments)
        for a, b in SUB_ENTITIES[dim][1]:
            for t in (1.0 / 3.0, 2.0 / 3.0):
                nodes.append((1.0 - t) * verts[a] + t * verts[b])
        if dim == 2:
            nodes.append(verts.mean(axis=0))
    elif degree != 1:
        raise NotImplementedError("transformation of sub-entity dof families above degree 3")
    return np.array(nodes)


def _node_permutation(dim, degree, vertex_image):
    """``P[i, j] = 1`` when the affine map sending reference vertex v to ``vertex_image[v]`` takes node j to node i."""
    verts = REF_VERTICES[dim]
    nodes = _lagrange_nodes(dim, degree)
    lam = np.concatenate([1.0 - nodes.sum(axis=1, keepdims=True), nodes], axis=1)  # barycentric coordinates
    mapped = lam @ verts[np.array(vertex_image)]
    P = np.zeros((len(nodes), len(nodes)))
    for j, x in enumerate(mapped):
        i = int(np.argmin(np.abs(nodes - x).sum(axis=1)))
        if np.abs(nodes[i] - x).max() > 1e-12:
            raise AssertionError("symmetry does not permute the nodal family")
        P[i, j] = 1.0
    return P


def facet_family_degree(el):
    """Degree of the nodal family carried by the element's facets/edges that needs transforming."""
    if el.family == "RT":
        return el.degree - 1
    if el.family == "BDM":
        return el.degree
    if el.family == "Lagrange":
        return el.degree - 2  # interior edge nodes of P_k form a degree k-2 family (k = 3: 2 nodes = degree 1)
    raise NotImplementedError(el.family)


def entity_transformations(el):
    """``{"interval": [T], "triangle": [rotate, reflect]}`` like basix, for the entity types the cell has."""
    out = {}
    if el.dim < 2:
        return out
    if el.family in ("N1E", "N2E"):
        # N1E_1: one tangential moment per edge, reversing the edge flips its sign; N2E_1: moments against the Legendre
        # functions 1 and sqrt(3)(2s - 1) -- the even one flips with the tangent, the odd one does not; no face dofs
        out["interval"] = [-np.ones((1, 1))] if el.family == "N1E" else [np.diag([-1.0, 1.0])]
        if el.dim == 3:
            out["triangle"] = [np.zeros((0, 0)), np.zeros((0, 0))]
        return out
    hdiv = el.family in ("RT", "BDM")
    if el.dim == 2 or not hdiv:
        # dofs attached to edges: H(div) facets of a triangle, or interior edge nodes of Lagrange P3
        deg = facet_family_degree(el) if hdiv else el.degree - 2
        n = len(el.entity_dofs[1][0])
        if n:
            P = _node_permutation(1, deg if hdiv else 1 if n == 2 else 0, [1, 0]) if n > 1 else np.ones((1, 1))
            out["interval"] = [(-1.0 if hdiv else 1.0) * P]
        else:
            out["interval"] = [np.zeros((0, 0))]
    else:
        out["interval"] = [np.zeros((0, 0))]
    if el.dim == 3:
        n = len(el.entity_dofs[2][0])
        if n == 0:
            out["triangle"] = [np.zeros((0, 0)), np.zeros((0, 0))]
        elif hdiv:
            deg = facet_family_degree(el)
            shift = [1, 2, 0] if ROTATION_IS_FORWARD_SHIFT else [2, 0, 1]
            rot = _node_permutation(2, deg, shift) if n > 1 else np.ones((1, 1))
            ref = -(_node_permutation(2, deg, [0, 2, 1]) if n > 1 else np.ones((1, 1)))
            out["triangle"] = [rot, ref]
        else:
            out["triangle"] = [np.eye(n), np.eye(n)]  # P3 face bubble: one node at the centroid
    return out
