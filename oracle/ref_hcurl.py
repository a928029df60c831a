"""ORACLE (test infrastructure only -- never imported by the product path).

Lowest-order first-kind Nedelec element ("N1E", the Whitney edge element) written directly in PHYSICAL coordinates,
independently of the product's reference tables and covariant map (nfempy_b200/basis/hcurl.py, csrc/nfa_hcurl.cu):

    phi_e = s_e (lam_a grad lam_b - lam_b grad lam_a),   e = (a, b) in the reference's local edge order
    (R/src/mesh/mesh.py:187-203: tetrahedron (2,3),(1,3),(1,2),(0,3),(0,2),(0,1); triangle (1,2),(0,2),(0,1)),

with the physical gradients of the barycentric coordinates -- which is what basix' covariant push-forward of the
reference Whitney functions gives (R/src/basis/finite_element.py:156-186).  ``s_e = -1`` for an edge the reference finds
mis-oriented (R/src/basis/permute_and_transform.py:8-19 in 2-D: local edge vertices differ from the edge cell's own
node_tags; :38-54 in 3-D: they do not ascend), since ``transformations["interval"][0] = [-1]`` for this element.

``nedelec2`` does the same for the second-kind element of degree 1 ("N2E").

PARITY: N1E of degree 1 is pinned against the reference's own pipeline (its ``FiniteElement`` covariant map and
``permute_and_transform`` edge signs over a textbook Whitney table: tests/golden/refcode_projector_n1e_*.npz, agreement 5e-16);
N2E has no counterpart run -- the reference's ``family_by_name("N2E")`` returns the N1E family (element_family.py:10).
basix's own N1E scaling stays unpinned (see oracle/ref_basis.py).
"""
import numpy as np

from .ref_basis import SUB


def whitney(space, cell_id, points):
    """``(x[nq, 3], det, phi[nq, n_edges, 3])`` of the oriented Whitney functions of one cell at reference ``points``."""
    mesh = space.mesh
    cell = mesh.cells[cell_id]
    dim = cell.dimension
    X = mesh.points[cell.node_tags]  # [dim+1, 3]
    J = (X[1:] - X[0]).T  # [3, dim]
    G = J.T @ J
    det = np.sqrt(np.linalg.det(G))
    Jinv = np.linalg.solve(G, J.T)  # [dim, 3]
    glam = np.concatenate([-Jinv.sum(axis=0, keepdims=True), Jinv], axis=0)  # [dim+1, 3] physical gradients
    points = np.asarray(points, dtype=float).reshape(-1, dim)
    lam = np.concatenate([1.0 - points.sum(axis=1, keepdims=True), points], axis=1)  # [nq, dim+1]
    x = lam @ X
    phi = np.zeros((len(points), len(SUB[dim][1]), 3))
    for e, (a, b) in enumerate(SUB[dim][1]):
        local = cell.node_tags[[a, b]]
        if dim == 2:
            own = mesh.cells[cell.sub_cells_ids[1][e]].node_tags
            oriented = np.array_equal(local, own)
        else:
            oriented = local[0] < local[1]
        s = 1.0 if oriented else -1.0
        phi[:, e, :] = s * (lam[:, [a]] * glam[b][None, :] - lam[:, [b]] * glam[a][None, :])
    return x, det, phi


def nedelec2(space, cell_id, points):
    """``(x, det, phi[nq, 2 n_edges, 3])`` of the second-kind Nedelec functions of degree 1 ("N2E": the full ``[P_1]^d``),
    written in physical coordinates.  Dual basis of the tangential moments of edge ``e = (a, b)`` (parameter s from a to
    b, tangent ``t = x_b - x_a``) against the orthonormal Legendre functions 1 and ``sqrt(3) (2 s - 1)`` (basix' legendre
    variant): with ``u = lam_a grad lam_b``, ``v = lam_b grad lam_a`` one has ``u . t = 1 - s``, ``v . t = -s`` on the edge
    and zero tangential trace on every other edge, hence

        phi_{e,0} = u - v            (moments 1, 0: the Whitney function)
        phi_{e,1} = -sqrt(3) (u + v) (moments 0, 1: a gradient, -sqrt(3) grad(lam_a lam_b)).

    Reversing the edge flips ``phi_{e,0}`` and leaves ``phi_{e,1}`` unchanged: the reference's transformation of a
    mis-oriented edge is ``diag(-1, +1)`` (R/src/basis/permute_and_transform.py:22-35, 110-140)."""
    mesh = space.mesh
    cell = mesh.cells[cell_id]
    dim = cell.dimension
    X = mesh.points[cell.node_tags]
    J = (X[1:] - X[0]).T
    G = J.T @ J
    det = np.sqrt(np.linalg.det(G))
    Jinv = np.linalg.solve(G, J.T)
    glam = np.concatenate([-Jinv.sum(axis=0, keepdims=True), Jinv], axis=0)
    points = np.asarray(points, dtype=float).reshape(-1, dim)
    lam = np.concatenate([1.0 - points.sum(axis=1, keepdims=True), points], axis=1)
    x = lam @ X
    phi = np.zeros((len(points), 2 * len(SUB[dim][1]), 3))
    for e, (a, b) in enumerate(SUB[dim][1]):
        local = cell.node_tags[[a, b]]
        if dim == 2:
            oriented = np.array_equal(local, mesh.cells[cell.sub_cells_ids[1][e]].node_tags)
        else:
            oriented = local[0] < local[1]
        u = lam[:, [a]] * glam[b][None, :]
        v = lam[:, [b]] * glam[a][None, :]
        phi[:, 2 * e, :] = (1.0 if oriented else -1.0) * (u - v)
        phi[:, 2 * e + 1, :] = -np.sqrt(3.0) * (u + v)
    return x, det, phi


def l2_projector_n2e(space, functions, iel, alpha):
    """The same form for a space whose single field is N2E of degree 1."""
    (name, sp), = space.discrete_spaces.items()
    dim = sp.dimension
    points, weights = space.quadrature[dim]
    x, det, phi = nedelec2(sp, sp.element_ids[iel], points)
    f = np.asarray(functions[name](x[:, 0], x[:, 1], x[:, 2]))[0]
    j_el = np.einsum("q,qad,qbd->ab", det * weights, phi, phi)
    r_el = -np.einsum("q,qad,dq->a", det * weights, phi, f)
    return r_el, j_el


def l2_projector(space, functions, iel, alpha):
    """L2ProjectorWeakForm.evaluate_form (R/src/weak_forms/l2_projector_weak_form.py:8-93) for a space whose single field
    is N1E of degree 1: ``J = int phi . phi``, ``r = -int f . phi``."""
    (name, sp), = space.discrete_spaces.items()
    dim = sp.dimension
    points, weights = space.quadrature[dim]
    x, det, phi = whitney(sp, sp.element_ids[iel], points)
    f = np.asarray(functions[name](x[:, 0], x[:, 1], x[:, 2]))[0]  # [3, nq]
    j_el = np.einsum("q,qad,qbd->ab", det * weights, phi, phi)
    r_el = -np.einsum("q,qad,dq->a", det * weights, phi, f)
    return r_el, j_el
