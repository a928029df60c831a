"""H(curl) reference element N1E_1 (nfempy_b200/basis/hcurl.py): the Whitney functions, their tangential moments, and
tangential continuity of the covariantly mapped functions across physical cells once the reference's edge-orientation
rule (ascending global vertex ids, R/src/basis/permute_and_transform.py:38-54) is applied with
``entity_transformations()["interval"][0] = [[-1]]``."""
import numpy as np
import pytest

from nfempy_b200.basis.hcurl import covariant_piola
from nfempy_b200.basis.reference_element import REF_VERTICES, SUB_ENTITIES, create_element
from nfempy_b200.mesh import box_tet_mesh, square_tri_mesh


@pytest.mark.parametrize("dim", [2, 3])
def test_basis_is_the_whitney_family_with_unit_tangential_moments(dim):
    el = create_element("N1E", dim, 1)
    edges = SUB_ENTITIES[dim][1]
    assert el.n_phi == len(edges) and el.value_size == dim and not el.dof_transformations_are_identity
    rng = np.random.default_rng(3)
    pts = rng.random((20, dim)) / dim
    tab = el.tabulate(1, pts)[0]  # [nq, n_phi, dim]
    lam = np.concatenate([1.0 - pts.sum(axis=1, keepdims=True), pts], axis=1)
    glam = np.concatenate([-np.ones((1, dim)), np.eye(dim)], axis=0)
    for i, (a, b) in enumerate(edges):
        whitney = lam[:, [a]] * glam[b][None, :] - lam[:, [b]] * glam[a][None, :]
        assert np.abs(tab[:, i, :] - whitney).max() < 1e-14
    verts = REF_VERTICES[dim]
    s = (np.arange(8) + 0.5) / 8.0  # midpoint rule is exact for the (affine) tangential component
    for j, (a, b) in enumerate(edges):
        x = verts[a][None, :] + s[:, None] * (verts[b] - verts[a])[None, :]
        moments = np.einsum("qic,c->i", el.tabulate(0, x)[0], verts[b] - verts[a]) / len(s)
        assert np.abs(moments - np.eye(len(edges))[j]).max() < 1e-14


@pytest.mark.parametrize("dim", [2, 3])
def test_tangential_continuity_across_cells(dim):
    mesh = box_tet_mesh(2, jitter=0.2, shuffle_local_nodes=True) if dim == 3 else square_tri_mesh(3, jitter=0.2)
    el = create_element("N1E", dim, 1)
    flip = el.entity_transformations()["interval"][0][0, 0]
    cells = mesh.vol_nodes if dim == 3 else mesh.face_nodes
    edges = SUB_ENTITIES[dim][1]
    verts = REF_VERTICES[dim]
    seen = {}
    checked = 0
    for c, nodes in enumerate(cells):
        X = mesh.points[nodes]
        J = (X[1:] - X[0]).T  # [3, dim]
        for i, (a, b) in enumerate(edges):
            ga, gb = int(nodes[a]), int(nodes[b])
            sign = 1.0 if ga < gb else flip  # the reference transforms the functions of mis-oriented edges
            xi = 0.3 * verts[a] + 0.7 * verts[b] if ga < gb else 0.7 * verts[a] + 0.3 * verts[b]  # same physical point
            phi = covariant_piola(el.tabulate(0, xi[None, :])[0][0], J)[i] * sign  # [3]
            lo, hi = (ga, gb) if ga < gb else (gb, ga)
            tangent = mesh.points[hi] - mesh.points[lo]
            t_comp = float(phi @ tangent)
            key = (lo, hi)
            if key in seen:
                assert abs(t_comp - seen[key]) < 1e-12, (key, t_comp, seen[key])
                checked += 1
            else:
                seen[key] = t_comp
            assert abs(t_comp - 1.0) < 1e-12  # unit tangential moment along the globally oriented edge
    assert checked > 10
