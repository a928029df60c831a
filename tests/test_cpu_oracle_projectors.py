"""The reference's own test of this path (R/tests/test_projectors/test_projectors.py:86-216) restated against the ORACLE:
``l2_projector`` of a function that lies in the discrete space, then ``l2_error`` close to zero -- with the reference's
``s_functions`` / ``v_functions`` and its order rule (the H(div) / H(curl) loops raise ``k_order`` by one before the first
family and keep it raised), for every family and order the oracle restates.  It is a property any valid basis of the
space satisfies, so it pins the oracle's spaces (numbering, Piola maps, orientation), not basix' choice of basis."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spl

from nfempy_b200.mesh import box_tet_mesh, square_tri_mesh
from oracle import ref_assembly, ref_forms, ref_hcurl, ref_postprocess
from oracle.ref_mesh import RefMesh
from oracle.ref_spaces import RefProductSpace

S_FUNCTIONS = [
    lambda x, y, z: np.array([x + y]),
    lambda x, y, z: np.array([x * (1.0 - x) + y * (1.0 - y)]),
    lambda x, y, z: np.array([(x**2) * (1.0 - x) + (y**2) * (1.0 - y)]),
]
V_FUNCTIONS = [
    lambda x, y, z: np.array([[y, -x, -z]]),
    lambda x, y, z: np.array([[(1 - y) * y**1, -(1 - x) * x**1, -(1 - z) * z**1]]),
    lambda x, y, z: np.array([[(1 - y) * y**2, -(1 - x) * x**2, -(1 - z) * z**2]]),
]


def _space(dim, family, k, n_comp=1, discontinuous=False):
    mesh = box_tet_mesh(1, jitter=0.2, shuffle_local_nodes=True) if dim == 3 else square_tri_mesh(2, jitter=0.2)
    r = RefMesh(dim, mesh.points, mesh._blocks)
    r.build_conformal_mesh()
    space = RefProductSpace({"f": (dim, n_comp, family, k, r)})
    space.make_subspaces_discontinuous({"f": discontinuous})
    space.build_structures({"f": [1]}, {})
    return space


def _project(space, form, functions):
    n = len(space.discrete_spaces["f"].element_ids)
    indptr, indices, data, res = ref_assembly.assemble(space, [("form", form, functions, n)])
    return spl.spsolve(sp.csr_matrix((data, indices, indptr)).tocsc(), -res)


@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("discontinuous", [True, False])
def test_scalar_h1_projector(k, dim, discontinuous):
    space = _space(dim, "Lagrange", k, discontinuous=discontinuous)
    functions = {"f": S_FUNCTIONS[k - 1]}
    alpha = _project(space, ref_forms.l2_projector, functions)
    assert ref_postprocess.l2_error(dim, space, functions, alpha)[0] <= 1e-12


@pytest.mark.parametrize("family,k_function,k_space", [("RT", 1, 2), ("RT", 2, 3), ("BDM", 1, 2), ("BDM", 2, 3), ("BDM", 1, 1)])
@pytest.mark.parametrize("dim", [2, 3])
def test_vector_hdiv_projector(family, k_function, k_space, dim):
    """The reference projects ``v_functions[k - 1]`` onto RT_{k+1} and -- because its loop keeps the raised order -- onto
    BDM_{k+1}; BDM_1 with the linear field is added (it is in the space)."""
    space = _space(dim, family, k_space)
    functions = {"f": V_FUNCTIONS[k_function - 1]}
    alpha = _project(space, ref_forms.l2_projector, functions)
    assert ref_postprocess.l2_error(dim, space, functions, alpha)[0] <= 1e-11


@pytest.mark.parametrize("family,dim", [("N1E", 2), ("N2E", 2), ("N2E", 3)])
def test_vector_hcurl_projector(family, dim):
    """Degree 1, the order the oracle restates: ``v_functions[0] = (y, -x, -z)`` is a rotation field in 2-D (in N1E_1) and a
    linear field (in N2E_1) in both dimensions."""
    space = _space(dim, family, 1)
    functions = {"f": V_FUNCTIONS[0]}
    form, basis = (ref_hcurl.l2_projector, ref_hcurl.whitney) if family == "N1E" else (ref_hcurl.l2_projector_n2e, ref_hcurl.nedelec2)
    alpha = _project(space, form, functions)
    fs = space.discrete_spaces["f"]
    points, weights = space.quadrature[dim]
    total = 0.0
    for i in range(len(fs.element_ids)):
        x, det, phi = basis(fs, fs.element_ids[i], points)
        f_h = np.einsum("j,qjc->qc", alpha[space.destination_indexes(i)], phi)
        f_e = np.asarray(functions["f"](x[:, 0], x[:, 1], x[:, 2]))[0].T.copy()
        if dim == 2:
            f_e[:, 2] = 0.0
        total += float(np.sum(det * weights * np.sum((f_e - f_h) ** 2, axis=1)))
    assert np.sqrt(total) <= 1e-12
