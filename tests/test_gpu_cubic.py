"""GPU tests of CONTINUOUS cubic Lagrange fields, the first Lagrange degree whose dof transformations are not the
identity: the reference exchanges the two basis functions of a mis-oriented edge (R/src/basis/permute_and_transform.py:22-35,
110-140), the product scatters the untransformed local matrix through exchanged destinations
(``DiscreteSpace.destination_table``).  Volume forms only: the reference's Dirichlet forms pick ONE dof per sub-entity
(R/src/weak_forms/laplace_primal_weak_form.py:120-133) and do not run at degree 3 themselves."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spl

from nfempy_b200.mesh import box_tet_mesh, square_tri_mesh
from oracle import ref_forms
from tests.parity_helpers import Case, compare_csr, laplace_case

pytestmark = pytest.mark.gpu
RTOL = 1e-12


@pytest.mark.parametrize("dim,fused", [(2, True), (2, False), (3, True)])
def test_laplace_cubic_continuous_matches_the_oracle(dim, fused):
    case = laplace_case(n=3 if dim == 2 else 2, dim=dim, k=3, with_bc=False)
    compare_csr(case.run_gpu(allow_fused=fused), case.run_oracle(), RTOL)


def _cubic(x, y, z):
    return np.array([1.0 + x - 2.0 * y + 0.5 * z + x * y * (x - y) + z * z * z - 0.75 * x * x * z])


@pytest.mark.parametrize("dim", [2, 3])
def test_projection_of_a_cubic_onto_continuous_cubic_lagrange_is_exact(dim):
    """The reference's projector property test (R/tests/test_projectors/test_projectors.py:86-216) at degree 3: only a
    conforming space reproduces the polynomial."""
    from nfempy_b200.weak_forms import L2ProjectorWeakForm

    mesh = box_tet_mesh(2, jitter=0.2, shuffle_local_nodes=True) if dim == 3 else square_tri_mesh(3, jitter=0.2)
    functions = {"u": _cubic}

    def gpu_forms(space, a):
        wf = L2ProjectorWeakForm(space)
        wf.functions = functions
        return {"proj": ["form", range(len(space.discrete_spaces["u"].elements)), wf, a]}

    case = Case(mesh, lambda g: {"u": (dim, 1, "Lagrange", 3, g)}, {"u": [1]}, {}, gpu_forms, None)
    indptr, indices, data, rhs = case.run_gpu()
    alpha = spl.spsolve(sp.csr_matrix((data, indices, indptr)).tocsc(), -rhs)
    space = case.ref_space()
    us = space.discrete_spaces["u"]
    points, _ = space.quadrature[dim]
    worst = 0.0
    for i, cid in enumerate(us.element_ids):
        el = ref_forms.RefElement(us, cid)
        x, jac, det, inv = el.evaluate_mapping(points)
        tab = el.evaluate_basis(points, jac, det, inv)
        uh = tab[0, :, :, 0] @ alpha[space.destination_indexes(i)]
        worst = max(worst, np.abs(uh - _cubic(x[:, 0], x[:, 1], x[:, 2])[0]).max())
    assert worst <= 1e-11
