"""The CUDA path against the COMMITTED golden fixtures (tests/golden/, generator tools/make_golden.py): CSR pattern
bit-exact, values and right-hand side within the stated tolerance, on the reference's own gmsh meshes rebuilt from the
committed mesh arrays (no /root/reference at run time)."""
import numpy as np
import pytest

from tests.golden_cases import (GOLDEN, LAPLACE_BC_FUNCTIONS, LAPLACE_FUNCTIONS, LE_BC_FUNCTIONS, LE_FUNCTIONS,
                                fixture_mesh, multi_cases)
from tests.parity_helpers import compare_csr

pytestmark = pytest.mark.gpu


def _assemble(space, forms):
    from nfempy_b200.assembly import GPUAssembler

    asm = GPUAssembler(space, None, np.zeros(space.n_dof))
    asm.form_to_input_list = forms
    return asm.assemble()


def test_elasticity_p2_on_the_3d_fixture():
    from nfempy_b200.spaces import ProductSpace
    from nfempy_b200.weak_forms import LEPrimalWeakForm, LEPrimalWeakFormBCDirichlet

    space = ProductSpace({"u": (3, 3, "Lagrange", 2, fixture_mesh(3))})
    space.build_structures({"u": [1]}, {"u": [2, 3, 4, 5, 6, 7]})
    vol, bc = LEPrimalWeakForm(space), LEPrimalWeakFormBCDirichlet(space)
    vol.functions, bc.functions = LE_FUNCTIONS, LE_BC_FUNCTIONS
    u = space.discrete_spaces["u"]
    got = _assemble(space, {"volume": ["form", range(len(u.elements)), vol, np.zeros(space.n_dof)],
                            "bc": ["bc_form", range(len(u.bc_elements)), bc, np.zeros(space.n_dof)]})
    z = np.load(GOLDEN + "/le_p2_ex3_3d_l0.npz")
    compare_csr(got, (z["indptr"], z["indices"], z["data"], z["rhs"]), 1e-12)


def test_laplace_p1_on_the_2d_fixture():
    from nfempy_b200.spaces import ProductSpace
    from nfempy_b200.weak_forms import LaplacePrimalWeakForm, LaplacePrimalWeakFormBCDirichlet

    z = np.load(GOLDEN + "/laplace_p1_ex3_2d_l0.npz")
    space = ProductSpace({"p": (2, 1, "Lagrange", 1, fixture_mesh(2))})
    space.build_structures({"p": [int(t) for t in z["tags"]]}, {"p": [int(t) for t in z["btags"]]})
    vol, bc = LaplacePrimalWeakForm(space), LaplacePrimalWeakFormBCDirichlet(space)
    vol.functions, bc.functions = LAPLACE_FUNCTIONS, LAPLACE_BC_FUNCTIONS
    p = space.discrete_spaces["p"]
    got = _assemble(space, {"volume": ["form", range(len(p.elements)), vol, np.zeros(space.n_dof)],
                            "bc": ["bc_form", range(len(p.bc_elements)), bc, np.zeros(space.n_dof)]})
    compare_csr(got, (z["indptr"], z["indices"], z["data"], z["rhs"]), 1e-12)


@pytest.mark.parametrize("name", ["lce_primal", "taylor_hood", "mixed_poisson_bdm3", "lce_dual_neumann"])
def test_later_form_families_on_the_2d_fixture(name):
    case, tol = multi_cases()[name]
    z = np.load(GOLDEN + "/multi_ex3_2d_l0.npz")
    compare_csr(case.run_gpu(), (z[name + "_indptr"], z[name + "_indices"], z[name + "_data"], z[name + "_rhs"]), tol)
