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


def _le_fixture_system(level, cap, with_bc=True):
    import os

    from nfempy_b200.spaces import ProductSpace
    from nfempy_b200.weak_forms import LEPrimalWeakForm, LEPrimalWeakFormBCDirichlet

    old = os.environ.get("NFA_P2TET_CAP")
    if cap is not None:
        os.environ["NFA_P2TET_CAP"] = str(cap)
    try:
        space = ProductSpace({"u": (3, 3, "Lagrange", 2, fixture_mesh(3, level))})
        space.build_structures({"u": [1]}, {"u": [2, 3, 4, 5, 6, 7]})
        vol, bc = LEPrimalWeakForm(space), LEPrimalWeakFormBCDirichlet(space)
        vol.functions, bc.functions = LE_FUNCTIONS, LE_BC_FUNCTIONS
        u = space.discrete_spaces["u"]
        forms = {"volume": ["form", range(len(u.elements)), vol, np.zeros(space.n_dof)]}
        if with_bc:
            forms["bc"] = ["bc_form", range(len(u.bc_elements)), bc, np.zeros(space.n_dof)]
        return space, _assemble(space, forms)
    finally:
        if old is None:
            os.environ.pop("NFA_P2TET_CAP", None)
        else:
            os.environ["NFA_P2TET_CAP"] = old


@pytest.mark.parametrize("cap", [64, None])
def test_elasticity_p2_on_the_level_1_fixture(cap):
    """1 325 tets of the reference's gmsh fixture against the cell-by-cell oracle's committed CSR.  ``cap = 64``: chunks of
    64 blocks, so every CTA of the row-item kernel walks many chunks and hands its tiles over many times -- the multi-chunk
    path meets the ORACLE here, not another CUDA kernel."""
    _, got = _le_fixture_system(1, cap)
    z = np.load(GOLDEN + "/le_p2_ex3_3d_l1.npz")
    compare_csr(got, (z["indptr"], z["indices"], z["data"], z["rhs"]), 1e-12)


def test_elasticity_p2_on_the_level_2_fixture():
    """4 261 tets: pattern, right-hand side, diagonal and four matrix-vector probes of the cell-by-cell oracle (committed
    fingerprint), and EVERY value of the volume form against oracle/strong_cpu.py assembled here (itself pinned to the
    cell-by-cell oracle in tests/test_cpu_strong_baseline.py)."""
    import scipy.sparse as sp

    from oracle import strong_cpu

    space, (indptr, indices, data, rhs) = _le_fixture_system(2, 64)
    z = np.load(GOLDEN + "/le_p2_ex3_3d_l2_probe.npz")
    assert np.array_equal(indptr, z["indptr"]) and np.array_equal(indices, z["indices"])
    n = space.n_dof
    K = sp.csr_matrix((data, indices, indptr), shape=(n, n))
    V = np.random.default_rng(int(z["seed"])).standard_normal((n, 4))
    scale = z["abs_row_sum"][:, None] * np.abs(V).max()
    assert (np.abs(K @ V - z["KV"]) <= 1e-12 * scale).all()
    assert (np.abs(K.diagonal() - z["diag"]) <= 1e-12 * np.abs(z["diag"])).all()
    assert np.abs(rhs - z["rhs"]).max() <= 1e-12 * np.abs(z["rhs"]).max()
    # all values, volume form only
    space, (indptr, indices, data, _) = _le_fixture_system(2, 64, with_bc=False)
    u = space.discrete_spaces["u"]
    ref = strong_cpu.assemble_elasticity(u.mesh.points, np.asarray(u.cell_nodes(), dtype=np.int64),
                                         np.asarray(space.destination_table(), dtype=np.int64), n, 1.0, 1.0)
    compare_csr((indptr, indices, data, np.zeros(1)), ref + (np.zeros(1),), 1e-12)
