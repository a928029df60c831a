"""Mixed-dimensional assembly on the GPU against the oracle: mixed Poisson in the fractured 3-D box, mixed Poisson on the
fracture plane, natural boundary conditions on the outer facets of both, and the Robin coupling over the interface
triples -- all in ONE stacked system (``GPUAssembler([space_c0, space_c1])``, ``dof_shift`` per space), as the driver
R/examples/md_elliptic_scalar/md_elliptic_scalar.py:205-342 builds it by hand."""
import numpy as np
import pytest

from oracle import ref_forms, ref_md
from tests.md_helpers import COUPLING, KAPPA_C0, KAPPA_C1, P_EXACT, RHS, md_problem
from tests.parity_helpers import compare_csr

pytestmark = pytest.mark.gpu


def _systems(k, alpha_seed=None):
    from nfempy_b200.assembly import GPUAssembler
    from nfempy_b200.weak_forms import CouplingWeakForm, LaplaceDualWeakForm, LaplaceDualWeakFormBCDirichlet

    mesh, spaces, ref_spaces, triples, outer = md_problem(2, k)
    n = spaces[1].dof_shift + spaces[1].n_dof
    alpha = np.zeros(n) if alpha_seed is None else np.random.default_rng(alpha_seed).standard_normal(n)
    f0, f1 = {"rhs": RHS, "kappa": KAPPA_C0}, {"rhs": RHS, "kappa": KAPPA_C1}
    bc = {"p": P_EXACT}
    forms = {}
    for tag, space, fun in (("c0", spaces[0], f0), ("c1", spaces[1], f1)):
        vol, bnd = LaplaceDualWeakForm(space), LaplaceDualWeakFormBCDirichlet(space)
        vol.functions, bnd.functions = fun, bc
        u = space.discrete_spaces["u"]
        forms["vol_" + tag] = ["form", range(len(u.elements)), vol, alpha]
        forms["bc_" + tag] = ["bc_form", outer if tag == "c0" else range(len(u.bc_elements)), bnd, alpha]
    cpl = CouplingWeakForm(spaces)
    cpl.functions = COUPLING
    forms["coupling"] = ["interface_form", triples, cpl, alpha]
    asm = GPUAssembler(spaces, None, np.zeros(n))
    asm.form_to_input_list = forms
    got = asm.assemble()
    u0, u1 = ref_spaces[0].discrete_spaces["u"], ref_spaces[1].discrete_spaces["u"]
    want = ref_md.assemble_md(
        ref_spaces,
        [[("form", ref_forms.laplace_dual, f0, len(u0.element_ids)), ("bc_form", ref_forms.laplace_dual_bc_dirichlet, bc, outer)],
         [("form", ref_forms.laplace_dual, f1, len(u1.element_ids)),
          ("bc_form", ref_forms.laplace_dual_bc_dirichlet, bc, len(u1.bc_element_ids))]],
        triples, COUPLING, alpha)
    return got, want, cpl, ref_spaces, triples


@pytest.mark.parametrize("k", [1, 2])
def test_md_elliptic_system_matches_the_oracle(k):
    got, want, *_ = _systems(k)
    compare_csr(got, want, 1e-12)


def test_md_coupling_residual_follows_the_state_vector():
    """Non-zero alpha: the coupling form's residual is J alpha_l (the volume forms of the reference discard alpha)."""
    got, want, *_ = _systems(1, alpha_seed=5)
    compare_csr(got, want, 1e-12)


def test_coupling_evaluate_form_single_cell():
    got, want, cpl, ref_spaces, triples = _systems(1)
    a = np.random.default_rng(1).standard_normal(cpl.n_loc())
    r, j = cpl.evaluate_form(*[int(v) for v in triples[2]], a)
    r_ref, j_ref = ref_md.coupling(ref_spaces, COUPLING, *[int(v) for v in triples[2]], a)
    assert np.abs(j - j_ref).max() <= 1e-12 * np.abs(j_ref).max() and np.abs(r - r_ref).max() <= 1e-12 * np.abs(r_ref).max()
