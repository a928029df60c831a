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


def test_md_two_flux_transport_step_matches_the_oracle():
    """BASELINE config 5 as stated -- the parabolic two-flux transport model on a mixed-dimensional fractured 3-D mesh: Sundus
    forms (time dependent, nonlinear reaction) in the box and on the fracture plane, their boundary forms, and the coupling
    of (mp, p) and (mc, c) across the fracture, in one stacked system."""
    from nfempy_b200.assembly import GPUAssembler
    from nfempy_b200.mesh import fractured_box_tet_mesh
    from nfempy_b200.spaces import ProductSpace
    from nfempy_b200.weak_forms import interface_triples
    from oracle import ref_assembly
    from oracle.ref_mesh import RefMesh
    from oracle.ref_spaces import RefProductSpace
    from tests.md_helpers import SUNDUS, SUNDUS_BC, md_sundus_forms, md_sundus_spaces

    mesh, interface = fractured_box_tet_mesh(2, jitter=0.15)
    spaces = md_sundus_spaces(mesh, ProductSpace)
    rmesh = RefMesh(3, mesh.points, mesh._blocks)
    rmesh.build_conformal_mesh()
    rspaces = md_sundus_spaces(rmesh, RefProductSpace)
    triples = interface_triples(spaces[0], spaces[1], interface, "mp")
    mp0 = spaces[0].discrete_spaces["mp"]
    outer = [i for i, cid in enumerate(mp0.bc_element_ids) if mesh.material[cid] != 50]
    n = spaces[1].dof_shift + spaces[1].n_dof
    a_n, a = 0.3 * np.cos(np.arange(n) * 0.37) + 0.05, 0.2 * np.sin(np.arange(n) * 0.11)
    t = 0.3
    asm = GPUAssembler(spaces, None, np.zeros(n))
    asm.form_to_input_list = md_sundus_forms(spaces, triples, outer, (a_n, a), t)
    got = asm.assemble()
    res_g, jac_g = np.zeros(n), ref_assembly.DokMatrix(n)
    for space, bc_items in ((rspaces[0], outer), (rspaces[1], range(len(rspaces[1].discrete_spaces["mp"].bc_element_ids)))):
        for i in range(len(space.discrete_spaces["mp"].element_ids)):
            ref_assembly.scatter_time_dependent_form(space, ref_forms.sundus_dual, SUNDUS, i, res_g, jac_g, a_n, a, t)
        for i in bc_items:
            ref_assembly.scatter_time_dependent_bc_form(space, ref_forms.sundus_dual_bc_dirichlet, SUNDUS_BC, i, res_g, jac_g, a_n, a, t)
    for pair in (("mp", "p"), ("mc", "c")):
        ref_md.scatter_coupling(rspaces, COUPLING, triples, res_g, jac_g, a_n, pair)
    compare_csr(got, jac_g.getValuesCSR() + (res_g,), 1e-12)
