"""CPU tests pinning the oracle: known-answer and invariant tests of the restated weak forms (the reference's own
tests hold no golden vectors for this path), plus the committed golden fixtures."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import ref_assembly, ref_forms
from oracle.ref_mesh import RefMesh
from oracle.ref_spaces import RefProductSpace
from tests.golden_cases import (GOLDEN, LAPLACE_BC_FUNCTIONS, LAPLACE_FUNCTIONS, LE_BC_FUNCTIONS, LE_FUNCTIONS,
                                load_mesh_blocks)
from tests.parity_helpers import elasticity_case, laplace_case


def _matrix(case):
    indptr, indices, data, rhs = case.run_oracle()
    return sp.csr_matrix((data, indices, indptr)), rhs


def test_elasticity_rigid_body_modes_and_symmetry():
    case = elasticity_case(n=1, with_bc=False)
    A, _ = _matrix(case)
    scale = abs(A).max()
    assert abs(A - A.T).max() <= 1e-14 * scale
    space = case.ref_space()
    # dof (node, c) = 3 node + c; node coordinates: vertices then edge midpoints
    mesh = space.discrete_spaces["u"].mesh
    n_nodes = space.n_dof // 3
    X = np.zeros((n_nodes, 3))
    us = space.discrete_spaces["u"]
    for i, cid in enumerate(us.element_ids):
        cell = mesh.cells[cid]
        dest = space.destination_indexes(i)[::3] // 3
        pts = mesh.points[cell.node_tags]
        X[dest[:4]] = pts
        for e, (a, b) in enumerate([(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)]):
            X[dest[4 + e]] = 0.5 * (pts[a] + pts[b])
    modes = []
    for c in range(3):
        t = np.zeros((n_nodes, 3))
        t[:, c] = 1.0
        modes.append(t.ravel())
    for axis in range(3):  # infinitesimal rotations
        w = np.zeros(3)
        w[axis] = 1.0
        modes.append(np.cross(w, X).ravel())
    for m in modes:
        assert np.abs(A @ m).max() <= 1e-12 * scale * max(1.0, np.abs(m).max())


def test_elasticity_constant_strain_patch():
    """u = E x is reproduced exactly by P2: K u equals the boundary-traction load, interior rows vanish for constant
    stress (div sigma = 0)."""
    case = elasticity_case(n=2, with_bc=False)
    A, _ = _matrix(case)
    space = case.ref_space()
    mesh = space.discrete_spaces["u"].mesh
    us = space.discrete_spaces["u"]
    n_nodes = space.n_dof // 3
    X = np.zeros((n_nodes, 3))
    for i, cid in enumerate(us.element_ids):
        cell = mesh.cells[cid]
        dest = space.destination_indexes(i)[::3] // 3
        pts = mesh.points[cell.node_tags]
        X[dest[:4]] = pts
        for e, (a, b) in enumerate([(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)]):
            X[dest[4 + e]] = 0.5 * (pts[a] + pts[b])
    E = np.array([[0.3, 0.1, -0.2], [0.05, -0.4, 0.2], [0.0, 0.3, 0.1]])
    u = (X @ E.T).ravel()
    r = (A @ u).reshape(n_nodes, 3)
    interior = np.all((X > 1e-9) & (X < 1 - 1e-9), axis=1)
    assert interior.any()
    assert np.abs(r[interior]).max() <= 1e-12 * abs(A).max()


def test_laplace_constants_in_null_space_and_load_sum():
    case = laplace_case(n=4, dim=2, k=2, with_bc=False)
    A, rhs = _matrix(case)
    assert np.abs(A @ np.ones(A.shape[0])).max() <= 1e-13 * abs(A).max()
    # sum_i r_i = -int f over the unit square (partition of unity), f = 2 pi^2 sin sin + 0
    assert abs(rhs.sum() + 8.0) <= 2e-2  # int_0^1 int_0^1 2 pi^2 sin(pi x) sin(pi y) = 8, 9-point rule on 32 cells


def test_dirichlet_penalty_block_is_scaled_facet_mass():
    case = laplace_case(n=2, dim=2, k=1, with_bc=True)
    space = case.ref_space()
    r, j = ref_forms.laplace_primal_bc_dirichlet(space, {"p": lambda x, y, z: np.array([1.0])}, 0, np.zeros(2))
    h = 0.5
    assert np.allclose(j, 1e12 * h / 6.0 * np.array([[2, 1], [1, 2]]), rtol=1e-13)
    assert np.allclose(r, -1e12 * h / 2.0, rtol=1e-13)


def test_oracle_reproduces_golden_elasticity_fixture():
    points, blocks = load_mesh_blocks("mesh_ex3_3d_l0.npz")
    mesh = RefMesh(3, points, blocks)
    mesh.build_conformal_mesh()
    space = RefProductSpace({"u": (3, 3, "Lagrange", 2, mesh)})
    space.build_structures({"u": [1]}, {"u": [2, 3, 4, 5, 6, 7]})
    z = np.load(GOLDEN + "/le_p2_ex3_3d_l0.npz")
    us = space.discrete_spaces["u"]
    # the whole fixture (298 tets + every boundary facet) is re-assembled cell by cell and must reproduce the committed
    # CSR bit for bit: pattern, values and right-hand side
    forms = [
        ("form", ref_forms.le_primal, LE_FUNCTIONS, len(us.element_ids)),
        ("bc_form", ref_forms.le_primal_bc_dirichlet, LE_BC_FUNCTIONS, len(us.bc_element_ids)),
    ]
    indptr, indices, data, rhs = ref_assembly.assemble(space, forms)
    assert z["indptr"].dtype == np.int32 and len(z["rhs"]) == space.n_dof
    assert np.array_equal(indptr, z["indptr"]) and np.array_equal(indices, z["indices"])
    assert np.array_equal(data, z["data"]) and np.array_equal(rhs, z["rhs"])
    assert len(us.bc_element_ids) > 0


def test_oracle_reproduces_golden_laplace_fixture():
    points, blocks = load_mesh_blocks("mesh_ex3_2d_l0.npz")
    z = np.load(GOLDEN + "/laplace_p1_ex3_2d_l0.npz")
    mesh = RefMesh(2, points, blocks)
    mesh.build_conformal_mesh()
    space = RefProductSpace({"p": (2, 1, "Lagrange", 1, mesh)})
    space.build_structures({"p": [int(t) for t in z["tags"]]}, {"p": [int(t) for t in z["btags"]]})
    ps = space.discrete_spaces["p"]
    forms = [
        ("form", ref_forms.laplace_primal, LAPLACE_FUNCTIONS, len(ps.element_ids)),
        ("bc_form", ref_forms.laplace_primal_bc_dirichlet, LAPLACE_BC_FUNCTIONS, len(ps.bc_element_ids)),
    ]
    indptr, indices, data, rhs = ref_assembly.assemble(space, forms)
    assert np.array_equal(indptr, z["indptr"]) and np.array_equal(indices, z["indices"])
    assert np.array_equal(data, z["data"]) and np.array_equal(rhs, z["rhs"])


def test_oracle_reproduces_golden_multi_form_fixture():
    """multi_ex3_2d_l0.npz (tools/make_golden.py multi): primal Cosserat, Taylor-Hood, BDM_3 mixed Poisson and the mixed
    Cosserat form with both stress boundary conditions on the reference's 2-D gmsh mesh."""
    from tests.golden_cases import multi_cases

    z = np.load(GOLDEN + "/multi_ex3_2d_l0.npz")
    for name, (case, _) in multi_cases().items():
        if name == "mixed_poisson_bdm3":
            continue  # 20 s of oracle time: covered by the GPU comparison against the same fixture
        indptr, indices, data, rhs = case.run_oracle()
        assert np.array_equal(indptr, z[name + "_indptr"]) and np.array_equal(indices, z[name + "_indices"]), name
        assert np.abs(data - z[name + "_data"]).max() <= 1e-13 * np.abs(data).max(), name
        assert np.abs(rhs - z[name + "_rhs"]).max() <= 1e-13 * max(np.abs(rhs).max(), 1e-300), name


def test_sundus_oracle_tangent_is_the_derivative_of_its_residual():
    """The nonlinear two-flux form (reaction 1 + eta c^2): the oracle's complex-step Jacobian against central differences
    of its own residual, its structure (no (mp, mc) / (p, c) couplings, the reaction tangent only on (c, c)) and the
    linearity in everything but c -- what the CUDA kernels' hand-derived tangent is compared with."""
    from oracle import ref_forms
    from tests.parity_helpers import SundusCase

    case = SundusCase(dim=3, k=1, n=1)
    space = case.ref_space()
    n_loc = len(space.destination_indexes(0))
    rng = np.random.default_rng(11)
    a, a_n = rng.standard_normal(n_loc), rng.standard_normal(n_loc)
    r0, J = ref_forms.sundus_dual(space, case.functions, 2, a, a_n, case.t)
    h = 1e-6
    for j in range(n_loc):
        e = np.zeros(n_loc)
        e[j] = h
        rp, _ = ref_forms.sundus_dual(space, case.functions, 2, a + e, a_n, case.t)
        rm, _ = ref_forms.sundus_dual(space, case.functions, 2, a - e, a_n, case.t)
        assert np.abs((rp - rm) / (2 * h) - J[:, j]).max() <= 1e-7 * max(1.0, np.abs(J).max())
    n_u = (n_loc - 2) // 2
    mp, mc, p, c = slice(0, n_u), slice(n_u, 2 * n_u), 2 * n_u, 2 * n_u + 1
    assert not J[mp, mc].any() and not J[mc, mp].any() and J[p, c] == 0.0 and J[c, p] == 0.0
    assert not J[mp, c].any() and not J[mc, p].any() and not J[p, mc].any() and not J[c, mp].any()
    assert np.allclose(J[mp, p], -J[p, mp]) and np.allclose(J[mc, c], -J[c, mc]) and np.allclose(J[mp, mp], J[mp, mp].T)
    # only the (c, c) entry depends on the iterate: J(a) - J(a') is zero elsewhere
    a2 = a.copy()
    a2[c] += 0.7
    _, J2 = ref_forms.sundus_dual(space, case.functions, 2, a2, a_n, case.t)
    D = J2 - J
    assert abs(D[c, c]) > 1e-6
    D[c, c] = 0.0
    assert np.abs(D).max() <= 1e-13 * np.abs(J).max()


@pytest.mark.parametrize("dim", [2, 3])
def test_oracle_lowest_order_mixed_poisson_reproduces_a_constant_flux(dim):
    """Known answer for RT_1 x P_0 (element scaling, Piola map, facet orientation and the sign of the natural boundary
    term all enter): with f = 0, constant kappa and linear boundary data p_D the discrete solution is the exact constant
    flux u = -kappa grad p_D, and the pressure dofs are the cell averages of p_D, i.e. its centroid values."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl

    from oracle import ref_assembly, ref_forms
    from tests.parity_helpers import mixed_poisson_case

    case = mixed_poisson_case(dim=dim, k=1, n=2)
    space = case.ref_space()
    grad = np.array([1.0, 2.0, -1.0 if dim == 3 else 0.0])
    p_d = lambda x, y, z: np.array([grad[0] * x + grad[1] * y + grad[2] * z + 0.25])
    u_sp = space.discrete_spaces["u"]
    forms = [
        ("form", ref_forms.laplace_dual, {"rhs": lambda x, y, z: np.array([0.0 * x]), "kappa": lambda x, y, z: 2.0}, len(u_sp.element_ids)),
        ("bc_form", ref_forms.laplace_dual_bc_dirichlet, {"p": p_d}, len(u_sp.bc_element_ids)),
    ]
    indptr, indices, data, rhs = ref_assembly.assemble(space, forms)
    A = sp.csr_matrix((data, indices, indptr)).tocsc()
    sol = spl.spsolve(A, -rhs)
    assert np.abs(A @ sol + rhs).max() < 1e-11
    n_u = u_sp.n_dof
    mesh = space.discrete_spaces["p"].mesh
    for i, cid in enumerate(space.discrete_spaces["p"].element_ids):
        centroid = mesh.points[mesh.cells[cid].node_tags].mean(axis=0)
        dof = space.destination_indexes(i)[-1]
        assert dof >= n_u and abs(sol[dof] - float(p_d(*centroid)[0])) < 1e-10
    # the discrete flux is the exact one: evaluate u_h at the centroid of a few cells through the oracle's basis
    from oracle.ref_forms import RefElement

    exact = -2.0 * grad[:dim]
    for i in range(0, len(u_sp.element_ids), 5):
        el = RefElement(u_sp, u_sp.element_ids[i])
        pts = np.full((1, dim), 1.0 / (dim + 1))
        x, jac, det_jac, inv_jac = el.evaluate_mapping(pts)
        tab = el.evaluate_basis(pts, jac, det_jac, inv_jac)
        dest = space.destination_indexes(i)[: tab.shape[2]]
        u_h = sol[dest] @ tab[0, 0, :, :dim]
        assert np.abs(u_h - exact).max() < 1e-10


@pytest.mark.parametrize("family,k,dim,n", [
    ("RT", 2, 2, 3), ("RT", 3, 2, 2), ("BDM", 1, 2, 3), ("BDM", 2, 2, 3), ("BDM", 3, 2, 2),
    ("RT", 2, 3, 2), ("RT", 3, 3, 1), ("BDM", 1, 3, 2), ("BDM", 2, 3, 1), ("BDM", 3, 3, 1),
])
def test_oracle_mixed_poisson_is_exact_for_polynomial_fluxes(family, k, dim, n):
    """Known answer for every H(div) element of the oracle (facet and interior moments, their scaling, the Piola map and
    the orientation transformations of all facets of a mesh with shuffled local vertices): for p = s^k / k with
    s = a . x, constant kappa and f = div u the flux u = -kappa grad p has degree k - 1, lies in RT_k and BDM_k, and
    div u lies in the pressure space -- so (u, P_h p) solves the discrete equations and the computed flux is the exact one
    at every point."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl

    from oracle import ref_assembly, ref_forms
    from oracle.ref_forms import RefElement
    from tests.parity_helpers import mixed_poisson_case

    kappa = 2.0
    a = np.array([1.0, 2.0, -1.0 if dim == 3 else 0.0])
    s = lambda x, y, z: a[0] * x + a[1] * y + a[2] * z + 0.3
    p_d = lambda x, y, z: np.array([s(x, y, z) ** k / k])
    flux = lambda x, y, z: -kappa * s(x, y, z) ** (k - 1) * a[:dim]
    rhs = lambda x, y, z: np.array([-kappa * (k - 1) * s(x, y, z) ** max(k - 2, 0) * float(a @ a) + 0.0 * x])
    case = mixed_poisson_case(dim=dim, k=k, n=n, family=family)
    space = case.ref_space()
    u_sp = space.discrete_spaces["u"]
    forms = [
        ("form", ref_forms.laplace_dual, {"rhs": rhs, "kappa": lambda x, y, z: kappa}, len(u_sp.element_ids)),
        ("bc_form", ref_forms.laplace_dual_bc_dirichlet, {"p": p_d}, len(u_sp.bc_element_ids)),
    ]
    indptr, indices, data, res = ref_assembly.assemble(space, forms)
    A = sp.csr_matrix((data, indices, indptr)).tocsc()
    sol = spl.spsolve(A, -res)
    assert np.abs(A @ sol + res).max() < 1e-9 * max(1.0, np.abs(res).max())
    pts = np.array([[0.2, 0.3, 0.1][:dim], [0.6, 0.1, 0.25][:dim]])
    scale = max(1.0, np.abs(flux(1.0, 1.0, 0.0)).max())
    for i in range(len(u_sp.element_ids)):
        el = RefElement(u_sp, u_sp.element_ids[i])
        x, jac, det_jac, inv_jac = el.evaluate_mapping(pts)
        tab = el.evaluate_basis(pts, jac, det_jac, inv_jac)
        dest = space.destination_indexes(i)[: tab.shape[2]]
        for q in range(len(pts)):
            u_h = sol[dest] @ tab[0, q, :, :dim]
            assert np.abs(u_h - flux(*x[q])).max() < 1e-8 * scale, (i, q, u_h, flux(*x[q]))


@pytest.mark.parametrize("dim,k", [(2, 1), (3, 1), (2, 2)])
def test_oracle_mixed_elasticity_reproduces_a_constant_stress(dim, k):
    """Known answer for the weakly symmetric mixed elasticity oracle (config 4's form family: row-wise BDM stresses,
    discontinuous displacement and rotation multipliers, natural displacement condition): for a linear displacement
    u = G x + b with constant Lame parameters the stress sigma = 2 mu sym(G) + lambda tr(G) I is constant, lies in the
    stress space with div sigma = 0, and the rotation is constant -- so the discrete solution carries the exact stress
    at every point, whatever the orientation of the facets."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl

    from oracle import ref_assembly
    from oracle import ref_tensor_forms as rt
    from oracle.ref_forms import RefElement
    from tests.parity_helpers import tensor_mixed_case

    lam, mu = 1.5, 0.75
    G = np.array([[0.3, -0.2, 0.5], [0.7, 0.1, -0.4], [-0.6, 0.25, 0.2]])[:dim, :dim]
    b = np.array([0.1, -0.3, 0.2])[:dim]
    sigma = mu * (G + G.T) + lam * np.trace(G) * np.eye(dim)
    tc = 3 if dim == 3 else 1
    u_d = lambda x, y, z: np.array([sum(G[r, c] * (x, y, z)[c] for c in range(dim)) + b[r] for r in range(dim)])
    functions = {"rhs": lambda x, y, z: np.zeros((dim + tc,) + np.shape(x)), "lambda": lambda x, y, z: lam, "mu": lambda x, y, z: mu}
    case = tensor_mixed_case(kind="le", dim=dim, s_family="BDM", k=k, n=2 if dim == 2 or k == 1 else 1)
    space = case.ref_space()
    s_sp = space.discrete_spaces["s"]
    forms = [("form", rt.le_dual, functions, len(s_sp.element_ids)),
             ("bc_form", rt.le_dual_bc_dirichlet, {"u": u_d}, len(s_sp.bc_element_ids))]
    indptr, indices, data, res = ref_assembly.assemble(space, forms)
    A = sp.csr_matrix((data, indices, indptr)).tocsc()
    sol = spl.spsolve(A, -res)
    assert np.abs(A @ sol + res).max() < 1e-10 * max(1.0, np.abs(res).max())
    pts = np.array([[0.2, 0.3, 0.1][:dim], [0.55, 0.15, 0.2][:dim]])
    for i in range(len(s_sp.element_ids)):
        el = RefElement(s_sp, s_sp.element_ids[i])
        x, jac, det_jac, inv_jac = el.evaluate_mapping(pts)
        tab = el.evaluate_basis(pts, jac, det_jac, inv_jac)
        n_phi = tab.shape[2]
        coef = sol[space.destination_indexes(i)[: n_phi * dim]].reshape(n_phi, dim)  # [basis, row]: component-minor dofs
        for q in range(len(pts)):
            s_h = coef.T @ tab[0, q, :, :dim]  # row r = sum_j coef[j, r] psi_j
            assert np.abs(s_h - sigma).max() < 1e-9 * np.abs(sigma).max(), (i, q, s_h, sigma)


@pytest.mark.parametrize("kind", ["lce", "lce_scaled"])
@pytest.mark.parametrize("dim", [2, 3])
def test_oracle_mixed_cosserat_reproduces_a_constant_stress(kind, dim):
    """Known answer for the mixed Cosserat oracles (BASELINE config 4's forms): a linear displacement with ANY gradient G
    and the constant micro-rotation that cancels its skew part (the convention of the primal oracle, Theta(t) = -skew G)
    give the constant symmetric stress 2 mu_s sym(G) + lambda_s tr(G) I and zero couple stress; both lie in the discrete
    spaces with zero divergence and the angular balance holds, so the discrete solution is exact -- for the plain and the
    length-scaled formulation (variable length scale l(x)).  With the opposite rotation the stress is off by O(1): the
    asymmetry coupling, the boundary term of the rotation and the primal oracle's convention are tied together."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl

    from oracle import ref_assembly
    from oracle import ref_tensor_forms as rt
    from oracle.ref_forms import RefElement
    from tests.parity_helpers import tensor_mixed_case

    lam, mu = 1.5, 0.75
    G = np.array([[0.3, -0.2, 0.5], [0.7, 0.1, -0.4], [-0.6, 0.25, 0.2]])[:dim, :dim]
    b = np.array([0.1, -0.3, 0.2])[:dim]
    sigma = mu * (G + G.T) + lam * np.trace(G) * np.eye(dim)
    t_ex = np.array([0.5 * (G[0, 1] - G[1, 0])]) if dim == 2 else \
        np.array([0.5 * (G[1, 2] - G[2, 1]), 0.5 * (G[2, 0] - G[0, 2]), 0.5 * (G[0, 1] - G[1, 0])])
    tc = len(t_ex)
    one = lambda x: np.ones_like(x)
    functions = {
        "rhs": lambda x, y, z: np.zeros((dim + tc,) + np.shape(x)),
        "lambda_s": lambda x, y, z: lam * one(x), "mu_s": lambda x, y, z: mu * one(x), "kappa_s": lambda x, y, z: 0.35 * one(x),
        "lambda_o": lambda x, y, z: 0.7 * one(x), "mu_o": lambda x, y, z: 1.3 * one(x), "kappa_o": lambda x, y, z: 0.4 * one(x),
        "l": lambda x, y, z: 0.5 + 0.25 * x + 0.1 * y * y, "grad_l": lambda x, y, z: np.array([0.25 * one(x), 0.2 * y, 0.0 * x]),
    }
    case = tensor_mixed_case(kind=kind, dim=dim, s_family="BDM", m_family="RT", k=1, n=2)
    space = case.ref_space()
    s_sp, m_sp = space.discrete_spaces["s"], space.discrete_spaces["m"]
    vol, bc = (rt.lce_dual, rt.lce_dual_bc_dirichlet) if kind == "lce" else (rt.lce_scaled_dual, rt.lce_scaled_dual_bc_dirichlet)
    pts = np.array([[0.2, 0.3, 0.1][:dim], [0.55, 0.15, 0.2][:dim]])
    mc = m_sp.n_comp

    def worst(sign):
        bc_functions = {
            "u": lambda x, y, z: np.array([sum(G[r, c] * (x, y, z)[c] for c in range(dim)) + b[r] for r in range(dim)]),
            "t": lambda x, y, z: np.array([sign * t_ex[c] * one(x) for c in range(tc)]), "l": functions["l"], "grad_l": functions["grad_l"],
        }
        forms = [("form", vol, functions, len(s_sp.element_ids)), ("bc_form", bc, bc_functions, len(s_sp.bc_element_ids))]
        indptr, indices, data, res = ref_assembly.assemble(space, forms)
        A = sp.csr_matrix((data, indices, indptr)).tocsc()
        sol = spl.spsolve(A, -res)
        assert np.abs(A @ sol + res).max() < 1e-10 * max(1.0, np.abs(res).max())
        err_s = err_m = 0.0
        for i in range(len(s_sp.element_ids)):
            el_s, el_m = RefElement(s_sp, s_sp.element_ids[i]), RefElement(m_sp, m_sp.element_ids[i])
            x, jac, det_jac, inv_jac = el_s.evaluate_mapping(pts)
            tab_s = el_s.evaluate_basis(pts, jac, det_jac, inv_jac)
            tab_m = el_m.evaluate_basis(pts, jac, det_jac, inv_jac)
            dest = space.destination_indexes(i)
            ns, nm = tab_s.shape[2], tab_m.shape[2]
            cs = sol[dest[: ns * dim]].reshape(ns, dim)
            cm = sol[dest[ns * dim : ns * dim + nm * mc]].reshape(nm, mc)
            for q in range(len(pts)):
                err_s = max(err_s, np.abs(cs.T @ tab_s[0, q, :, :dim] - sigma).max())
                err_m = max(err_m, np.abs(cm.T @ tab_m[0, q, :, :dim]).max())
        return err_s, err_m

    err_s, err_m = worst(1.0)
    assert err_s < 1e-9 * np.abs(sigma).max() and err_m < 1e-9 * np.abs(sigma).max()
    assert worst(-1.0)[0] > 1e-2 * np.abs(sigma).max()


@pytest.mark.parametrize("kind", ["degenerate", "unscaled"])
@pytest.mark.parametrize("dim,k", [(2, 1), (3, 1), (2, 2), (3, 2)])
def test_oracle_degenerate_elliptic_forms_reproduce_a_constant_flux(kind, dim, k):
    """Known answer for the Arbogast-Taicher forms of config 3 (flux mass, generalised divergence coupling AND the (q, q)
    mass block): with porosity = d_phi = 1 both forms read (v, psi) - (q, div psi) = -<q_D, psi . n>,
    (div v, w) + (q, w) = (f, w); for linear q and f = q the flux v = -grad q is constant and (v, P_h q) solves the
    discrete equations, so the computed flux is exact at every point."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl

    from oracle import ref_assembly, ref_forms
    from oracle.ref_forms import RefElement
    from tests.parity_helpers import degenerate_case

    a = np.array([1.0, 2.0, -1.0 if dim == 3 else 0.0])
    q_ex = lambda x, y, z: np.array([a[0] * x + a[1] * y + a[2] * z + 0.25])
    one = lambda x: np.ones_like(x)
    flux_name, scalar = ("u", "p") if kind == "unscaled" else ("v", "q")
    functions = {"rhs": q_ex, "porosity": lambda x, y, z: one(x), "d_phi": lambda x, y, z: one(x),
                 "grad_d_phi": lambda x, y, z: np.array([0.0 * x, 0.0 * y, 0.0 * z])}
    bc_functions = {scalar: q_ex, "porosity": functions["porosity"], "d_phi": functions["d_phi"]}
    space = degenerate_case(kind=kind, dim=dim, k=k, n=2).ref_space()
    v_sp = space.discrete_spaces[flux_name]
    vol, bc = (ref_forms.unscaled_elliptic, ref_forms.unscaled_elliptic_bc_dirichlet) if kind == "unscaled" else \
        (ref_forms.degenerate_elliptic, ref_forms.degenerate_elliptic_bc_dirichlet)
    forms = [("form", vol, functions, len(v_sp.element_ids)), ("bc_form", bc, bc_functions, len(v_sp.bc_element_ids))]
    indptr, indices, data, res = ref_assembly.assemble(space, forms)
    A = sp.csr_matrix((data, indices, indptr)).tocsc()
    sol = spl.spsolve(A, -res)
    assert np.abs(A @ sol + res).max() < 1e-10 * max(1.0, np.abs(res).max())
    pts = np.array([[0.2, 0.3, 0.1][:dim], [0.55, 0.15, 0.2][:dim]])
    for i in range(len(v_sp.element_ids)):
        el = RefElement(v_sp, v_sp.element_ids[i])
        x, jac, det_jac, inv_jac = el.evaluate_mapping(pts)
        tab = el.evaluate_basis(pts, jac, det_jac, inv_jac)
        dest = space.destination_indexes(i)[: tab.shape[2]]
        for q in range(len(pts)):
            assert np.abs(sol[dest] @ tab[0, q, :, :dim] + a[:dim]).max() < 1e-9


@pytest.mark.parametrize("dim", [2, 3])
def test_sundus_oracle_steady_linear_state_is_a_fixed_point(dim):
    """Known answer for the two-flux transport oracle (config 5) without the nonlinearity (eta = 0): linear p and c that
    do not change in time (previous level = cell averages of the same fields), constant kappa and delta, f_f = 0 and
    f_r = 1 (the reaction's constant part): one Newton step from zero lands on the exact constant fluxes
    mp = -kappa grad p, mc = -delta grad c and on the cell averages of p and c -- time-derivative terms, both
    divergence couplings, the boundary terms and their signs all enter."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl

    from oracle import ref_assembly, ref_forms
    from oracle.ref_forms import RefElement
    from tests.parity_helpers import SundusCase

    kappa, delta = 2.0, 1.5
    gp = np.array([1.0, 2.0, -1.0 if dim == 3 else 0.0])
    gc = np.array([-0.5, 0.75, 0.4 if dim == 3 else 0.0])
    p_ex = lambda x, y, z, t=0.0: np.array([gp[0] * x + gp[1] * y + gp[2] * z + 0.2])
    c_ex = lambda x, y, z, t=0.0: np.array([gc[0] * x + gc[1] * y + gc[2] * z - 0.1])
    one = lambda x: np.ones_like(x)
    functions = {"rhs_f": lambda x, y, z, t: np.array([0.0 * x]), "rhs_r": lambda x, y, z, t: np.array([one(x)]),
                 "kappa": lambda x, y, z: kappa, "delta": lambda x, y, z: delta, "eta": lambda x, y, z: 0.0, "delta_t": 0.1}
    case = SundusCase(dim=dim, k=1, n=2)
    space = case.ref_space()
    mp = space.discrete_spaces["mp"]
    n_cells = len(mp.element_ids)
    mesh = mp.mesh
    previous = np.zeros(space.n_dof)
    for i in range(n_cells):
        centroid = mesh.points[mesh.cells[mp.element_ids[i]].node_tags].mean(axis=0)
        dest = space.destination_indexes(i)
        previous[dest[-2]], previous[dest[-1]] = p_ex(*centroid)[0], c_ex(*centroid)[0]
    forms = [("time_dependent_form", ref_forms.sundus_dual, functions, n_cells),
             ("time_dependent_bc_form", ref_forms.sundus_dual_bc_dirichlet, {"p": p_ex, "c": c_ex}, len(mp.bc_element_ids))]
    indptr, indices, data, res = ref_assembly.assemble_time_dependent(space, forms, (np.zeros(space.n_dof), previous), 0.3)
    A = sp.csr_matrix((data, indices, indptr)).tocsc()
    sol = spl.spsolve(A, -res)
    pts = np.array([[0.2, 0.3, 0.1][:dim], [0.55, 0.15, 0.2][:dim]])
    for i in range(n_cells):
        el = RefElement(mp, mp.element_ids[i])
        x, jac, det_jac, inv_jac = el.evaluate_mapping(pts)
        tab = el.evaluate_basis(pts, jac, det_jac, inv_jac)
        n_u = tab.shape[2]
        dest = space.destination_indexes(i)
        for q in range(len(pts)):
            assert np.abs(sol[dest[:n_u]] @ tab[0, q, :, :dim] + kappa * gp[:dim]).max() < 1e-9
            assert np.abs(sol[dest[n_u : 2 * n_u]] @ tab[0, q, :, :dim] + delta * gc[:dim]).max() < 1e-9
        assert abs(sol[dest[-2]] - previous[dest[-2]]) < 1e-9 and abs(sol[dest[-1]] - previous[dest[-1]]) < 1e-9


def test_taylor_hood_oracle_reproduces_a_linear_displacement():
    """Known answer for the displacement-pressure (Taylor-Hood P2-P1) elasticity oracle: a linear displacement has a
    constant stress, so with f = 0 its nodal interpolant together with the constant pressure (lambda + mu) div u makes
    the residual of every row vanish -- the rows of the boundary nodes too, once the penalty term (which is zero for the
    exact boundary values) is left out.  (Solving the penalised system instead would lose six digits to the 1e12
    penalty in the factorisation.)"""
    from oracle import ref_assembly, ref_h1_multi
    from oracle.ref_basis import SUB
    from tests.parity_helpers import taylor_hood_case

    lam, mu = 3.0, 0.75
    G = np.array([[0.3, -0.2], [0.7, 0.1]])
    b = np.array([0.1, -0.3])
    u_ex = lambda x, y, z: np.array([G[0, 0] * x + G[0, 1] * y + b[0], G[1, 0] * x + G[1, 1] * y + b[1]])
    functions = {"lambda": lambda x, y, z: lam + 0.0 * x, "mu": lambda x, y, z: mu + 0.0 * x,
                 "rhs": lambda x, y, z: np.array([0.0 * x, 0.0 * x])}
    space = taylor_hood_case(n=3).ref_space()
    u_sp = space.discrete_spaces["u"]
    mesh = u_sp.mesh
    exact = np.zeros(space.n_dof)
    for i in range(len(u_sp.element_ids)):
        cell = mesh.cells[u_sp.element_ids[i]]
        X = mesh.points[cell.node_tags]
        dest = space.destination_indexes(i)
        for v in range(3):
            exact[dest[2 * v : 2 * v + 2]] = u_ex(*X[v])
        for e, (a, c) in enumerate(SUB[2][1]):
            exact[dest[6 + 2 * e : 8 + 2 * e]] = u_ex(*(0.5 * (X[a] + X[c])))
        exact[dest[12:15]] = (lam + mu) * np.trace(G)
    # interior equilibrium: volume form alone, all rows whose node is not on the boundary
    indptr, indices, data, res = ref_assembly.assemble(
        space, [("form", ref_h1_multi.le_taylor_hood, functions, len(u_sp.element_ids))], exact)
    boundary = np.zeros(space.n_dof, dtype=bool)
    for i in range(len(u_sp.bc_element_ids)):
        boundary[space.bc_destination_indexes(i)] = True
    scale = np.abs(data).max() * np.abs(exact).max()
    assert np.abs(res[~boundary]).max() < 1e-12 * scale and (~boundary).sum() > 50
    # and the penalty equations hold at the exact boundary values (the boundary form ignores the state vector: its
    # residual is -beta int u_D phi, its matrix beta int phi phi)
    ip, ix, dat, res_bc = ref_assembly.assemble(
        space, [("bc_form", ref_h1_multi.le_taylor_hood_bc_dirichlet, {"u": u_ex}, len(u_sp.bc_element_ids))])
    A_bc = sp.csr_matrix((dat, ix, ip), shape=(space.n_dof, space.n_dof))
    assert np.abs(A_bc @ exact + res_bc).max() < 1e-12 * np.abs(res_bc).max()


@pytest.mark.parametrize("dim", [2, 3])
def test_primal_cosserat_oracle_balances_a_linear_displacement_with_its_rotation(dim):
    """Known answer for the primal Cosserat oracle (fields u, t): for a linear displacement with ANY gradient G and the
    constant micro-rotation that cancels its skew part (Theta(t) = -skew G in the oracle's convention
    e = grad u + Theta(t)) the strain is symmetric, the stress 2 mu sym(G) + lambda tr(G) I is constant and symmetric,
    the couple stress vanishes -- so with zero loads every interior equation holds for the nodal interpolant.  A sign
    error between the rotation's entry into the strain and the stress' skew part entering the rotation equation would
    leave a residual proportional to kappa."""
    from oracle import ref_assembly, ref_h1_multi
    from oracle.ref_basis import SUB
    from tests.parity_helpers import lce_primal_case

    G = np.array([[0.3, -0.2, 0.5], [0.7, 0.1, -0.4], [-0.6, 0.25, 0.2]])[:dim, :dim]
    b = np.array([0.1, -0.3, 0.2])[:dim]
    if dim == 2:
        t_ex = np.array([0.5 * (G[0, 1] - G[1, 0])])
    else:
        t_ex = np.array([0.5 * (G[1, 2] - G[2, 1]), 0.5 * (G[2, 0] - G[0, 2]), 0.5 * (G[0, 1] - G[1, 0])])
    nt = len(t_ex)
    u_ex = lambda X: G @ X[:dim] + b
    functions = {"lambda": lambda x, y, z: 3.0, "mu": lambda x, y, z: 0.5, "kappa": lambda x, y, z: 0.35, "gamma": lambda x, y, z: 2.0,
                 "rhs": lambda x, y, z: np.zeros((dim + nt,) + np.shape(x))}
    case = lce_primal_case(dim=dim, k=2, kt=1, n=3 if dim == 2 else 2)
    space = case.ref_space()
    u_sp = space.discrete_spaces["u"]
    mesh = u_sp.mesh
    n_v, edges = dim + 1, SUB[dim][1]
    exact = np.zeros(space.n_dof)
    for i in range(len(u_sp.element_ids)):
        X = mesh.points[mesh.cells[u_sp.element_ids[i]].node_tags]
        dest = space.destination_indexes(i)
        nodes = [X[v] for v in range(n_v)] + [0.5 * (X[a] + X[c]) for a, c in edges]
        for j, xn in enumerate(nodes):
            exact[dest[dim * j : dim * j + dim]] = u_ex(xn)
        off = dim * len(nodes)
        for v in range(n_v):
            exact[dest[off + nt * v : off + nt * v + nt]] = t_ex
    _, _, data, res = ref_assembly.assemble(space, [("form", ref_h1_multi.lce_primal, functions, len(u_sp.element_ids))], exact)
    boundary = np.zeros(space.n_dof, dtype=bool)
    for i in range(len(u_sp.bc_element_ids)):
        boundary[space.bc_destination_indexes(i)] = True
    scale = np.abs(data).max() * max(1.0, np.abs(exact).max())
    assert (~boundary).sum() > 20 and np.abs(res[~boundary]).max() < 1e-12 * scale
    # the test has teeth: with the opposite rotation the rotation equations are violated at order kappa
    wrong = exact.copy()
    for i in range(len(u_sp.element_ids)):
        dest = space.destination_indexes(i)
        off = dim * (n_v + len(edges))
        wrong[dest[off:]] = -np.tile(t_ex, n_v)
    _, _, _, res_w = ref_assembly.assemble(space, [("form", ref_h1_multi.lce_primal, functions, len(u_sp.element_ids))], wrong)
    assert np.abs(res_w[~boundary]).max() > 1e-4 * scale
