"""Host logic of the multi-field Lagrange forms on the CPU: the term lists read off a form's constitutive map
(``H1MultiFieldForm.term_lists``), contracted with the mapped basis exactly the way ``nfa_local_h1_multifield`` does,
reproduce the oracle's local matrices (oracle/ref_h1_multi.py restates the reference cell by cell)."""
import numpy as np
import pytest

from nfempy_b200.weak_forms.h1_common import scalar_table
from oracle import ref_h1_multi
from tests.parity_helpers import lce_primal_case, stress_constraint_case, taylor_hood_case


def _walk_terms(form, space, cell, coef_fns):
    """numpy version of the kernel's entry loop for one cell (coefficients evaluated at the quadrature points)."""
    fields = [space.discrete_spaces[n] for n in form.field_names]
    dim = fields[0].dimension
    points, weights = space.quadrature[dim]
    tabs = [scalar_table(s.ref_element, points) for s in fields]
    tab = np.concatenate(tabs, axis=2)
    dof_a, dof_cr, a0, cr0 = [], [], 0, 0
    for s, t in zip(fields, tabs):
        for a in range(t.shape[2]):
            for c in range(s.n_comp):
                dof_a.append(a0 + a)
                dof_cr.append(cr0 + c)
        a0 += t.shape[2]
        cr0 += s.n_comp
    ptr, code, val = form.term_lists(dim, cr0)
    X = fields[0].mesh.points[fields[0].cell_nodes()[cell]]
    J = (X[1:] - X[0]).T  # [3, dim]
    Jinv = np.linalg.pinv(J)
    det = np.sqrt(np.linalg.det(J.T @ J))
    lam = np.array([1.0 - points[:, :dim].sum(axis=1)] + [points[:, d] for d in range(dim)]).T
    xq = lam @ X
    D = np.empty((len(weights), tab.shape[2], 1 + dim))
    D[:, :, 0] = tab[0]
    D[:, :, 1:] = np.einsum("kqa,kx->qax", tab[1:], Jinv)[:, :, :dim]
    coefs = np.array([f(xq) for f in coef_fns])  # [M, n_q]
    n_loc = len(dof_a)
    K = np.zeros((n_loc, n_loc))
    for i in range(n_loc):
        for j in range(n_loc):
            pair = dof_cr[i] * cr0 + dof_cr[j]
            for t in range(ptr[pair], ptr[pair + 1]):
                s_, t_, m = code[t] & 15, (code[t] >> 4) & 15, code[t] >> 8
                K[i, j] += val[t] * np.sum(weights * det * coefs[m] * D[:, dof_a[i], s_] * D[:, dof_a[j], t_])
    return K


@pytest.mark.parametrize("which", ["lce3", "lce2", "taylor_hood", "stress3"])
def test_term_lists_reproduce_the_oracle_matrix(which):
    if which == "lce3":
        case, fn = lce_primal_case(dim=3, k=2, kt=1, n=1, varying=True), ref_h1_multi.lce_primal
    elif which == "lce2":
        case, fn = lce_primal_case(dim=2, k=2, kt=2, n=2, varying=True), ref_h1_multi.lce_primal
    elif which == "taylor_hood":
        case, fn = taylor_hood_case(n=2, varying=True), ref_h1_multi.le_taylor_hood
    else:
        case, fn = stress_constraint_case(dim=3, n=1), ref_h1_multi.le_stress_constraint
    space, rspace = case.space(), case.ref_space()
    form = case.gpu_forms(space, None)["volume"][2]
    F = form.functions
    ev = lambda f: (lambda x: f(x[:, 0], x[:, 1], x[:, 2]) * np.ones(len(x)))
    if which.startswith("lce"):
        coef_fns = [ev(F[k]) for k in ("lambda", "mu", "kappa", "gamma")]
    elif which == "taylor_hood":
        coef_fns = [ev(F["lambda"]), ev(F["mu"]), lambda x: np.ones(len(x)),
                    lambda x: 1.0 / (ev(F["lambda"])(x) + ev(F["mu"])(x))]
    else:
        coef_fns = [lambda x: 1.0e6 * ev(F["lambda"])(x), lambda x: 1.0e6 * ev(F["mu"])(x)]
    functions = case.ref_forms_fn(rspace)[0][2]
    for cell in (0, len(space.discrete_spaces["u"].elements) - 1):
        K = _walk_terms(form, space, cell, coef_fns)
        _, j_ref = fn(rspace, functions, cell, np.zeros(K.shape[0]))
        assert np.abs(K - j_ref).max() <= 1e-12 * np.abs(j_ref).max()


def test_field_order_is_checked():
    from nfempy_b200 import _lib
    from nfempy_b200.mesh import square_tri_mesh
    from nfempy_b200.spaces import ProductSpace
    from nfempy_b200.weak_forms import LCEPrimalWeakForm

    mesh = square_tri_mesh(2)
    space = ProductSpace({"t": (2, 1, "Lagrange", 1, mesh), "u": (2, 2, "Lagrange", 1, mesh)})
    space.build_structures({"u": [1], "t": [1]}, {})
    with pytest.raises(_lib.NfaError):
        LCEPrimalWeakForm(space)._fields()


@pytest.mark.parametrize("dim", [2, 3])
def test_primal_cosserat_rigid_motions_are_in_the_null_space(dim):
    """Known answer for the oracle's LCE primal restatement (sign of Theta(t), skew part, couple stress): a translation
    and an infinitesimal rigid rotation u = W x with the matching micro-rotation t (Theta(t) = -W, grad t = 0) produce no
    strain, no curvature and therefore a vanishing homogeneous residual J alpha on every cell."""
    from oracle import ref_assembly

    case = lce_primal_case(dim=dim, k=2, kt=1, n=1 if dim == 3 else 2, varying=True, with_bc=False)
    rspace = case.ref_space()
    functions = dict(case.ref_forms_fn(rspace)[0][2])
    functions["rhs"] = lambda x, y, z: np.zeros((6 if dim == 3 else 3,) + np.shape(x))
    u, t = rspace.discrete_spaces["u"], rspace.discrete_spaces["t"]
    nt = 3 if dim == 3 else 1
    # nodal coordinates of the Lagrange dofs: interpolate by evaluating at the cell's nodes through the oracle's own basis
    omega = np.array([0.3, -0.2, 0.5])
    W = np.array([[0.0, -omega[2], omega[1]], [omega[2], 0.0, -omega[0]], [-omega[1], omega[0], 0.0]])
    n_cells = len(u.element_ids)
    for motion in ("translation", "rotation"):
        for iel in range(n_cells):
            cell = u.mesh.cells[u.element_ids[iel]]
            X = u.mesh.points[cell.node_tags]  # vertices
            edges = [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)] if dim == 3 else [(1, 2), (0, 2), (0, 1)]
            nodes_u = np.vstack([X] + [0.5 * (X[a] + X[b])[None, :] for a, b in edges])  # P2 nodes: vertices, edge midpoints
            nodes_t = X  # P1 nodes
            if motion == "translation":
                a_u = np.tile(np.array([0.7, -0.4, 0.2])[:dim], len(nodes_u))
                a_t = np.zeros(len(nodes_t) * nt)
            else:
                Wd = W if dim == 3 else W[:2, :2]
                a_u = (nodes_u[:, :dim] @ Wd.T).reshape(-1)
                # Theta(t) = [[0, -t2, t1], [t2, 0, -t0], [-t1, t0, 0]] must equal -W: t = -omega (2-D: t = -omega_z)
                tv = -omega if dim == 3 else np.array([-omega[2]])
                a_t = np.tile(tv, len(nodes_t))
            alpha_l = np.concatenate([a_u, a_t])
            r_el, j_el = ref_h1_multi.lce_primal(rspace, functions, iel, alpha_l)
            assert np.abs(r_el).max() <= 1e-12 * np.abs(j_el).max(), (motion, iel)
