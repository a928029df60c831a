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
