"""ORACLE (test infrastructure only -- never imported by the product path).

Per-cell restatement, one quadrature point at a time and in the reference's own order of operations, of the primal
forms on several Lagrange fields:

* ``lce_primal``               -- R/src/weak_forms/lce_primal_weak_form.py:9-220 (2-D branch :73-118, 3-D branch :120-206)
* ``lce_primal_bc_dirichlet``  -- R/src/weak_forms/lce_primal_weak_form.py:223-297
* ``le_taylor_hood``           -- R/src/weak_forms/le_taylor_hood_weak_form.py:9-186 (2-D branch :82-129; 3-D asserts False)
* ``le_taylor_hood_bc_dirichlet`` -- R/src/weak_forms/le_taylor_hood_weak_form.py:189-237
* ``le_stress_constraint``     -- R/src/weak_forms/le_primal_stress_constraint_weak_form.py:10-126

The reference differentiates its residual with ``auto_diff``; the forms are linear in ``alpha``, so the Jacobian is the
homogeneous residual of the unit vectors (``ref_forms._linearise``, see that module's header).  PARITY: the reference's own classes, run over
stand-ins for basix / auto_diff (tools/make_reference_pins.py), give these matrices bit for bit
(tests/golden/refcode_lce_primal_*, refcode_taylor_hood_*, refcode_stress_constraint_*); P1 / P2 Lagrange bases are unique,
so nothing of basix's conventions enters.
"""
import numpy as np

from .ref_forms import RefElement, _grad_phi, _linearise


def _volume_setup(space, names, iel):
    spaces = [space.discrete_spaces[n] for n in names]
    els = [RefElement(s, s.element_ids[iel]) for s in spaces]
    dim = els[0].cell.dimension
    points, weights = space.quadrature[dim]
    x, jac, det_jac, inv_jac = els[0].evaluate_mapping(points)
    tabs = [e.evaluate_basis(points, jac, det_jac, inv_jac) for e in els]
    return spaces, dim, weights, x, det_jac, inv_jac, tabs


def lce_primal(space, functions, iel, alpha):
    (u_space, t_space), dim, weights, x, det_jac, inv_jac, (u_tab, t_tab) = _volume_setup(space, ("u", "t"), iel)
    uc, tc = u_space.n_comp, t_space.n_comp
    n_u_phi, n_t_phi = u_tab.shape[2], t_tab.shape[2]
    n_u_dof, n_t_dof = n_u_phi * uc, n_t_phi * tc
    n_dof = n_u_dof + n_t_dof
    f_val = np.asarray(functions["rhs"](x[:, 0], x[:, 1], x[:, 2]))
    f_lambda, f_mu, f_kappa, f_gamma = (functions[k] for k in ("lambda", "mu", "kappa", "gamma"))
    eye = np.identity(dim)

    def residual(a, load):
        el_form = np.zeros(n_dof)
        if load:
            for c in range(uc):
                el_form[c : c + n_u_dof : uc] += det_jac * weights * u_tab[0, :, :, 0].T @ f_val[c]
            for c in range(tc):
                b = c + n_u_dof
                el_form[b : b + n_t_dof : tc] += det_jac * weights * t_tab[0, :, :, 0].T @ f_val[c + uc]
        for i, omega in enumerate(weights):
            xv = x[i]
            gu = _grad_phi(dim, inv_jac[i], u_tab, i)
            gt = _grad_phi(dim, inv_jac[i], t_tab, i)
            grad_uh = np.vstack([a[c : n_u_dof + c : uc] @ gu for c in range(dim)])
            if dim == 2:
                a_t = a[n_u_dof : n_t_dof + n_u_dof : tc]
                grad_th = (a_t @ gt).reshape(1, dim)
                th = a_t @ t_tab[0, i, :, 0:tc]  # [1]
                theta_outer = th[0] * np.array([[0.0, -1.0], [1.0, 0.0]])
            else:
                a_tc = [a[c + n_u_dof : n_u_dof + n_t_dof + c : tc] for c in range(3)]
                grad_th = np.vstack([a_tc[c] @ gt for c in range(3)])
                th = np.array([a_tc[c] @ t_tab[0, i, :, 0] for c in range(3)])
                theta_outer = np.array([[0.0, -th[2], +th[1]], [+th[2], 0.0, -th[0]], [-th[1], +th[0], 0.0]])
            eh = grad_uh + theta_outer
            symm_eh, skew_eh = 0.5 * (eh + eh.T), 0.5 * (eh - eh.T)
            sh = (2.0 * f_mu(xv[0], xv[1], xv[2]) * symm_eh + 2.0 * f_kappa(xv[0], xv[1], xv[2]) * skew_eh
                  + f_lambda(xv[0], xv[1], xv[2]) * np.trace(eh) * eye)
            skew_sh = 0.5 * (sh - sh.T)
            if dim == 2:
                s_cross = np.array([[skew_sh[1, 0] - skew_sh[0, 1]]])
            else:
                s_cross = np.array([[skew_sh[2, 1] - skew_sh[1, 2], skew_sh[0, 2] - skew_sh[2, 0],
                                     skew_sh[1, 0] - skew_sh[0, 1]]])
            k = f_gamma(xv[0], xv[1], xv[2]) * grad_th
            strain_energy_h = (gu @ sh.T).reshape((n_u_dof,))
            curvature_energy_h = (gt @ k.T + t_tab[0, i, :, 0:1] @ s_cross).reshape((n_t_dof,))
            integrand = np.concatenate([strain_energy_h, curvature_energy_h])
            el_form = el_form + det_jac[i] * omega * integrand
        return el_form

    return _linearise(residual, alpha, n_dof)


def _penalty_blocks(space, field, data_fn, iel, beta=1.0e12):
    f_space = space.discrete_spaces[field]
    el = RefElement(f_space, f_space.bc_element_ids[iel], bc=True)
    dim = el.cell.dimension
    points, weights = space.bc_quadrature[dim]
    x, jac, det_jac, inv_jac = el.evaluate_mapping(points)
    tab = el.evaluate_basis(points, jac, det_jac, inv_jac)
    nc, n_phi = f_space.n_comp, tab.shape[2]
    n_dof = n_phi * nc
    j_el, r_el = np.zeros((n_dof, n_dof)), np.zeros(n_dof)
    block = np.zeros((n_phi, n_phi))
    for i, omega in enumerate(weights):
        phi = tab[0, i, :, 0]
        block += beta * det_jac[i] * omega * np.outer(phi, phi)
    for c in range(nc):
        res = np.zeros(n_phi)
        for i, omega in enumerate(weights):
            phi = tab[0, i, :, 0]
            d_v = np.asarray(data_fn(x[i, 0], x[i, 1], x[i, 2])).reshape(-1)
            res -= beta * det_jac[i] * omega * d_v[c] * phi
        r_el[c : c + n_dof : nc] += res
        j_el[c : c + n_dof : nc, c : c + n_dof : nc] += block
    return r_el, j_el


def lce_primal_bc_dirichlet(space, functions, iel, alpha):
    r_u, j_u = _penalty_blocks(space, "u", functions["u"], iel)
    r_t, j_t = _penalty_blocks(space, "t", functions["t"], iel)
    n_u, n_t = len(r_u), len(r_t)
    j_el = np.zeros((n_u + n_t, n_u + n_t))
    j_el[:n_u, :n_u] = j_u
    j_el[n_u:, n_u:] = j_t
    return np.concatenate([r_u, r_t]), j_el


def le_taylor_hood(space, functions, iel, alpha):
    (u_space, p_space), dim, weights, x, det_jac, inv_jac, (u_tab, p_tab) = _volume_setup(space, ("u", "p"), iel)
    assert dim == 2, "the reference's 3-D branch starts with `assert False`"
    uc = u_space.n_comp
    n_u_phi, n_p_phi = u_tab.shape[2], p_tab.shape[2]
    n_u_dof, n_p_dof = n_u_phi * uc, n_p_phi
    n_dof = n_u_dof + n_p_dof
    f_val = np.asarray(functions["rhs"](x[:, 0], x[:, 1], x[:, 2]))
    lam = functions["lambda"](x[:, 0], x[:, 1], x[:, 2]) * np.ones(len(weights))
    mu = functions["mu"](x[:, 0], x[:, 1], x[:, 2]) * np.ones(len(weights))
    eye = np.identity(dim)

    def residual(a, load):
        el_form = np.zeros(n_dof)
        if load:
            for c in range(uc):
                el_form[c : c + n_u_dof : uc] += det_jac * weights * u_tab[0, :, :, 0].T @ f_val[c]
        for i, omega in enumerate(weights):
            gu = _grad_phi(dim, inv_jac[i], u_tab, i)
            phi_p = p_tab[0, i, :, 0:1]  # [n_p_phi, 1]
            grad_uh = np.vstack([a[c : n_u_dof + c : uc] @ gu for c in range(dim)])
            eh = 0.5 * (grad_uh + grad_uh.T)
            div_uh = np.trace(grad_uh)
            ph = (a[n_u_dof : n_u_dof + n_p_dof] @ phi_p).reshape(1, 1)
            sh = 2.0 * mu[i] * eh + lam[i] * np.trace(eh) * eye
            sh_dev = sh - (1.0 / dim) * np.trace(sh) * eye
            equ_1 = gu @ (sh_dev + ph * eye).T
            equ_2 = phi_p * div_uh - (1.0 / (lam[i] + mu[i])) * phi_p @ ph.T
            integrand = np.concatenate([equ_1.reshape(n_u_dof), equ_2.reshape(n_p_dof)])
            el_form = el_form + det_jac[i] * omega * integrand
        return el_form

    return _linearise(residual, alpha, n_dof)


def le_taylor_hood_bc_dirichlet(space, functions, iel, alpha):
    return _penalty_blocks(space, "u", functions["u"], iel)


def le_stress_constraint(space, functions, iel, alpha):
    (u_space,), dim, weights, x, det_jac, inv_jac, (u_tab,) = _volume_setup(space, ("u",), iel)
    uc = u_space.n_comp
    n_dof = u_tab.shape[2] * uc
    beta = 1.0e6
    st_all = np.asarray(functions["s_target"](x[:, 0], x[:, 1], x[:, 2]), dtype=float) * np.ones((dim, dim, len(weights)))
    lam = functions["lambda"](x[:, 0], x[:, 1], x[:, 2]) * np.ones(len(weights))
    mu = functions["mu"](x[:, 0], x[:, 1], x[:, 2]) * np.ones(len(weights))
    eye = np.identity(dim)

    def residual(a, load):
        el_form = np.zeros(n_dof)
        for i, omega in enumerate(weights):
            gu = _grad_phi(dim, inv_jac[i], u_tab, i)
            grad_uh = np.vstack([a[c : n_dof + c : uc] @ gu for c in range(dim)])
            eh = 0.5 * (grad_uh + grad_uh.T)
            sh = 2.0 * mu[i] * eh + lam[i] * np.trace(eh) * eye
            st = st_all[:, :, i] if load else np.zeros((dim, dim))
            strain_energy_h = (gu @ (beta * (sh.T - st.T))).reshape((n_dof,))
            el_form = el_form + det_jac[i] * omega * strain_energy_h
        return el_form

    return _linearise(residual, alpha, n_dof)
