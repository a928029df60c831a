"""ORACLE (test infrastructure only -- never imported by the product path).

Restatement of the reference's closed-form (``evaluate_form_vectorized``) mixed forms on row-wise H(div) stresses:

* ``le_dual``            -- R/src/weak_forms/le_dual_weak_form.py:354-510
* ``lce_dual``           -- R/src/weak_forms/lce_dual_weak_form.py:410-648 (``scaled=True``:
                            R/src/weak_forms/lce_scaled_dual_weak_form.py:445-699)
* ``le_dual_bc_dirichlet`` / ``lce_dual_bc_dirichlet`` -- le_dual_weak_form.py:513-578, lce_dual_weak_form.py:651-756,
                            lce_scaled_dual_weak_form.py:702-811
* ``le_riesz_map`` / ``lce_riesz_map`` -- le_riesz_map_weak_form.py:258-360, lce_riesz_map_weak_form.py:372-546,
                            lce_scaled_riesz_map_weak_form.py:415-598

The block operators are built the way the reference builds them (``outer_fun`` / ``inner_fun`` with ``np.tensordot`` and
``np.block`` per quadrature point, then a weighted sum over the points); the Jacobian is the closed form and the
residual is ``J alpha + load`` (what the reference's AD variant ``evaluate_form`` returns for these linear forms; the
vectorised variant returns the load only -- ``alpha=None`` reproduces that).  PARITY: pinned against the reference's own source above the basix boundary, through BOTH of its
entry points (``evaluate_form_vectorized`` at alpha = 0, the AD quadrature loop ``evaluate_form`` at a non-zero state, config 4's order
included: tests/golden/refcode_l*.npz, tools/make_reference_pins.py); basix's element conventions stay unpinned (ref_basis.py).
"""
import numpy as np

from . import ref_hdiv
from .ref_forms import RefElement, _bc_setup, normal, transform_lower_to_higher


def outer_fun(phi):  # lce_dual_weak_form.py:503-506
    n = phi.shape[0]
    t = np.tensordot(phi, phi.T, axes=0)
    return 0.5 * np.block([[t[j, :, :, i] for j in range(n)] for i in range(n)])


def inner_fun(phi):  # lce_dual_weak_form.py:508-512
    n, nc = phi.shape
    t = np.tensordot(phi, phi.T, axes=1)
    return 0.5 * np.block([[t[j, i] * np.identity(nc) for j in range(n)] for i in range(n)])


def _tables(space, names, iel):
    sp = [space.discrete_spaces[n] for n in names]
    els = [RefElement(s, s.element_ids[iel]) for s in sp]
    dim = els[0].cell.dimension
    points, weights = space.quadrature[dim]
    x, jac, det_jac, inv_jac = els[0].evaluate_mapping(points)
    tabs = [e.evaluate_basis(points, jac, det_jac, inv_jac) for e in els]
    return sp, dim, np.asarray(weights), x, det_jac, tabs


def _compliance(phi_star, wd, lam, mu, kappa, dim, scale=1.0):
    """(s,s) / (m,m) tensor block: symm / (2 mu) + skew / (2 kappa) - vol (lce_dual_weak_form.py:514-529)."""
    gen_outer = np.array([outer_fun(p) for p in phi_star]).T
    gen_inner = np.array([inner_fun(p) for p in phi_star]).T
    outer = np.array([np.outer(p, p) for p in phi_star]).T
    vol_factor = (1.0 / (2.0 * mu)) * (lam / (2.0 * mu + dim * lam))
    return (
        (gen_inner + gen_outer) @ (wd * scale / (2.0 * mu))
        + (gen_inner - gen_outer) @ (wd * scale / (2.0 * kappa))
        - outer @ (wd * scale * vol_factor)
    )


def _asym_blocks(j_el, t_tab, s_star, wd, dim, t_off, n_t, tc, s_off, n_s, sc):
    """(t,s) and (s,t) blocks (lce_dual_weak_form.py:531-581)."""
    A = [np.einsum("q,qk,qa->ka", wd, t_tab[0, :, :, 0], s_star[:, :, d]) for d in range(dim)]
    if dim == 3:
        comp_map = {0: [2, 1], 1: [0, 2], 2: [1, 0]}
        op_map = {0: (A[1], -A[2]), 1: (A[2], -A[0]), 2: (A[0], -A[1])}
    else:
        comp_map = {0: [1, 0]}
        op_map = {0: (A[0], -A[1])}
    for c in range(tc):
        b, e = c + t_off, c + t_off + n_t
        for cs, op in zip(comp_map[c], op_map[c]):
            bs, es = s_off + cs, s_off + cs + n_s
            j_el[b:e:tc, bs:es:sc] += -1.0 * op
            j_el[bs:es:sc, b:e:tc] += -1.0 * op.T


def _div_blocks(j_el, l_tab, weights, div_star, l_off, n_l, lc_, h_off, n_h, hc):
    """(u,s)/(s,u) or (t,m)/(m,t): same component only; ``div_star[q, a]`` already carries every factor but w phi."""
    block = (weights * l_tab[0, :, :, 0].T) @ div_star
    for c in range(lc_):
        if c >= hc:
            continue
        lb, le = l_off + c, l_off + c + n_l
        hb, he = h_off + c, h_off + c + n_h
        j_el[lb:le:lc_, hb:he:hc] += block
        j_el[hb:he:hc, lb:le:lc_] += block.T


def _finish(j_el, r_load, alpha):
    if alpha is None:
        return r_load, j_el
    return j_el @ np.asarray(alpha, dtype=float) + r_load, j_el


def le_dual(space, functions, iel, alpha):
    sp, dim, weights, x, det_jac, (s_tab, u_tab, t_tab) = _tables(space, ("s", "u", "t"), iel)
    sc, uc, tc = (s.n_comp for s in sp)
    n_s, n_u, n_t = s_tab.shape[2] * sc, u_tab.shape[2] * uc, t_tab.shape[2] * tc
    n_dof = n_s + n_u + n_t
    j_el, r_el = np.zeros((n_dof, n_dof)), np.zeros(n_dof)
    f = np.asarray(functions["rhs"](x[:, 0], x[:, 1], x[:, 2]))
    lam = functions["lambda"](x[:, 0], x[:, 1], x[:, 2]) * np.ones(len(weights))
    mu = functions["mu"](x[:, 0], x[:, 1], x[:, 2]) * np.ones(len(weights))
    wd = det_jac * weights
    for c in range(uc):
        r_el[n_s + c : n_s + c + n_u : uc] += -1.0 * (wd * u_tab[0, :, :, 0].T) @ f[c]
    for c in range(tc):
        r_el[n_s + n_u + c : n_s + n_u + c + n_t : tc] += -1.0 * (wd * t_tab[0, :, :, 0].T) @ f[c + uc]
    s_star = s_tab[0, :, :, 0:dim]
    mass = np.array([p @ p.T for p in s_star]).T @ (wd / (2.0 * mu))
    for c in range(sc):
        j_el[c:n_s:sc, c:n_s:sc] += mass
    vol_factor = (1.0 / (2.0 * mu)) * (lam / (2.0 * mu + dim * lam))
    j_el[0:n_s, 0:n_s] += -1.0 * np.array([np.outer(p, p) for p in s_star]).T @ (vol_factor * wd)
    _asym_blocks(j_el, t_tab, s_star, wd, dim, n_s + n_u, n_t, tc, 0, n_s, sc)
    div_s = np.trace(s_tab[1 : 1 + dim, :, :, 0:dim], axis1=0, axis2=3)
    _div_blocks(j_el, u_tab, weights, div_s, n_s, n_u, uc, 0, n_s, sc)
    return _finish(j_el, r_el, alpha)


def lce_dual(space, functions, iel, alpha, scaled=False):
    sp, dim, weights, x, det_jac, (s_tab, m_tab, u_tab, t_tab) = _tables(space, ("s", "m", "u", "t"), iel)
    sc, mc, uc, tc = (s.n_comp for s in sp)
    n_s, n_m = s_tab.shape[2] * sc, m_tab.shape[2] * mc
    n_u, n_t = u_tab.shape[2] * uc, t_tab.shape[2] * tc
    n_dof = n_s + n_m + n_u + n_t
    j_el, r_el = np.zeros((n_dof, n_dof)), np.zeros(n_dof)
    ones = np.ones(len(weights))
    ev = lambda k: functions[k](x[:, 0], x[:, 1], x[:, 2]) * ones
    f = np.asarray(functions["rhs"](x[:, 0], x[:, 1], x[:, 2]))
    wd = det_jac * weights
    o_u, o_t = n_s + n_m, n_s + n_m + n_u
    for c in range(uc):
        r_el[o_u + c : o_u + c + n_u : uc] += -1.0 * (wd * u_tab[0, :, :, 0].T) @ f[c]
    for c in range(tc):
        r_el[o_t + c : o_t + c + n_t : tc] += -1.0 * (wd * t_tab[0, :, :, 0].T) @ f[c + uc]
    s_star = s_tab[0, :, :, 0:dim]
    j_el[0:n_s, 0:n_s] += _compliance(s_star, wd, ev("lambda_s"), ev("mu_s"), ev("kappa_s"), dim)
    _asym_blocks(j_el, t_tab, s_star, wd, dim, o_t, n_t, tc, 0, n_s, sc)
    lc_v = ev("l")
    m_scale = ones if scaled else 1.0 / lc_v
    m_star = m_tab[0, :, :, 0:dim]
    if dim == 3:
        j_el[n_s:o_u, n_s:o_u] += _compliance(m_star, wd, ev("lambda_o"), ev("mu_o"), ev("kappa_o"), dim, m_scale)
    else:
        m_inner = np.array([p @ p.T for p in m_star]).T
        block = m_inner @ (wd * m_scale * (1.0 / (ev("mu_o") + ev("kappa_o"))))
        for c in range(mc):
            j_el[n_s + c : o_u : mc, n_s + c : o_u : mc] += block
    div_s = np.trace(s_tab[1 : 1 + dim, :, :, 0:dim], axis1=0, axis2=3)
    _div_blocks(j_el, u_tab, weights, div_s, o_u, n_u, uc, 0, n_s, sc)
    div_m = np.trace(m_tab[1 : 1 + dim, :, :, 0:dim], axis1=0, axis2=3)
    if scaled:  # grad l . m det + l div^ m  (lce_scaled_dual_weak_form.py:669-687)
        grad_l = np.asarray(functions["grad_l"](x[:, 0], x[:, 1], x[:, 2]))
        gl_m = np.array([[grad_l[0:dim, i] @ m_tab[0, i, j, 0:dim] for j in range(m_tab.shape[2])] for i in range(len(weights))])
        div_m = gl_m * det_jac[:, None] + div_m * lc_v[:, None]
    _div_blocks(j_el, t_tab, weights, div_m, o_t, n_t, tc, n_s, n_m, mc)
    return _finish(j_el, r_el, alpha)


def lce_scaled_dual(space, functions, iel, alpha):
    return lce_dual(space, functions, iel, alpha, scaled=True)


def _trace_normal_load(space, field, iel, value_at, n_comp):
    """``-int_F g_c (psi . n)`` on the facet dofs of the first neighbour, component-minor (lce_dual_weak_form.py:735-754)."""
    f_space, el, weights, x, det_jac, _ = _bc_setup(space, field, iel)
    cell = el.cell
    points, _ = space.bc_quadrature[cell.dimension]
    neigh = RefElement(f_space, f_space.first_neighbour(cell.id))
    mapped = transform_lower_to_higher(np.asarray(points).reshape(len(weights), -1), cell, neigh.cell)
    _, jac0, det0, inv0 = neigh.evaluate_mapping(mapped)
    tr_tab = neigh.evaluate_basis(mapped, jac0, det0, inv0)
    facet_index = neigh.cell.sub_cells_ids[cell.dimension].tolist().index(cell.id)
    dofs = ref_hdiv.entity_dofs(neigh.family, neigh.cell.dimension, neigh.k)[cell.dimension][facet_index]
    n = normal(f_space.mesh, neigh.cell, cell)
    dim = neigh.cell.dimension
    n_phi = len(dofs)
    r = np.zeros(n_phi * n_comp)
    for c in range(n_comp):
        res = np.zeros(n_phi)
        for i, omega in enumerate(weights):
            res -= det_jac[i] * omega * value_at(x[i], c) * (tr_tab[0, i, dofs, 0:dim] @ n[0:dim])
        r[c : c + n_phi * n_comp : n_comp] += res
    return r


def _trace_normal_penalty(space, field, iel, value_at, n_comp, beta=1.0e12):
    """Penalty blocks of one field: ``J_cc = beta int (psi.n)(psi.n)^T``, ``r_c = -beta int g_c (psi.n)`` on the facet dofs
    of the first neighbour (lce_dual_weak_form.py:835-869)."""
    f_space, el, weights, x, det_jac, _ = _bc_setup(space, field, iel)
    cell = el.cell
    points, _ = space.bc_quadrature[cell.dimension]
    neigh = RefElement(f_space, f_space.first_neighbour(cell.id))
    mapped = transform_lower_to_higher(np.asarray(points).reshape(len(weights), -1), cell, neigh.cell)
    _, jac0, det0, inv0 = neigh.evaluate_mapping(mapped)
    tr_tab = neigh.evaluate_basis(mapped, jac0, det0, inv0)
    facet_index = neigh.cell.sub_cells_ids[cell.dimension].tolist().index(cell.id)
    dofs = ref_hdiv.entity_dofs(neigh.family, neigh.cell.dimension, neigh.k)[cell.dimension][facet_index]
    n = normal(f_space.mesh, neigh.cell, cell)
    dim = neigh.cell.dimension
    n_phi = len(dofs)
    n_dof = n_phi * n_comp
    r, j = np.zeros(n_dof), np.zeros((n_dof, n_dof))
    for c in range(n_comp):
        res, block = np.zeros(n_phi), np.zeros((n_phi, n_phi))
        for i, omega in enumerate(weights):
            phi = tr_tab[0, i, dofs, 0:dim] @ n[0:dim]
            res -= beta * det_jac[i] * omega * value_at(x[i], c) * phi
            block += beta * det_jac[i] * omega * np.outer(phi, phi)
        r[c : c + n_dof : n_comp] += res
        j[c : c + n_dof : n_comp, c : c + n_dof : n_comp] += block
    return r, j


def le_dual_bc_neumann(space, functions, iel, alpha):
    """le_dual_weak_form.py:581-644 (there named LCEDualWeakFormBCNeumann): stress field only."""
    sc = space.discrete_spaces["s"].n_comp
    return _trace_normal_penalty(space, "s", iel, lambda p, c: functions["s"](p[0], p[1], p[2])[c], sc)


def lce_dual_bc_neumann(space, functions, iel, alpha):
    """lce_dual_weak_form.py:759-871 (= lce_scaled_dual_weak_form.py:814-917)."""
    sc, mc = space.discrete_spaces["s"].n_comp, space.discrete_spaces["m"].n_comp
    r_s, j_s = _trace_normal_penalty(space, "s", iel, lambda p, c: functions["s"](p[0], p[1], p[2])[c], sc)
    r_m, j_m = _trace_normal_penalty(space, "m", iel, lambda p, c: functions["m"](p[0], p[1], p[2])[c], mc)
    n_s, n_m = len(r_s), len(r_m)
    j = np.zeros((n_s + n_m, n_s + n_m))
    j[:n_s, :n_s] = j_s
    j[n_s:, n_s:] = j_m
    return np.concatenate([r_s, r_m]), j


def le_dual_bc_dirichlet(space, functions, iel, alpha):
    sc = space.discrete_spaces["s"].n_comp
    r = _trace_normal_load(space, "s", iel, lambda p, c: functions["u"](p[0], p[1], p[2])[c], sc)
    return r, np.zeros((len(r), len(r)))


def lce_dual_bc_dirichlet(space, functions, iel, alpha, scaled=False):
    sc, mc = space.discrete_spaces["s"].n_comp, space.discrete_spaces["m"].n_comp
    r_s = _trace_normal_load(space, "s", iel, lambda p, c: functions["u"](p[0], p[1], p[2])[c], sc)
    if scaled:
        t_val = lambda p, c: functions["t"](p[0], p[1], p[2])[c] * (functions["l"](p[0], p[1], p[2]) + 1.0e-16)
    else:
        t_val = lambda p, c: functions["t"](p[0], p[1], p[2])[c]
    r_m = _trace_normal_load(space, "m", iel, t_val, mc)
    r = np.concatenate([r_s, r_m])
    return r, np.zeros((len(r), len(r)))


def lce_scaled_dual_bc_dirichlet(space, functions, iel, alpha):
    return lce_dual_bc_dirichlet(space, functions, iel, alpha, scaled=True)


def _mass_blocks(j_el, tab, wd, off, n, nc):
    block = np.array([np.outer(p, p) for p in tab[0, :, :, 0]]).T @ wd
    for c in range(nc):
        j_el[off + c : off + n : nc, off + c : off + n : nc] += block


def _divdiv_blocks(j_el, div_phys, wd, off, n, nc):
    """``div_phys[q, a]``: the (generalised) physical divergence of every vector function."""
    block = np.array([np.outer(d, d) for d in div_phys]).T @ wd
    for c in range(nc):
        j_el[off + c : off + n : nc, off + c : off + n : nc] += block


def le_riesz_map(space, functions, iel, alpha):
    """LERieszMapWeakForm.evaluate_form_vectorized, R/src/weak_forms/le_riesz_map_weak_form.py:258-360."""
    sp, dim, weights, x, det_jac, (s_tab, u_tab, t_tab) = _tables(space, ("s", "u", "t"), iel)
    sc, uc, tc = (s.n_comp for s in sp)
    n_s, n_u, n_t = s_tab.shape[2] * sc, u_tab.shape[2] * uc, t_tab.shape[2] * tc
    n_dof = n_s + n_u + n_t
    j_el = np.zeros((n_dof, n_dof))
    lam = functions["lambda"](x[:, 0], x[:, 1], x[:, 2]) * np.ones(len(weights))
    mu = functions["mu"](x[:, 0], x[:, 1], x[:, 2]) * np.ones(len(weights))
    wd = det_jac * weights
    s_star = s_tab[0, :, :, 0:dim]
    mass = np.array([p @ p.T for p in s_star]).T @ (wd / (2.0 * mu))
    for c in range(sc):
        j_el[c:n_s:sc, c:n_s:sc] += mass
    vol_factor = (1.0 / (2.0 * mu)) * (lam / (2.0 * mu + dim * lam))
    j_el[0:n_s, 0:n_s] += -1.0 * np.array([np.outer(p, p) for p in s_star]).T @ (vol_factor * wd)
    div_s = np.trace(s_tab[1 : 1 + dim, :, :, 0:dim], axis1=0, axis2=3) / det_jac[:, None]
    _divdiv_blocks(j_el, div_s, wd, 0, n_s, sc)
    _mass_blocks(j_el, u_tab, wd, n_s, n_u, uc)
    _mass_blocks(j_el, t_tab, wd, n_s + n_u, n_t, tc)
    return _finish(j_el, np.zeros(n_dof), alpha)


def lce_riesz_map(space, functions, iel, alpha, scaled=False):
    """LCERieszMapWeakForm / LCEScaledRieszMapWeakForm .evaluate_form_vectorized,
    R/src/weak_forms/lce_riesz_map_weak_form.py:372-546, lce_scaled_riesz_map_weak_form.py:415-598."""
    sp, dim, weights, x, det_jac, (s_tab, m_tab, u_tab, t_tab) = _tables(space, ("s", "m", "u", "t"), iel)
    sc, mc, uc, tc = (s.n_comp for s in sp)
    n_s, n_m = s_tab.shape[2] * sc, m_tab.shape[2] * mc
    n_u, n_t = u_tab.shape[2] * uc, t_tab.shape[2] * tc
    n_dof = n_s + n_m + n_u + n_t
    j_el = np.zeros((n_dof, n_dof))
    ones = np.ones(len(weights))
    ev = lambda k: functions[k](x[:, 0], x[:, 1], x[:, 2]) * ones
    wd = det_jac * weights
    o_u, o_t = n_s + n_m, n_s + n_m + n_u
    s_star, m_star = s_tab[0, :, :, 0:dim], m_tab[0, :, :, 0:dim]
    j_el[0:n_s, 0:n_s] += _compliance(s_star, wd, ev("lambda_s"), ev("mu_s"), ev("kappa_s"), dim)
    lc_v = ev("l")
    m_scale = ones if scaled else 1.0 / lc_v
    if dim == 3:
        j_el[n_s:o_u, n_s:o_u] += _compliance(m_star, wd, ev("lambda_o"), ev("mu_o"), ev("kappa_o"), dim, m_scale)
    else:
        block = np.array([p @ p.T for p in m_star]).T @ (wd * m_scale * (1.0 / (ev("mu_o") + ev("kappa_o"))))
        for c in range(mc):
            j_el[n_s + c : o_u : mc, n_s + c : o_u : mc] += block
    div_s = np.trace(s_tab[1 : 1 + dim, :, :, 0:dim], axis1=0, axis2=3) / det_jac[:, None]
    _divdiv_blocks(j_el, div_s, wd, 0, n_s, sc)
    div_m = np.trace(m_tab[1 : 1 + dim, :, :, 0:dim], axis1=0, axis2=3) / det_jac[:, None]
    if scaled:
        grad_l = np.asarray(functions["grad_l"](x[:, 0], x[:, 1], x[:, 2]))
        gl_m = np.array([[grad_l[0:dim, i] @ m_tab[0, i, j, 0:dim] for j in range(m_tab.shape[2])] for i in range(len(weights))])
        div_m = gl_m + div_m * lc_v[:, None]
    _divdiv_blocks(j_el, div_m, wd, n_s, n_m, mc)
    _mass_blocks(j_el, u_tab, wd, o_u, n_u, uc)
    _mass_blocks(j_el, t_tab, wd, o_t, n_t, tc)
    return _finish(j_el, np.zeros(n_dof), alpha)


def lce_scaled_riesz_map(space, functions, iel, alpha):
    return lce_riesz_map(space, functions, iel, alpha, scaled=True)
