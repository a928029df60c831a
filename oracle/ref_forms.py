"""ORACLE (test infrastructure only -- never imported by the product path).

Per-cell restatement of the reference weak forms, one quadrature point at a time, in the reference's own order of
operations.  The reference differentiates its residual with ``auto_diff`` (third-party, absent); every library form
is linear in ``alpha`` (SURVEY.md section 8a row AD), so the Jacobian is obtained here the way forward-mode AD
would: the homogeneous part of the residual is pushed through the same code for the ``n_loc`` unit vectors
(``_linearise``).  That is exact for linear forms and keeps the arithmetic of each column identical to the residual's.

Forms restated:
* ``le_primal``               -- R/src/weak_forms/le_primal_weak_form.py:13-129
* ``le_primal_bc_dirichlet``  -- R/src/weak_forms/le_primal_weak_form.py:131-176
* ``le_primal_bc_neumann``    -- R/src/weak_forms/le_primal_weak_form.py:179-221
* ``laplace_primal``          -- R/src/weak_forms/laplace_primal_weak_form.py:10-74 (with ``quadrature[dim]``)
* ``laplace_primal_bc_dirichlet`` -- R/src/weak_forms/laplace_primal_weak_form.py:77-157

PARITY: pinned against the reference's own source above the basix boundary.  tools/make_reference_pins.py imports
/root/reference/src unmodified over stand-ins for meshio / basix / auto_diff (tools/ref_shims) and runs the classes restated
here through ``SequentialAssembler.scatter_forms``; tests/test_cpu_reference_pins.py compares this module with the result
(tests/golden/refcode_*.npz): bit-identical for every form of this file except the Sundus tangent (complex step here, AD
there: 5e-16).  ``laplace_primal*`` alone has no counterpart run: the reference's class is stale (it unpacks
``space.quadrature`` as a pair, R/src/weak_forms/laplace_primal_weak_form.py:26, where ProductSpace now keeps a list per
dimension) -- it is checked by known answers only.  basix's own element conventions stay unpinned (oracle/ref_basis.py).
"""
import numpy as np

from . import ref_basis


class RefElement:
    """What ``space.elements[i]`` offers the forms: mapping + mapped basis of one cell."""

    def __init__(self, space, cell_id, bc=False):
        self.space = space
        self.mesh = space.mesh
        self.cell = space.mesh.cells[cell_id]
        if bc:
            self.family, self.k, self.disc = space.bc_family, space.bc_k_order, space.bc_discontinuous
        else:
            self.family, self.k, self.disc = space.family, space.k_order, space.discontinuous
        if self.cell.dimension == 1 and space.generator_family in ("RT", "BDM", "N1E", "N2E"):
            self.family = "Lagrange"

    def evaluate_mapping(self, points):  # R/src/basis/finite_element.py:128-154
        dim = self.cell.dimension
        cell_points = self.mesh.points[self.cell.node_tags]
        if dim == 0:
            return ref_basis.evaluate_mapping(0, np.ones(1), cell_points)
        phi = ref_basis.evaluate_linear_shapes(dim, points)
        return ref_basis.evaluate_mapping(dim, phi, cell_points)

    def evaluate_basis(self, points, jac, det_jac, inv_jac):  # R/src/basis/finite_element.py:156-186
        dim = self.cell.dimension
        if dim == 0:
            return np.ones((1, 1, 1, 1))
        tab = ref_basis.tabulate(self.family, dim, self.k, points, self.disc)
        mapped = ref_basis.push_forward(self.family, tab[0], jac, det_jac, inv_jac)
        if tab[0].shape == mapped.shape:
            tab[0] = mapped
        else:
            tab = np.insert(tab, 2, 0, axis=3)
            tab[0] = mapped
        if self.family != "Lagrange" or self.k > 2:
            from . import ref_hdiv

            tab = ref_hdiv.permute_and_transform(tab, self)
        return tab


def _linearise(residual, alpha, n_dof):
    """``r = residual(alpha, load=True)``; column ``j`` of the Jacobian = homogeneous residual of ``e_j``."""
    r_el = residual(np.asarray(alpha, dtype=float), True)
    j_el = np.zeros((n_dof, n_dof))
    for j in range(n_dof):
        e = np.zeros(n_dof)
        e[j] = 1.0
        j_el[:, j] = residual(e, False)
    return r_el, j_el


def _grad_phi(dim, inv_jac_i, tab, i):
    rows = [inv_jac_i @ np.eye(3)[d] for d in range(dim)]
    return (np.vstack(rows) @ tab[1 : tab.shape[0] + 1, i, :, 0]).T  # [nphi, dim]


def le_primal(space, functions, iel, alpha):
    u_space = space.discrete_spaces["u"]
    nc = u_space.n_comp
    el = RefElement(u_space, u_space.element_ids[iel])
    dim = el.cell.dimension
    points, weights = space.quadrature[dim]
    x, jac, det_jac, inv_jac = el.evaluate_mapping(points)
    tab = el.evaluate_basis(points, jac, det_jac, inv_jac)
    n_phi = tab.shape[2]
    n_dof = n_phi * nc
    f_val = np.asarray(functions["rhs"](x[:, 0], x[:, 1], x[:, 2]))
    lam = functions["lambda"](x[:, 0], x[:, 1], x[:, 2]) * np.ones(len(weights))
    mu = functions["mu"](x[:, 0], x[:, 1], x[:, 2]) * np.ones(len(weights))
    eye = np.identity(dim)

    def residual(a, load):
        el_form = np.zeros(n_dof)
        if load:
            for c in range(nc):
                el_form[c : c + n_dof : nc] += det_jac * weights * tab[0, :, :, 0].T @ f_val[c]
        for i, omega in enumerate(weights):
            g = _grad_phi(dim, inv_jac[i], tab, i)
            grad_uh = np.vstack([a[c : n_dof + c : nc] @ g for c in range(dim)])
            eh = 0.5 * (grad_uh + grad_uh.T)
            sh = 2.0 * mu[i] * eh + lam[i] * np.trace(eh) * eye
            el_form = el_form + det_jac[i] * omega * (g @ sh.T).reshape((n_dof,))
        return el_form

    return _linearise(residual, alpha, n_dof)


def _bc_setup(space, field, iel):
    f_space = space.discrete_spaces[field]
    el = RefElement(f_space, f_space.bc_element_ids[iel], bc=True)
    dim = el.cell.dimension
    points, weights = space.bc_quadrature[dim]
    x, jac, det_jac, inv_jac = el.evaluate_mapping(points)
    tab = el.evaluate_basis(points, jac, det_jac, inv_jac)
    return f_space, el, weights, x, det_jac, tab


def le_primal_bc_dirichlet(space, functions, iel, alpha):
    u_space, el, weights, x, det_jac, tab = _bc_setup(space, "u", iel)
    nc = u_space.n_comp
    n_phi = tab.shape[2]
    n_dof = n_phi * nc
    j_el, r_el = np.zeros((n_dof, n_dof)), np.zeros(n_dof)
    beta = 1.0e12
    block = np.zeros((n_phi, n_phi))
    for i, omega in enumerate(weights):
        phi = tab[0, i, :, 0]
        block += beta * det_jac[i] * omega * np.outer(phi, phi)
    for c in range(nc):
        res = np.zeros(n_phi)
        for i, omega in enumerate(weights):
            phi = tab[0, i, :, 0]
            u_d = functions["u"](x[i, 0], x[i, 1], x[i, 2])
            res -= beta * det_jac[i] * omega * u_d[c] * phi
        r_el[c : c + n_dof : nc] += res
        j_el[c : c + n_dof : nc, c : c + n_dof : nc] += block
    return r_el, j_el


def le_primal_bc_neumann(space, functions, iel, alpha):
    u_space, el, weights, x, det_jac, tab = _bc_setup(space, "u", iel)
    nc = u_space.n_comp
    n_phi = tab.shape[2]
    n_dof = n_phi * nc
    j_el, r_el = np.zeros((n_dof, n_dof)), np.zeros(n_dof)
    for c in range(nc):
        res = np.zeros(n_phi)
        for i, omega in enumerate(weights):
            phi = tab[0, i, :, 0]
            t_n = functions["t"](x[i, 0], x[i, 1], x[i, 2])
            res += det_jac[i] * omega * t_n[c] * phi
        r_el[c : c + n_dof : nc] += res
    return r_el, j_el


def laplace_primal(space, functions, iel, alpha):
    p_space = space.discrete_spaces["p"]
    nc = p_space.n_comp
    el = RefElement(p_space, p_space.element_ids[iel])
    dim = el.cell.dimension
    points, weights = space.quadrature[dim]
    x, jac, det_jac, inv_jac = el.evaluate_mapping(points)
    tab = el.evaluate_basis(points, jac, det_jac, inv_jac)
    n_dof = tab.shape[2] * nc
    f_val = np.asarray(functions["rhs"](x[:, 0], x[:, 1], x[:, 2]))
    phi_s_star = det_jac * weights * tab[0, :, :, 0].T

    def residual(a, load):
        el_form = np.zeros(n_dof)
        if load:
            for c in range(nc):
                el_form[c : c + n_dof : nc] -= (phi_s_star @ f_val[c].T).ravel()
        for i, omega in enumerate(weights):
            xv = x[i]
            g = _grad_phi(dim, inv_jac[i], tab, i)
            grad_uh = (a @ g) * functions["kappa"](xv[0], xv[1], xv[2])
            el_form = el_form + det_jac[i] * omega * (g @ grad_uh.T).reshape((n_dof,))
        return el_form

    return _linearise(residual, alpha, n_dof)


def transform_lower_to_higher(points, cell, neigh_cell):  # R/src/basis/parametric_transformation.py:8-98
    dim = cell.dimension
    facet_index = neigh_cell.sub_cells_ids[dim].tolist().index(cell.id)
    ref_vertices = np.concatenate([np.zeros((1, dim + 1)), np.eye(dim + 1)], axis=0)
    facet = list(ref_basis.SUB[dim + 1][dim][facet_index])
    sub = neigh_cell.node_tags[facet]
    own = cell.node_tags
    if tuple(own) != tuple(sub):
        index_map = dict(zip(sub.tolist(), facet))
        facet = [index_map[k] for k in own.tolist()]
    if dim == 0:
        return np.array([ref_vertices[facet[0]] * xi for xi in points])
    if dim == 1:
        return np.array([ref_vertices[facet[0]] * (1 - xi) + ref_vertices[facet[1]] * xi for xi in points[:, 0]])
    return np.array(
        [
            ref_vertices[facet[0]] * (1 - xi - eta) + ref_vertices[facet[1]] * xi + ref_vertices[facet[2]] * eta
            for xi, eta in points
        ]
    )


def laplace_primal_bc_dirichlet(space, functions, iel, alpha):
    p_space, el, weights, x, det_jac, _ = _bc_setup(space, "p", iel)
    nc = p_space.n_comp
    cell = el.cell
    points, _ = space.bc_quadrature[cell.dimension]
    neigh_id = p_space.first_neighbour(cell.id)
    neigh = RefElement(p_space, neigh_id)
    mapped = transform_lower_to_higher(points, cell, neigh.cell)
    _, jac0, det0, inv0 = neigh.evaluate_mapping(mapped)
    tr_tab = neigh.evaluate_basis(mapped, jac0, det0, inv0)
    ned = ref_basis.num_entity_dofs(neigh.family, neigh.cell.dimension, neigh.k, neigh.disc)
    # dof index of the first dof of each sub-entity of the neighbour (R/...laplace_primal_weak_form.py:119-131)
    first_dof, run = [], 0
    for d, per in enumerate(ned):
        first_dof.append([])
        for n in per:
            first_dof[d].append(run if n > 0 else None)
            run += n
    dof_idx = []
    for d in range(cell.dimension + 1):
        for sid in cell.sub_cells_ids[d]:
            idx = neigh.cell.sub_cells_ids[d].tolist().index(sid)
            if first_dof[d][idx] is not None:
                dof_idx.append(first_dof[d][idx])
    dof_idx = np.array(dof_idx)
    n_phi = len(dof_idx)
    n_dof = n_phi * nc
    j_el, r_el = np.zeros((n_dof, n_dof)), np.zeros(n_dof)
    beta = 1.0e12
    block = np.zeros((n_phi, n_phi))
    for i, omega in enumerate(weights):
        phi = tr_tab[0, i, dof_idx, 0]
        block += beta * det_jac[i] * omega * np.outer(phi, phi)
    for c in range(nc):
        res = np.zeros(n_phi)
        for i, omega in enumerate(weights):
            phi = tr_tab[0, i, dof_idx, 0]
            p_d = functions["p"](x[i, 0], x[i, 1], x[i, 2])
            res -= beta * det_jac[i] * omega * p_d[c] * phi
        r_el[c : c + n_dof : nc] += res
        j_el[c : c + n_dof : nc, c : c + n_dof : nc] += block
    return r_el, j_el


# ---- mixed Poisson (H(div) x L2) ----------------------------------------------------------------------
def normal(mesh, cell_c0, cell_c1):  # R/src/geometry/compute_normal.py:7-24
    xc_c1 = np.mean(mesh.points[cell_c1.node_tags], axis=0)
    xc_c0 = np.mean(mesh.points[cell_c0.node_tags], axis=0)
    outward = (xc_c1 - xc_c0) / np.linalg.norm(xc_c1 - xc_c0)
    if cell_c1.dimension == 0:
        n = outward / np.linalg.norm(outward)
    elif cell_c1.dimension == 1:
        v = mesh.points[cell_c1.node_tags[1]] - mesh.points[cell_c1.node_tags[0]]
        v = v / np.linalg.norm(v)
        w = np.cross(v, outward) / np.linalg.norm(np.cross(v, outward))
        n = np.cross(v, w)
    else:
        v = mesh.points[cell_c1.node_tags[0]] - xc_c1
        w = mesh.points[cell_c1.node_tags[1]] - xc_c1
        n = np.cross(v, w) / np.linalg.norm(np.cross(v, w))
    if n @ outward < 0.0:
        n = n * -1.0
    return n


def laplace_dual(space, functions, iel, alpha):
    """LaplaceDualWeakForm.evaluate_form, R/src/weak_forms/laplace_dual_weak_form.py:10-87.  The reference
    overwrites ``alpha`` with zeros (:50), so the residual is the load only."""
    u_space, p_space = space.discrete_spaces["u"], space.discrete_spaces["p"]
    nu_c, np_c = u_space.n_comp, p_space.n_comp
    el_u = RefElement(u_space, u_space.element_ids[iel])
    el_p = RefElement(p_space, p_space.element_ids[iel])
    dim = el_u.cell.dimension
    points, weights = space.quadrature[dim]
    x, jac, det_jac, inv_jac = el_u.evaluate_mapping(points)
    u_tab = el_u.evaluate_basis(points, jac, det_jac, inv_jac)
    p_tab = el_p.evaluate_basis(points, jac, det_jac, inv_jac)
    n_q_dof, n_p_dof = u_tab.shape[2] * nu_c, p_tab.shape[2] * np_c
    n_dof = n_q_dof + n_p_dof
    f_val = np.asarray(functions["rhs"](x[:, 0], x[:, 1], x[:, 2]))
    phi_s_star = det_jac * weights * p_tab[0, :, :, 0].T

    def residual(a, load):
        a = a.reshape(1, -1)
        el_form = np.zeros(n_dof)
        if load:
            for c in range(np_c):
                el_form[c + n_q_dof : c + n_q_dof + n_p_dof : np_c] -= phi_s_star @ f_val[c].T
        for i, omega in enumerate(weights):
            xv = x[i]
            psi_h = u_tab[0, i, :, 0:dim]
            w_h = p_tab[0, i, :, 0:dim]
            grad_psi_h = u_tab[1 : u_tab.shape[0] + 1, i, :, 0:dim]
            div_psi_h = np.array([np.trace(grad_psi_h, axis1=0, axis2=2) / det_jac[i]])
            u_h = a[:, 0:n_q_dof] @ psi_h
            u_h = u_h * (1.0 / functions["kappa"](xv[0], xv[1], xv[2]))
            p_h = a[:, n_q_dof:n_dof] @ w_h
            div_uh = a[:, 0:n_q_dof] @ div_psi_h.T
            integrand = np.zeros((1, n_dof))
            integrand[:, 0:n_q_dof] = (u_h @ psi_h.T) - (p_h @ div_psi_h)
            integrand[:, n_q_dof:n_dof] = div_uh @ w_h.T
            el_form = el_form + det_jac[i] * omega * integrand.reshape((n_dof,))
        return el_form

    return _linearise(residual, np.zeros(n_dof), n_dof)


def laplace_dual_bc_dirichlet(space, functions, iel, alpha):
    """LaplaceDualWeakFormBCDirichlet.evaluate_form, R/src/weak_forms/laplace_dual_weak_form.py:90-154."""
    u_space, el, weights, x, det_jac, q_tab = _bc_setup(space, "u", iel)
    nc = u_space.n_comp
    cell = el.cell
    points, _ = space.bc_quadrature[cell.dimension]
    n_q_phi = q_tab.shape[2]
    n_dof = n_q_phi * nc
    j_el, r_el = np.zeros((n_dof, n_dof)), np.zeros(n_dof)
    neigh = RefElement(u_space, u_space.first_neighbour(cell.id))
    mapped = transform_lower_to_higher(np.asarray(points).reshape(len(weights), -1), cell, neigh.cell)
    _, jac0, det0, inv0 = neigh.evaluate_mapping(mapped)
    tr_tab = neigh.evaluate_basis(mapped, jac0, det0, inv0)
    facet_index = neigh.cell.sub_cells_ids[cell.dimension].tolist().index(cell.id)
    from . import ref_hdiv

    dof_n_index = ref_hdiv.entity_dofs(neigh.family, neigh.cell.dimension, neigh.k)[cell.dimension][facet_index]
    n = normal(u_space.mesh, neigh.cell, cell)
    dim = neigh.cell.dimension
    for c in range(nc):
        res = np.zeros(n_q_phi)
        for i, omega in enumerate(weights):
            p_d = functions["p"](x[i, 0], x[i, 1], x[i, 2])
            phi = tr_tab[0, i, dof_n_index, 0:dim] @ n[0:dim]
            res += det_jac[i] * omega * p_d[c] * phi
        r_el[c : c + n_dof : nc] += res
    return r_el, j_el


def le_primal_bc_normal_dirichlet(space, functions, iel, alpha):
    """LEPrimalWeakFormBCNormalDirichlet.evaluate_form, R/src/weak_forms/le_primal_weak_form.py:224-286."""
    u_space, el, weights, x, det_jac, tab = _bc_setup(space, "u", iel)
    nc = u_space.n_comp
    cell = el.cell
    dim = cell.dimension
    n_phi = tab.shape[2]
    n_dof = n_phi * nc
    j_el, r_el = np.zeros((n_dof, n_dof)), np.zeros(n_dof)
    neigh_cell = u_space.mesh.cells[u_space.first_neighbour(cell.id)]
    n = normal(u_space.mesh, neigh_cell, cell)
    beta = 1.0e12
    for c in range(nc):
        block = np.zeros((n_phi, n_phi))
        for i, omega in enumerate(weights):
            phi = tab[0, i, :, 0] * n[c]
            block += beta * det_jac[i] * omega * np.outer(phi, phi)
        res = np.zeros(n_phi)
        for i, omega in enumerate(weights):
            phi = tab[0, i, :, 0] * n[c]
            u_d = functions["u"](x[i, 0], x[i, 1], x[i, 2])
            res -= beta * det_jac[i] * omega * np.dot(u_d, n[0 : dim + 1]) * phi
        r_el[c : c + n_dof : nc] += res
        j_el[c : c + n_dof : nc, c : c + n_dof : nc] += block
    return r_el, j_el


# ---- two-flux mixed transport with reaction (config 5) --------------------------------------------------------
def sundus_dual(space, functions, iel, alpha_n_p_1, alpha_n, t):
    """SundusDualWeakForm.evaluate_form, R/examples/parabolic_diffusive_problem/SundusWeakForm.py:11-177.
    The form is nonlinear in c (R = 1 + eta c^2); the reference differentiates it with forward AD.  Here the same
    residual code runs in complex arithmetic and the Jacobian is the complex-step derivative
    ``Im r(alpha + i h e_j) / h`` (h = 1e-30), which is exact to round-off for this polynomial residual."""
    sp = [space.discrete_spaces[n] for n in ("mp", "mc", "p", "c")]
    els = [RefElement(s, s.element_ids[iel]) for s in sp]
    dim = els[0].cell.dimension
    points, weights = space.quadrature[dim]
    x, jac, det_jac, inv_jac = els[0].evaluate_mapping(points)
    vp, vc, wp, wc = (e.evaluate_basis(points, jac, det_jac, inv_jac) for e in els)
    n_mp, n_mc, n_p, n_c = vp.shape[2], vc.shape[2], wp.shape[2], wc.shape[2]
    o_mc, o_p, o_c = n_mp, n_mp + n_mc, n_mp + n_mc + n_p
    n_dof = o_c + n_c
    f_f = np.asarray(functions["rhs_f"](x[:, 0], x[:, 1], x[:, 2], t))
    f_r = np.asarray(functions["rhs_r"](x[:, 0], x[:, 1], x[:, 2], t))
    wp_star = det_jac * weights * wp[0, :, :, 0].T
    wc_star = det_jac * weights * wc[0, :, :, 0].T
    delta_t = functions["delta_t"]
    alpha_n = np.asarray(alpha_n, dtype=float)

    def residual(a):
        a = a.reshape(1, -1)
        el_form = np.zeros(n_dof, dtype=a.dtype)
        el_form[o_p:o_c] -= wp_star @ f_f[0].T
        el_form[o_c:n_dof] -= wc_star @ f_r[0].T
        for i, omega in enumerate(weights):
            xv = x[i]
            vp_h, vc_h = vp[0, i, :, 0:dim], vc[0, i, :, 0:dim]
            wp_h, wc_h = wp[0, i, :, 0:dim], wc[0, i, :, 0:dim]
            div_vp = np.array([[np.trace(vp[1 : 1 + dim, i, j, 0:dim]) / det_jac[i] for j in range(n_mp)]])
            div_vc = np.array([[np.trace(vc[1 : 1 + dim, i, j, 0:dim]) / det_jac[i] for j in range(n_mc)]])
            mp_h = (a[:, 0:o_mc] @ vp_h) * (1.0 / functions["kappa"](xv[0], xv[1], xv[2]))
            mc_h = (a[:, o_mc:o_p] @ vc_h) * (1.0 / functions["delta"](xv[0], xv[1], xv[2]))
            p_h, c_h = a[:, o_p:o_c] @ wp_h, a[:, o_c:n_dof] @ wc_h
            p_n, c_n = alpha_n[o_p:o_c] @ wp_h, alpha_n[o_c:n_dof] @ wc_h
            div_mp, div_mc = a[:, 0:o_mc] @ div_vp.T, a[:, o_mc:o_p] @ div_vc.T
            r_h = 1.0 + functions["eta"](xv[0], xv[1], xv[2]) * c_h * c_h
            dph_dt, dch_dt = (p_h - p_n) / delta_t, (c_h - c_n) / delta_t
            integrand = np.zeros((1, n_dof), dtype=a.dtype)
            integrand[:, 0:o_mc] = (mp_h @ vp_h.T) - (p_h @ div_vp)
            integrand[:, o_mc:o_p] = (mc_h @ vc_h.T) - (c_h @ div_vc)
            integrand[:, o_p:o_c] = dph_dt @ wp_h.T + div_mp @ wp_h.T
            integrand[:, o_c:n_dof] = dch_dt @ wc_h.T + div_mc @ wc_h.T + r_h @ wc_h.T
            el_form = el_form + det_jac[i] * omega * integrand.reshape((n_dof,))
        return el_form

    a0 = np.asarray(alpha_n_p_1, dtype=float)
    r_el = residual(a0)
    j_el = np.zeros((n_dof, n_dof))
    h = 1.0e-30
    for j in range(n_dof):
        ac = a0.astype(complex)
        ac[j] += 1j * h
        j_el[:, j] = residual(ac).imag / h
    return r_el, j_el


def sundus_dual_bc_dirichlet(space, functions, iel, alpha, t):
    """SundusDualWeakFormBCDirichlet.evaluate_form, R/examples/parabolic_diffusive_problem/SundusWeakForm.py:180-269."""
    from . import ref_hdiv

    out = []
    for field, key in (("mp", "p"), ("mc", "c")):
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
        res = np.zeros(len(dofs))
        for i, omega in enumerate(weights):
            val = functions[key](x[i, 0], x[i, 1], x[i, 2], t)
            res += det_jac[i] * omega * val[0] * (tr_tab[0, i, dofs, 0:dim] @ n[0:dim])
        out.append(res)
    r_el = np.concatenate(out)
    return r_el, np.zeros((len(r_el), len(r_el)))


def l2_projector(space, functions, iel, alpha):
    """L2ProjectorWeakForm.evaluate_form, R/src/weak_forms/l2_projector_weak_form.py:8-93."""
    r_els, j_els = [], []
    for name, sp in space.discrete_spaces.items():
        el = RefElement(sp, sp.element_ids[iel])
        dim = el.cell.dimension
        points, weights = space.quadrature[dim]
        x, jac, det_jac, inv_jac = el.evaluate_mapping(points)
        tab = el.evaluate_basis(points, jac, det_jac, inv_jac)
        nc = sp.n_comp
        n_phi = tab.shape[2]
        n_dof = n_phi * nc
        j_el, r_el = np.zeros((n_dof, n_dof)), np.zeros(n_dof)
        f_val = np.asarray(functions[name](x[:, 0], x[:, 1], x[:, 2]))
        lagrange = sp.family == "Lagrange"
        if lagrange:
            phi_s_star = det_jac * weights * tab[0, :, :, 0].T
            for c in range(nc):
                r_el[c : c + n_dof : nc] -= (phi_s_star @ f_val[c].T).ravel()
        else:
            for c in range(nc):
                for i, omega in enumerate(weights):
                    phi_s = det_jac[i] * weights[i] * tab[0, i, :, 0:dim].T
                    r_el[c : c + n_dof : nc] -= (phi_s.T @ f_val[c, 0:dim, i]).ravel()
        for c in range(nc):
            for i, omega in enumerate(weights):
                if lagrange:
                    phi = tab[0, i, :, 0]
                    j_el[c : c + n_dof : nc, c : c + n_dof : nc] += det_jac[i] * omega * np.outer(phi, phi)
                else:
                    for d in range(dim):
                        j_el[c : c + n_dof : nc, c : c + n_dof : nc] += (
                            det_jac[i] * omega * np.outer(tab[0, i, :, d], tab[0, i, :, d]))
        r_els.append(r_el)
        j_els.append(j_el)
    sizes = [len(r) for r in r_els]
    off = np.add.accumulate([0] + sizes)
    n = off[-1]
    j_all, r_all = np.zeros((n, n)), np.zeros(n)
    for k in range(len(sizes)):
        r_all[off[k] : off[k + 1]] += r_els[k]
        j_all[off[k] : off[k + 1], off[k] : off[k + 1]] += j_els[k]
    return r_all, j_all


# ---- Arbogast-Taicher degenerate elliptic family ---------------------------------------------------------------
def _mixed_setup(space, flux, scalar, iel):
    v_space, q_space = space.discrete_spaces[flux], space.discrete_spaces[scalar]
    el_v = RefElement(v_space, v_space.element_ids[iel])
    el_q = RefElement(q_space, q_space.element_ids[iel])
    dim = el_v.cell.dimension
    points, weights = space.quadrature[dim]
    x, jac, det_jac, inv_jac = el_v.evaluate_mapping(points)
    v_tab = el_v.evaluate_basis(points, jac, det_jac, inv_jac)
    q_tab = el_q.evaluate_basis(points, jac, det_jac, inv_jac)
    return dim, weights, x, det_jac, v_tab, q_tab, v_tab.shape[2] * v_space.n_comp, q_tab.shape[2] * q_space.n_comp


def degenerate_elliptic(space, functions, iel, alpha):
    """DegenerateEllipticWeakForm.evaluate_form, R/src/weak_forms/degenerate_elliptic_weak_form.py:41-144."""
    dim, weights, x, det_jac, v_tab, q_tab, n_v, n_q = _mixed_setup(space, "v", "q", iel)
    n_dof = n_v + n_q
    n_v_phi = v_tab.shape[2]
    phi_star = functions["porosity"](x[:, 0], x[:, 1], x[:, 2]) * np.ones(len(weights))
    delta_star = functions["d_phi"](x[:, 0], x[:, 1], x[:, 2]) * np.ones(len(weights))
    grad_delta_star = np.asarray(functions["grad_d_phi"](x[:, 0], x[:, 1], x[:, 2]))
    f_val = np.asarray(functions["rhs"](x[:, 0], x[:, 1], x[:, 2]))
    phi_q_star = det_jac * weights * q_tab[0, :, :, 0].T

    def residual(a, load):
        a = a.reshape(1, -1)
        el_form = np.zeros(n_dof)
        if load:
            el_form[n_v:n_dof] -= phi_q_star @ f_val[0].T
        for i, omega in enumerate(weights):
            phi, delta, grad_delta = phi_star[i], delta_star[i], grad_delta_star[:, i]
            psi_h = v_tab[0, i, :, 0:dim]
            w_h = q_tab[0, i, :, 0:dim]
            v_h = a[:, 0:n_v] @ psi_h
            q_h = a[:, n_v:n_dof] @ w_h
            grad_psi_h = v_tab[1 : v_tab.shape[0] + 1, i, :, 0:dim]
            div_psi_h = np.array([np.trace(grad_psi_h, axis1=0, axis2=2) / det_jac[i]])
            gdp = np.array([[np.dot(grad_delta[0:dim], psi_h[j, 0:dim]) for j in range(n_v_phi)]])
            if not np.isclose(phi, 0.0) and phi > 0.0:
                div_phi_h_s = (delta * div_psi_h + gdp) / np.sqrt(phi)
            else:
                div_phi_h_s = 0.0 * (delta * div_psi_h + gdp)
            div_delta_v_h = a[:, 0:n_v] @ div_phi_h_s.T
            integrand = np.zeros((1, n_dof))
            integrand[:, 0:n_v] = (v_h @ psi_h.T) - (q_h @ div_phi_h_s)
            integrand[:, n_v:n_dof] = div_delta_v_h @ w_h.T + q_h @ w_h.T
            el_form = el_form + det_jac[i] * omega * integrand.reshape((n_dof,))
        return el_form

    return _linearise(residual, alpha, n_dof)


def unscaled_elliptic(space, functions, iel, alpha, phi_threshold=1.0e-16):
    """UnscaledEllipticWeakForm.evaluate_form, R/src/weak_forms/unscaled_elliptic_weak_form.py:59-153."""
    dim, weights, x, det_jac, u_tab, p_tab, n_u, n_p = _mixed_setup(space, "u", "p", iel)
    n_dof = n_u + n_p
    phi_star = functions["porosity"](x[:, 0], x[:, 1], x[:, 2]) * np.ones(len(weights))
    with np.errstate(invalid="ignore"):
        sqrt_phi_star = np.sqrt(phi_star)
    delta_star = functions["d_phi"](x[:, 0], x[:, 1], x[:, 2]) * np.ones(len(weights))
    f_val = np.asarray(functions["rhs"](x[:, 0], x[:, 1], x[:, 2]))
    phi_q_star = det_jac * weights * p_tab[0, :, :, 0].T

    def residual(a, load):
        a = a.reshape(1, -1)
        el_form = np.zeros(n_dof)
        if load:
            el_form[n_u:n_dof] -= (sqrt_phi_star * phi_q_star) @ f_val[0].T
        for i, omega in enumerate(weights):
            phi, delta = phi_star[i], delta_star[i]
            psi_h = u_tab[0, i, :, 0:dim]
            w_h = p_tab[0, i, :, 0:dim]
            u_h = a[:, 0:n_u] @ psi_h
            p_h = a[:, n_u:n_dof] @ w_h
            grad_psi_h = u_tab[1 : u_tab.shape[0] + 1, i, :, 0:dim]
            div_psi_h = np.array([np.trace(grad_psi_h, axis1=0, axis2=2) / det_jac[i]])
            if np.isclose(phi, 0.0) or phi < 0.0:
                phi = phi_threshold
            if np.isclose(delta, 0.0) or delta < 0.0:
                delta = phi_threshold
            div_u_h = a[:, 0:n_u] @ div_psi_h.T
            integrand = np.zeros((1, n_dof))
            integrand[:, 0:n_u] = ((1 / (delta * delta)) * u_h @ psi_h.T) - (p_h @ div_psi_h)
            integrand[:, n_u:n_dof] = div_u_h @ w_h.T + phi * p_h @ w_h.T
            el_form = el_form + det_jac[i] * omega * integrand.reshape((n_dof,))
        return el_form

    return _linearise(residual, alpha, n_dof)


def scaled_to_physical_l2_projection(space, functions, iel, alpha, alpha_scaled):
    """ScaledToPhysicalL2Projection.evaluate_form, R/src/weak_forms/scaled_to_physical_l2_projection.py:27-105
    (three-argument form: the scaled solution enters as data)."""
    dim, weights, x, det_jac, v_tab, q_tab, n_v, n_q = _mixed_setup(space, "v", "q", iel)
    n_dof = n_v + n_q
    phi_star = functions["porosity"](x[:, 0], x[:, 1], x[:, 2]) * np.ones(len(weights))
    delta_star = functions["d_phi"](x[:, 0], x[:, 1], x[:, 2]) * np.ones(len(weights))
    alpha_scaled = np.asarray(alpha_scaled, dtype=float)

    def residual(a, load):
        a = a.reshape(1, -1)
        el_form = np.zeros(n_dof)
        for i, omega in enumerate(weights):
            phi, delta = phi_star[i], delta_star[i]
            psi_h = v_tab[0, i, :, 0:dim]
            w_h = q_tab[0, i, :, 0:dim]
            u_h = a[:, 0:n_v] @ psi_h
            p_h = a[:, n_v:n_dof] @ w_h
            if load:
                v_h = alpha_scaled[0:n_v] @ psi_h
                q_h = alpha_scaled[n_v:n_dof] @ w_h
                if phi > 0.0:
                    v_h = v_h * delta
                    q_h = q_h * (1.0 / np.sqrt(phi))
            else:
                v_h, q_h = 0.0 * (alpha_scaled[0:n_v] @ psi_h), 0.0 * (alpha_scaled[n_v:n_dof] @ w_h)
            integrand = np.zeros((1, n_dof))
            integrand[:, 0:n_v] = (u_h - v_h) @ psi_h.T
            integrand[:, n_v:n_dof] = (p_h - q_h) @ w_h.T
            el_form = el_form + det_jac[i] * omega * integrand.reshape((n_dof,))
        return el_form

    return _linearise(residual, alpha, n_dof)


def _hdiv_natural_bc(space, flux, iel, value_at):
    """Shared body of the natural boundary forms of the H(div) family: ``int_F g (psi . n)`` on the facet dofs of
    the first neighbour (degenerate_elliptic_weak_form.py:147-216, unscaled_elliptic_weak_form.py:156-218)."""
    from . import ref_hdiv

    f_space, el, weights, x, det_jac, _ = _bc_setup(space, flux, iel)
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
    res = np.zeros(len(dofs))
    for i, omega in enumerate(weights):
        res += det_jac[i] * omega * value_at(x[i, 0], x[i, 1], x[i, 2]) * (tr_tab[0, i, dofs, 0:dim] @ n[0:dim])
    return res, np.zeros((len(res), len(res)))


def degenerate_elliptic_bc_dirichlet(space, functions, iel, alpha):
    def value_at(xx, yy, zz):
        q_d = np.array(functions["q"](xx, yy, zz), dtype=float)
        d_phi = functions["d_phi"](xx, yy, zz)
        phi_v = functions["porosity"](xx, yy, zz)
        if not np.isclose(phi_v, 0.0) and phi_v > 0.0:
            q_d = q_d * (d_phi / np.sqrt(phi_v))
        return q_d[0]

    return _hdiv_natural_bc(space, "v", iel, value_at)


def unscaled_elliptic_bc_dirichlet(space, functions, iel, alpha):
    return _hdiv_natural_bc(space, "u", iel, lambda xx, yy, zz: functions["p"](xx, yy, zz)[0])
