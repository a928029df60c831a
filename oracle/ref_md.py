"""ORACLE (test infrastructure only -- never imported by the product path).

Mixed-dimensional coupling, cell by cell:

* ``coupling``    -- CouplingWeakForm.evaluate_form, R/examples/md_elliptic_scalar/CouplingWeakForm.py:11-151 (traces of
  the neighbours' flux bases on the two clones of a fracture cell, Robin coupling with the fracture potential), in the
  reference's operation order; the Jacobian by pushing unit vectors through the residual (what forward AD gives for a
  form that is linear in ``alpha``);
* ``assemble_md`` -- the driver's scatter loops, R/examples/md_elliptic_scalar/md_elliptic_scalar.py:250-342: volume and
  boundary forms of every space with its ``dof_shift``, then ``scatter_coupling_form_data`` over the interface triples.

PARITY: pinned against the reference's own source above the basix boundary (tools/make_reference_pins.py runs
CouplingWeakForm + the stacked spaces over stand-ins for basix / auto_diff; tests/golden/refcode_md_elliptic_k*.npz: bit-identical);
basix's element conventions stay unpinned (oracle/ref_basis.py).
"""
import numpy as np

from . import ref_assembly, ref_hdiv
from .ref_forms import RefElement, _linearise, normal, transform_lower_to_higher


def _trace_space(md_space, iel_c0, iel_c1, field_pair=("u", "p")):
    """``compute_trace_space`` of CouplingWeakForm.py:19-69."""
    field_c0, field_c1 = field_pair
    space_c0 = md_space[0].discrete_spaces[field_c0]
    space_c1 = md_space[1].discrete_spaces[field_c1]
    el_c0 = RefElement(space_c0, space_c0.bc_element_ids[iel_c0], bc=True)
    cell_c0 = el_c0.cell
    dim = cell_c0.dimension
    points, weights = md_space[0].bc_quadrature[dim]
    x, jac, det_jac, inv_jac = el_c0.evaluate_mapping(points)
    neigh = RefElement(space_c0, space_c0.first_neighbour(cell_c0.id))
    mapped = transform_lower_to_higher(np.asarray(points).reshape(len(weights), -1), cell_c0, neigh.cell)
    _, jac0, det0, inv0 = neigh.evaluate_mapping(mapped)
    tr_tab = neigh.evaluate_basis(mapped, jac0, det0, inv0)
    facet_index = neigh.cell.sub_cells_ids[cell_c0.dimension].tolist().index(cell_c0.id)
    dof_n_index = np.array(ref_hdiv.entity_dofs(neigh.family, neigh.cell.dimension, neigh.k)[cell_c0.dimension][facet_index])
    el_c1 = RefElement(space_c1, space_c1.element_ids[iel_c1])
    p_tab = el_c1.evaluate_basis(points, jac, det_jac, inv_jac)
    n = normal(space_c0.mesh, neigh.cell, cell_c0)
    return tr_tab, dof_n_index, n, p_tab, (x, weights, det_jac)


def coupling(md_space, functions, index_c1, index_c0_p, index_c0_n, alpha, field_pair=("u", "p")):
    """``field_pair``: (flux field of the higher-dimensional space, potential field of the fracture space) -- ("u", "p")
    in the reference; the two-flux transport model couples ("mp", "p") and ("mc", "c") the same way."""
    du_p, idx_p, n_p, p_tab_p, map_p = _trace_space(md_space, index_c0_p, index_c1, field_pair)
    du_n, idx_n, n_n, p_tab_n, map_n = _trace_space(md_space, index_c0_n, index_c1, field_pair)
    assert np.all(np.isclose(p_tab_p, p_tab_n))
    for a, b in zip(map_p, map_n):
        assert np.all(np.isclose(a, b))
    p_tab = p_tab_p
    x, weights, det_jac = map_p
    dim = md_space[0].discrete_spaces[field_pair[0]].dimension
    n_u_p, n_u_n, n_pd = len(idx_p), len(idx_n), p_tab.shape[2]
    n_dof = n_u_p + n_u_n + n_pd
    s_up, s_un, s_p = slice(0, n_u_p), slice(n_u_p, n_u_p + n_u_n), slice(n_u_p + n_u_n, n_dof)

    def residual(a, load):
        a = a.reshape(1, -1)
        el_form = np.zeros(n_dof)
        for i, omega in enumerate(weights):
            kappa_n_v = functions["kappa_normal"](x[i, 0], x[i, 1], x[i, 2])
            delta_v = functions["delta"](x[i, 0], x[i, 1], x[i, 2])
            du_h_p = (du_p[0:1, i, idx_p, 0:dim] @ n_p[0:dim]).T
            du_h_n = (du_n[0:1, i, idx_n, 0:dim] @ n_n[0:dim]).T
            dp_h = p_tab[0, i, :, 0:dim]
            u_h_p = a[:, s_up] @ du_h_p
            u_h_n = a[:, s_un] @ du_h_n
            p_h_c1 = a[:, s_p] @ dp_h
            alpha_v = kappa_n_v / (delta_v / 2.0)
            integrand = np.zeros((1, n_dof))
            integrand[:, s_up] = ((1.0 / alpha_v) * u_h_p + p_h_c1) @ du_h_p.T
            integrand[:, s_un] = ((1.0 / alpha_v) * u_h_n + p_h_c1) @ du_h_n.T
            integrand[:, s_p] = u_h_p @ dp_h.T + u_h_n @ dp_h.T
            el_form = el_form + det_jac[i] * omega * integrand.reshape((n_dof,))
        return el_form

    return _linearise(residual, alpha, n_dof)


def scatter_coupling(md_space, functions, triples, res_g, jac_g, alpha, field_pair=("u", "p")):
    """``scatter_coupling_form_data`` over the interface triples, R/examples/md_elliptic_scalar/md_elliptic_scalar.py:284-342."""
    fu, fp = field_pair
    for c1_idx, c0_p_idx, c0_n_idx in triples:
        dest = np.concatenate([
            md_space[0].bc_destination_indexes(int(c0_p_idx), fu), md_space[0].bc_destination_indexes(int(c0_n_idx), fu),
            md_space[1].destination_indexes(int(c1_idx), fp),
        ])
        r_el, j_el = coupling(md_space, functions, int(c1_idx), int(c0_p_idx), int(c0_n_idx), alpha[dest], field_pair)
        ref_assembly.scatter_rhs(res_g, r_el, dest)
        ref_assembly.scatter_lhs(jac_g, j_el, dest)


def assemble_md(md_space, space_forms, triples, coupling_functions, alpha=None):
    """``space_forms[k]``: list of ``(kind, form_fn, functions, n_items)`` of space k (kinds "form" / "bc_form"; the
    spaces carry their ``dof_shift``); ``triples``: ``(c1_idx, c0_p_idx, c0_n_idx)`` rows.  Returns the CSR + rhs of the
    stacked system."""
    n = max(s.dof_shift + s.n_dof for s in md_space)
    alpha = np.zeros(n) if alpha is None else alpha
    res_g, jac_g = np.zeros(n), ref_assembly.DokMatrix(n)
    for space, forms in zip(md_space, space_forms):
        for kind, form, functions, n_items in forms:
            scatter = ref_assembly.scatter_form if kind == "form" else ref_assembly.scatter_bc_form
            for i in (range(n_items) if isinstance(n_items, int) else n_items):
                scatter(space, form, functions, i, res_g, jac_g, alpha)
    scatter_coupling(md_space, coupling_functions, triples, res_g, jac_g, alpha)
    indptr, indices, data = jac_g.getValuesCSR()
    return indptr, indices, data, res_g
