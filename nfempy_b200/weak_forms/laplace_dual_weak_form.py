"""Mixed (dual) Poisson on ``u`` in H(div) (RT / BDM) x ``p`` in L2 (discontinuous Lagrange) -- GPU versions of the
classes in R/src/weak_forms/laplace_dual_weak_form.py (same names, ``functions`` keys ``rhs``, ``kappa`` / ``p``).

* ``LaplaceDualWeakForm``            R/...:10-87    J = [[int kappa^-1 psi.psi, -int w div psi], [int div psi w, 0]],
  r = [0, -int f w]  (the reference zeroes ``alpha`` before differentiating, :50, so ``alpha`` is ignored);
* ``LaplaceDualWeakFormBCDirichlet`` R/...:90-154   natural condition: r = int_F p_D (psi . n) on the facet dofs of the
  first neighbour cell, no matrix.

Local matrices go through the two-stage path (local kernel -> deterministic reduce).
"""
import itertools

import numpy as np
import torch

from .. import _lib, coefficients
from ..basis.orientation import case_matrices, facet_cases
from ..basis.reference_element import REF_VERTICES, SUB_ENTITIES
from .weak_form import BatchedWeakForm, FormData, take_rows


def _hdiv_tables(space):
    u = space.discrete_spaces["u"]
    p = space.discrete_spaces["p"]
    if list(space.discrete_spaces.keys())[:2] != ["u", "p"]:
        raise _lib.NfaError("LaplaceDual forms expect the fields in the order u, p")
    if u.family not in ("RT", "BDM") or p.family != "Lagrange" or u.n_comp != 1 or p.n_comp != 1:
        raise _lib.NfaError("LaplaceDual forms need u in RT/BDM and a scalar Lagrange p")
    ref = u.ref_element
    dim = u.dimension
    nfd = len(ref.entity_dofs[dim - 1][0])
    for f, dofs in enumerate(ref.entity_dofs[dim - 1]):
        if list(dofs) != list(range(f * nfd, (f + 1) * nfd)):
            raise _lib.NfaError("H(div) element: facet dofs are not facet-major")
    return u, p, ref, dim, nfd


def _all_cases(space_u):
    if not hasattr(space_u, "_facet_cases"):
        topo = space_u.mesh_topology
        space_u._facet_cases = facet_cases(space_u.mesh, space_u.dimension, topo.selected_local)
    return space_u._facet_cases


class LaplaceDualWeakForm(BatchedWeakForm):
    form_kind = "form"

    def n_items_total(self):
        return len(self.space.discrete_spaces["u"].elements)

    def dest_table(self):
        return self.space.destination_table()

    def prepare(self, dev, items=None, arrays=None):
        self._check()
        u, p, ref, dim, nfd = _hdiv_tables(self.space)
        points, weights = self.space.quadrature[dim]
        key = ("hdivl2", u.family, u.k_order, p.k_order, dim, len(weights))

        def u_tabs():
            tab = ref.tabulate(1, points)
            return tab

        u_val = dev.tensor(key + ("uval",), lambda: u_tabs()[0])
        u_div = dev.tensor(key + ("udiv",), lambda: sum(u_tabs()[1 + d][:, :, d] for d in range(dim)))
        p_val = dev.tensor(key + ("pval",), lambda: p.ref_element.tabulate(0, points)[0, :, :, 0])
        w = dev.tensor(key + ("w",), lambda: weights)
        tcases = dev.tensor(key + ("T",), lambda: case_matrices(ref))
        cases_all = dev._get(key + ("cases",), lambda: torch.as_tensor(_all_cases(u)).to(dev.device))
        cases = take_rows(cases_all, items)
        cell_nodes = take_rows(dev.cell_nodes("u"), items)
        n, nq = int(cell_nodes.shape[0]), len(weights)
        if arrays is None:
            x = coefficients.physical_points(dev, dim, points, cell_nodes)
            kappa = coefficients.classify(coefficients.evaluate(self.functions["kappa"], x))
            f = coefficients.evaluate(self.functions["rhs"], x, 1)[:, :, 0]
        else:
            v = np.asarray(arrays["kappa"], dtype=np.float64)
            kappa = (v.reshape(1), 0, 0, "constant") if v.size == 1 else (
                (np.ascontiguousarray(v), nq, 1, "point") if v.ndim == 2 else (np.ascontiguousarray(v), 1, 0, "cell"))
            f = arrays.get("rhs")
        if f is not None and not np.any(f):
            f = None
        return FormData(
            n_items=n, items=items, dim=dim, n_q=nq, n_u=ref.n_phi, n_p=p.ref_element.n_phi, nfd=nfd, u_val=u_val,
            u_div=u_div, p_val=p_val, w=w, tcases=tcases, cases=cases, cell_nodes=cell_nodes,
            kappa=(torch.as_tensor(kappa[0]).to(dev.device), kappa[1], kappa[2], kappa[3]),
            f=None if f is None else torch.as_tensor(np.ascontiguousarray(f, dtype=np.float64)).to(dev.device),
        )

    def launch_local(self, dev, data, alpha_dev, dest_dev, Ke, re):
        k = data.kappa
        _lib.call(
            "nfa_local_hdiv_l2", data.dim, data.n_u, data.n_p, data.n_q, data.nfd, _lib.ptr(data.u_val),
            _lib.ptr(data.u_div), _lib.ptr(data.p_val), _lib.ptr(data.w), _lib.ptr(data.tcases), _lib.ptr(data.cases),
            _lib.ptr(dev.coords), _lib.ptr(data.cell_nodes), data.n_items, _lib.ptr(k[0]), k[1], k[2],
            _lib.ptr(data.f), None, _lib.ptr(Ke), _lib.ptr(re), _lib.stream_ptr(),
        )

    def row_spec(self, data):
        """Row-owner path (``assembly/hdiv_rows.py``, ``csrc/nfa_hdivrows.cu``) for RT_1 x P_0 -- BASELINE config 3 at
        k = 1: record layout and entry maps, or ``None`` for every other element."""
        from ..assembly import hdiv_rows

        u = self.space.discrete_spaces["u"]
        if u.family != "RT" or data.nfd != 1 or data.n_u != data.dim + 1 or data.n_p != 1 or data.n_q > 64 \
                or list(self.space.discrete_spaces.keys()) != ["u", "p"]:
            return None
        emap, ridx, stride = hdiv_rows.laplace_dual_maps(data.n_u)
        return {"emap": emap, "ridx": ridx, "stride": stride}

    def launch_records(self, dev, data, alpha_dev, alpha_prev_dev, dest_dev, rec, stride):
        k = data.kappa  # the reference discards alpha (:50): no state terms
        _lib.call(
            "nfa_hdiv_rt1_records", data.dim, 0, data.n_q, _lib.ptr(data.u_val), _lib.ptr(data.u_div), _lib.ptr(data.p_val),
            _lib.ptr(data.w), _lib.ptr(data.tcases), _lib.ptr(data.cases), _lib.ptr(dev.coords), _lib.ptr(data.cell_nodes),
            data.n_items, _lib.ptr(k[0]), k[1], k[2], None, 0, 0, None, 0, 0, 0.0, _lib.ptr(data.f), None, None, None, None,
            stride, _lib.ptr(rec), _lib.stream_ptr(),
        )

    def evaluate_form(self, element_index, alpha):
        return super().evaluate_form(element_index, None)  # the reference discards alpha (:50)


def _facet_normals(points, facet_nodes, neigh_nodes):
    """Vectorised ``normal(mesh, neighbour, facet)`` (R/src/geometry/compute_normal.py:7-24)."""
    P = points[facet_nodes]
    xc1 = P.mean(axis=1)
    xc0 = points[neigh_nodes].mean(axis=1)
    outward = xc1 - xc0
    outward /= np.linalg.norm(outward, axis=1, keepdims=True)
    if facet_nodes.shape[1] == 3:
        n = np.cross(P[:, 0] - xc1, P[:, 1] - xc1)
    else:
        v = P[:, 1] - P[:, 0]
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        w = np.cross(v, outward)
        w /= np.linalg.norm(w, axis=1, keepdims=True)
        n = np.cross(v, w)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    flip = np.einsum("ij,ij->i", n, outward) < 0.0
    n[flip] *= -1.0
    return n


def hdiv_facet_setup(space, field_name, dev, items=None):
    """Everything the natural-boundary kernels need for the boundary facets of one H(div) field: the neighbour's
    reference trace pre-tabulated per (local facet, vertex permutation) -- the vectorised form of
    ``transform_lower_to_higher`` (R/src/basis/parametric_transformation.py:58-98) --, orientation cases, normals
    (R/src/geometry/compute_normal.py:7-24) and the physical quadrature points ``x`` for the boundary data."""
    u = space.discrete_spaces[field_name]
    ref, dim = u.ref_element, u.dimension
    nfd = len(ref.entity_dofs[dim - 1][0])
    fdim = dim - 1
    points, weights = space.bc_quadrature[fdim]
    points = np.asarray(points, dtype=np.float64).reshape(len(weights), fdim)
    mesh = u.mesh
    facets = u.bc_element_ids if items is None else u.bc_element_ids[np.asarray(items)]
    rows, local_f = u.mesh_topology.first_neighbours(facets)
    neigh_nodes = u.cell_nodes()[rows]
    facet_nodes = mesh.node_tags_of(fdim)[facets - mesh.offset[fdim]]
    # local vertex (in the neighbour) of every facet node -> which pre-tabulated permutation to use
    g = (neigh_nodes[:, None, :] == facet_nodes[:, :, None]).argmax(axis=2)  # [n, fdim+1]
    perms = {}
    tabs = []
    verts = REF_VERTICES[dim]
    lam = np.concatenate([1.0 - points.sum(axis=1, keepdims=True), points], axis=1)  # [nq, fdim+1]
    for f, lv in enumerate(SUB_ENTITIES[dim][fdim]):
        for perm in itertools.permutations(lv):
            perms[(f,) + tuple(perm)] = len(tabs)
            mapped = lam @ verts[list(perm)]  # R/src/basis/parametric_transformation.py:58-98
            tabs.append(ref.tabulate(0, mapped)[0][:, f * nfd : (f + 1) * nfd, :])
    tidx = np.array([perms[(int(f),) + tuple(int(v) for v in row)] for f, row in zip(local_f, g)], dtype=np.int32)
    cases = _all_cases(u)[rows, local_f].astype(np.int8)
    normals = _facet_normals(mesh.points, facet_nodes, neigh_nodes)
    facet_nodes_dev = torch.as_tensor(np.ascontiguousarray(facet_nodes, dtype=np.int32)).to(dev.device)
    x = coefficients.physical_points(dev, fdim, points, facet_nodes_dev)
    to = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a, dtype=dt)).to(dev.device)
    return FormData(
        n_items=len(facets), items=items, dim=dim, nfd=nfd, n_q=len(weights), trace_tab=to(np.array(tabs), np.float64),
        tidx=to(tidx, np.int32), w=to(weights, np.float64), tcases=to(case_matrices(ref), np.float64),
        cases=to(cases, np.int8), facet_nodes=facet_nodes_dev, neigh_nodes=to(neigh_nodes, np.int32),
        normals=to(normals, np.float64), x=x,
    )


class LaplaceDualWeakFormBCDirichlet(BatchedWeakForm):
    form_kind = "bc_form"

    def n_items_total(self):
        return len(self.space.discrete_spaces["u"].bc_elements)

    def dest_table(self):
        return self.space.bc_destination_table()

    def prepare(self, dev, items=None, arrays=None):
        if self.space is None or self.functions is None:
            raise ValueError("weak form needs .space and .functions")
        _hdiv_tables(self.space)
        data = hdiv_facet_setup(self.space, "u", dev, items)
        if arrays is None:
            pd = coefficients.evaluate(self.functions["p"], data.x, 1)[:, :, 0]
        else:
            pd = arrays["p"]
        data.pd = torch.as_tensor(np.ascontiguousarray(pd, dtype=np.float64)).to(dev.device)
        return data

    def launch_local(self, dev, data, alpha_dev, dest_dev, Ke, re):
        if Ke is not None:
            Ke.zero_()  # j_el = 0: the entries are inserted as explicit zeros (R/...:108-110)
        _lib.call(
            "nfa_local_hdiv_facet", data.dim, data.nfd, data.n_q, _lib.ptr(data.trace_tab), _lib.ptr(data.tidx),
            _lib.ptr(data.w), _lib.ptr(data.tcases), _lib.ptr(data.cases), _lib.ptr(dev.coords),
            _lib.ptr(data.facet_nodes), _lib.ptr(data.neigh_nodes), _lib.ptr(data.normals), _lib.ptr(data.pd),
            data.n_items, _lib.ptr(re), _lib.stream_ptr(),
        )
