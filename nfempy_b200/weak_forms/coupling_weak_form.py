"""Mixed-dimensional Robin coupling across a fracture -- GPU version of the example-local ``CouplingWeakForm``
(R/examples/md_elliptic_scalar/CouplingWeakForm.py:11-151; same name, same ``functions`` keys ``kappa_normal`` and
``delta``, same local dof order).

``space`` is the list ``[space_c0, space_c1]`` the reference passes (``CouplingWeakForm(md_produc_space)``): the
higher-dimensional mixed space (flux ``u`` in RT / BDM, potential ``p``) and the space on the fracture cells.  One item =
one fracture cell with the boundary facet of the higher-dimensional domain on its positive and on its negative side:

    r_{u+} = int ((1/alpha) u+.n+ + p_c1) (psi+.n+),   r_{u-} likewise,   r_p = int (u+.n+ + u-.n-) w,
    alpha = kappa_normal / (delta / 2),

local dofs ``[u_+ facet dofs | u_- facet dofs | p of the fracture cell]`` and destination indices
``[bc_destination_indexes(c0_p, "u"), bc_destination_indexes(c0_n, "u"), space_c1.destination_indexes(c1, "p")]`` -- the
closure ``scatter_coupling_form_data`` of R/examples/md_elliptic_scalar/md_elliptic_scalar.py:284-302; each space
carries its ``dof_shift`` (:206-209).  The form is linear: the Jacobian is state-independent and ``r = J alpha_l``.

With ``GPUAssembler`` the form is listed as ``["interface_form", triples, form, alpha]``, ``triples`` an int array
``[m, 3]`` of ``(c1_idx, c0_p_idx, c0_n_idx)`` (element index in ``space_c1``, boundary-element indices in ``space_c0``),
see ``interface_triples``.
"""
import numpy as np
import torch

from .. import _lib, coefficients
from .laplace_dual_weak_form import hdiv_facet_setup
from .weak_form import BatchedWeakForm, FormData


def interface_triples(space_c0, space_c1, interface, flux_name="u"):
    """``int64[m, 3]`` ``(c1_idx, c0_p_idx, c0_n_idx)`` of a fracture description ``{"c1", "positive", "negative"}``
    (node tags, row by row) as returned by ``mesh.fractured_box_tet_mesh``: what the reference builds from its
    ``interfaces`` with ``id_to_element`` / ``id_to_bc_element`` (R/.../md_elliptic_scalar.py:324-337)."""
    u0, u1 = space_c0.discrete_spaces[flux_name], space_c1.discrete_spaces[flux_name]
    mesh = u0.mesh
    fdim = u0.dimension - 1
    c1 = mesh.find_cell_ids(fdim, interface["c1"])
    pos = mesh.find_cell_ids(fdim, interface["positive"])
    neg = mesh.find_cell_ids(fdim, interface["negative"])
    return np.stack([
        np.array([u1.id_to_element[int(c)] for c in c1], dtype=np.int64),
        np.array([u0.id_to_bc_element[int(c)] for c in pos], dtype=np.int64),
        np.array([u0.id_to_bc_element[int(c)] for c in neg], dtype=np.int64),
    ], axis=1)


class CouplingWeakForm(BatchedWeakForm):
    form_kind = "interface_form"
    flux_name = "u"
    potential_name = "p"

    def primary_space(self):
        return self.space[0]

    def n_items_total(self):
        raise _lib.NfaError("CouplingWeakForm: list the interface triples explicitly (interface_triples)")

    def interface_dest(self, triples):
        """``int64[m, n_loc]`` destination indices of the listed triples (both spaces' ``dof_shift`` included)."""
        s0, s1 = self.space
        t = np.asarray(triples, dtype=np.int64).reshape(-1, 3)
        bc = s0.bc_destination_table(self.flux_name)
        return np.concatenate([bc[t[:, 1]], bc[t[:, 2]], s1.destination_table(self.potential_name)[t[:, 0]]], axis=1)

    def n_loc(self):
        s0, s1 = self.space
        u = s0.discrete_spaces[self.flux_name]
        nfd = len(u.ref_element.entity_dofs[u.dimension - 1][0])
        return 2 * nfd + s1.discrete_spaces[self.potential_name].ref_element.n_phi

    def prepare(self, dev, items=None, arrays=None):
        if self.space is None or self.functions is None:
            raise ValueError("weak form needs .space and .functions")
        if items is None:
            raise _lib.NfaError("CouplingWeakForm: list the interface triples explicitly (interface_triples)")
        s0, s1 = self.space
        u0, p1 = s0.discrete_spaces[self.flux_name], s1.discrete_spaces[self.potential_name]
        if u0.family not in ("RT", "BDM") or u0.n_comp != 1 or p1.n_comp != 1 or p1.family != "Lagrange":
            raise _lib.NfaError("CouplingWeakForm needs a scalar-component RT / BDM flux and a Lagrange potential")
        if p1.dimension != u0.dimension - 1:
            raise _lib.NfaError("CouplingWeakForm: the second space lives one dimension below the first")
        if u0.mesh is not p1.mesh:
            raise _lib.NfaError("CouplingWeakForm: both spaces must be built on the same mesh")
        t = np.asarray(items, dtype=np.int64).reshape(-1, 3)

        class _View:  # hdiv_facet_setup reads space.discrete_spaces[field] and space.bc_quadrature
            discrete_spaces = {"u": u0}
            bc_quadrature = s0.bc_quadrature

        side_a = hdiv_facet_setup(_View, "u", dev, t[:, 1])
        side_b = hdiv_facet_setup(_View, "u", dev, t[:, 2])
        # the fracture cell and its two clones list their vertices in the same order (the clones are copies of the cell,
        # R/src/mesh/mesh_operations.py:159-170), so one set of facet quadrature points serves all three: the reference
        # asserts exactly this (CouplingWeakForm.py:85-88)
        if not np.allclose(side_a.x, side_b.x, rtol=0.0, atol=1e-12):
            raise _lib.NfaError("CouplingWeakForm: the two sides of an interface cell do not coincide point by point")
        fdim = u0.dimension - 1
        points, weights = s0.bc_quadrature[fdim]
        points = np.asarray(points, dtype=np.float64).reshape(len(weights), fdim)
        p_val = p1.ref_element.tabulate(0, points)[0, :, :, 0]
        if arrays is None:
            kn = coefficients.evaluate(self.functions["kappa_normal"], side_a.x)
            dl = coefficients.evaluate(self.functions["delta"], side_a.x)
        else:
            kn, dl = arrays["kappa_normal"], arrays["delta"]
        up = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev.device)
        return FormData(
            n_items=len(t), items=items, dim=u0.dimension, nfd=side_a.nfd, n_p=int(p_val.shape[1]), n_q=len(weights),
            a=side_a, b=side_b, p_val=up(p_val), ainv=up(0.5 * dl / kn),
        )

    def launch_local(self, dev, data, alpha_dev, dest_dev, Ke, re):
        a, b = data.a, data.b
        _lib.call(
            "nfa_local_hdiv_coupling", data.dim, data.nfd, data.n_p, data.n_q, _lib.ptr(a.trace_tab), _lib.ptr(a.w),
            _lib.ptr(a.tcases), _lib.ptr(dev.coords), _lib.ptr(a.tidx), _lib.ptr(a.cases), _lib.ptr(a.facet_nodes),
            _lib.ptr(a.neigh_nodes), _lib.ptr(a.normals), _lib.ptr(b.tidx), _lib.ptr(b.cases), _lib.ptr(b.facet_nodes),
            _lib.ptr(b.neigh_nodes), _lib.ptr(b.normals), _lib.ptr(data.p_val), _lib.ptr(data.ainv), data.n_items,
            _lib.ptr(alpha_dev), _lib.ptr(dest_dev), _lib.ptr(Ke), _lib.ptr(re), _lib.stream_ptr(),
        )

    def evaluate_form(self, index_c1, index_c0_p, index_c0_n, alpha):
        """Single interface cell through the same kernel: local ``alpha`` in, ``(r_el, j_el)`` out."""
        from ..device import space_device

        dev = space_device(self.space[0])
        data = self.prepare(dev, np.array([[int(index_c1), int(index_c0_p), int(index_c0_n)]], dtype=np.int64))
        nl = self.n_loc()
        alpha_dev = dest_dev = None
        if alpha is not None and np.any(np.asarray(alpha) != 0.0):
            alpha_dev = torch.as_tensor(np.ascontiguousarray(alpha, dtype=np.float64).reshape(-1)).to(dev.device)
            dest_dev = torch.arange(nl, dtype=torch.int32, device=dev.device).reshape(1, nl)
        Ke = torch.empty((1, nl, nl), dtype=torch.float64, device=dev.device)
        re = torch.empty((1, nl), dtype=torch.float64, device=dev.device)
        self.launch_local(dev, data, alpha_dev, dest_dev, Ke, re)
        return re[0].cpu().numpy(), Ke[0].cpu().numpy()
