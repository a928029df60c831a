"""Shared host side of the generalised H(div) x L2 forms (``nfa_local_hdiv_l2_general``): tables, orientation cases,
per-point coefficient arrays built on the host from the user's callables (with the reference's per-point branches
applied there), kernel launch."""
import numpy as np
import torch

from .. import _lib, coefficients
from ..basis.orientation import case_matrices
from .laplace_dual_weak_form import LaplaceDualWeakFormBCDirichlet, _all_cases
from .weak_form import BatchedWeakForm, FormData, take_rows


class HdivL2GeneralForm(BatchedWeakForm):
    """Concrete forms set ``flux_name`` / ``scalar_name`` and implement ``point_coefficients(x) -> dict`` with keys
    ``c_uu`` (required), ``c_d``, ``g``, ``c_pp``, ``load_p``, ``s_u``, ``s_p`` (numpy ``[n, n_q]``; g ``[n, n_q, dim]``)."""

    form_kind = "form"
    flux_name, scalar_name = "u", "p"
    uses_alpha = True

    def _fields(self):
        names = list(self.space.discrete_spaces.keys())
        if names[:2] != [self.flux_name, self.scalar_name]:
            raise _lib.NfaError("%s expects the fields %s, %s in this order" % (type(self).__name__, self.flux_name, self.scalar_name))
        u, p = self.space.discrete_spaces[self.flux_name], self.space.discrete_spaces[self.scalar_name]
        if u.family not in ("RT", "BDM") or p.family != "Lagrange" or u.n_comp != 1 or p.n_comp != 1:
            raise _lib.NfaError("%s needs an RT/BDM flux and a scalar Lagrange field" % type(self).__name__)
        return u, p

    def n_items_total(self):
        return len(self.space.discrete_spaces[self.flux_name].elements)

    def dest_table(self):
        return self.space.destination_table()

    def point_coefficients(self, x, dim):
        raise NotImplementedError

    def prepare(self, dev, items=None, arrays=None):
        self._check()
        u, p = self._fields()
        ref, dim = u.ref_element, u.dimension
        nfd = len(ref.entity_dofs[dim - 1][0])
        points, weights = self.space.quadrature[dim]
        tab = ref.tabulate(1, points)
        up = lambda a, dt=np.float64: torch.as_tensor(np.ascontiguousarray(a, dtype=dt)).to(dev.device)
        cell_nodes = take_rows(dev.cell_nodes(self.flux_name), items)
        cases = take_rows(torch.as_tensor(_all_cases(u)).to(dev.device), items)
        x = coefficients.physical_points(dev, dim, points, cell_nodes)
        co = self.point_coefficients(x, dim)
        dev_co = {k: (None if v is None else up(v)) for k, v in co.items()}
        for k in ("c_d", "g", "c_pp", "load_p", "s_u", "s_p"):
            dev_co.setdefault(k, None)
        return FormData(
            n_items=int(cell_nodes.shape[0]), items=items, dim=dim, n_q=len(weights), n_u=ref.n_phi,
            n_p=p.ref_element.n_phi, nfd=nfd, u_val=up(tab[0]), u_div=up(sum(tab[1 + d][:, :, d] for d in range(dim))),
            p_val=up(p.ref_element.tabulate(0, points)[0, :, :, 0]), w=up(weights), tcases=up(case_matrices(ref)),
            cases=cases, cell_nodes=cell_nodes, co=dev_co,
        )

    def row_spec(self, data):
        """Row-owner path (``assembly/hdiv_rows.py``, ``csrc/nfa_hdivrows.cu`` form 2) for RT_1 x P_0 -- config 3's other
        form classes at k = 1; ``None`` for every other element and for forms with a second state vector."""
        from ..assembly import hdiv_rows

        u = self.space.discrete_spaces[self.flux_name]
        c = data.co
        if u.family != "RT" or data.nfd != 1 or data.n_u != data.dim + 1 or data.n_p != 1 or data.n_q > 64 \
                or c["s_u"] is not None or c["s_p"] is not None or getattr(self, "alpha_scaled", None) is not None \
                or list(self.space.discrete_spaces.keys()) != [self.flux_name, self.scalar_name]:
            return None
        emap, ridx, stride = hdiv_rows.general_maps(data.n_u)
        return {"emap": emap, "ridx": ridx, "stride": stride}

    def launch_records(self, dev, data, alpha_dev, alpha_prev_dev, dest_dev, rec, stride):
        c, nq = data.co, data.n_q
        arr = lambda t: (_lib.ptr(t), nq if t is not None else 0, 1 if t is not None else 0)
        if alpha_dev is not None and not self.uses_alpha:
            alpha_dev = None
        _lib.call(
            "nfa_hdiv_rt1_records", data.dim, 2, nq, _lib.ptr(data.u_val), _lib.ptr(data.u_div), _lib.ptr(data.p_val),
            _lib.ptr(data.w), _lib.ptr(data.tcases), _lib.ptr(data.cases), _lib.ptr(dev.coords), _lib.ptr(data.cell_nodes),
            data.n_items, *arr(c["c_uu"]), *arr(c["c_d"]), *arr(c["c_pp"]), 0.0, _lib.ptr(c["load_p"]), _lib.ptr(c["g"]),
            _lib.ptr(alpha_dev), None, _lib.ptr(dest_dev), stride, _lib.ptr(rec), _lib.stream_ptr(),
        )

    def launch_local(self, dev, data, alpha_dev, dest_dev, Ke, re, beta_dev=None):
        c = data.co
        need_dest = alpha_dev is not None or beta_dev is not None
        if need_dest and dest_dev is None:
            raise _lib.NfaError("%s: alpha given without the destination table" % type(self).__name__)
        _lib.call(
            "nfa_local_hdiv_l2_general", data.dim, data.n_u, data.n_p, data.n_q, data.nfd, _lib.ptr(data.u_val),
            _lib.ptr(data.u_div), _lib.ptr(data.p_val), _lib.ptr(data.w), _lib.ptr(data.tcases), _lib.ptr(data.cases),
            _lib.ptr(dev.coords), _lib.ptr(data.cell_nodes), data.n_items, _lib.ptr(c["c_uu"]), _lib.ptr(c["c_d"]),
            _lib.ptr(c["g"]), _lib.ptr(c["c_pp"]), _lib.ptr(c["load_p"]), _lib.ptr(alpha_dev if self.uses_alpha else None),
            _lib.ptr(beta_dev), _lib.ptr(dest_dev), _lib.ptr(c["s_u"]), _lib.ptr(c["s_p"]), _lib.ptr(Ke), _lib.ptr(re),
            _lib.stream_ptr(),
        )


class _RenamedView:
    """Presents a (flux, scalar) space under the field names (u, p) the LaplaceDual boundary form reads."""

    def __init__(self, space, flux_name, scalar_name):
        self._space = space
        self.discrete_spaces = {"u": space.discrete_spaces[flux_name], "p": space.discrete_spaces[scalar_name]}
        self.quadrature = space.quadrature
        self.bc_quadrature = space.bc_quadrature

    def __getattr__(self, name):
        return getattr(self._space, name)


class HdivNaturalBC(BatchedWeakForm):
    """``r = int_F g (psi . n)`` on the facet dofs of the first neighbour, ``g`` built by ``boundary_values(x)``."""

    form_kind = "bc_form"
    flux_name, scalar_name = "u", "p"

    def n_items_total(self):
        return len(self.space.discrete_spaces[self.flux_name].bc_elements)

    def dest_table(self):
        return self.space.bc_destination_table()

    def boundary_values(self, x, y, z):
        raise NotImplementedError

    def prepare(self, dev, items=None, arrays=None):
        if self.space is None or self.functions is None:
            raise ValueError("weak form needs .space and .functions")
        inner = LaplaceDualWeakFormBCDirichlet(_RenamedView(self.space, self.flux_name, self.scalar_name))
        inner.functions = {"p": self.boundary_values}
        data = inner.prepare(dev, items)
        data.inner = inner
        return data

    def launch_local(self, dev, data, alpha_dev, dest_dev, Ke, re):
        data.inner.launch_local(dev, data, None, None, Ke, re)
