"""``L2ProjectorWeakForm`` -- GPU version of R/src/weak_forms/l2_projector_weak_form.py:8-93 (same name; ``functions``
holds one callable per field of the product space).

Per field: mass matrix and load, ``J_f = int phi phi^T`` per component (Lagrange) or ``int psi . psi`` (RT / BDM,
mapped with the contravariant Piola map and the orientation transformations; N1E of degree 1 with the covariant map
and the edge-orientation signs), ``r_f = -int f . phi``; the local system
is block diagonal over the fields and is inserted with its zero off-diagonal blocks (the reference scatters the whole
``n_loc x n_loc`` matrix).  Solving ``J alpha = -r`` gives the L2 projection, the only assembly the reference's own
tests exercise (R/tests/test_projectors/test_projectors.py:86-216).
"""
import numpy as np
import torch

from .. import _lib, coefficients
from ..basis.orientation import case_matrices, edge_dof_signs
from .laplace_dual_weak_form import _all_cases
from .weak_form import BatchedWeakForm, FormData, take_rows


class L2ProjectorWeakForm(BatchedWeakForm):
    form_kind = "form"

    def n_items_total(self):
        return len(next(iter(self.space.discrete_spaces.values())).elements)

    def dest_table(self):
        return self.space.destination_table()

    def prepare(self, dev, items=None, arrays=None):
        self._check()
        parts = []
        up = lambda a, dt=np.float64: torch.as_tensor(np.ascontiguousarray(a, dtype=dt)).to(dev.device)
        for name, s in self.space.discrete_spaces.items():
            dim = s.dimension
            points, weights = self.space.quadrature[dim]
            ref = s.ref_element
            cell_nodes = take_rows(dev.cell_nodes(name), items)
            x = coefficients.physical_points(dev, dim, points, cell_nodes)
            n, nq = int(cell_nodes.shape[0]), len(weights)
            fn = self.functions[name]
            if s.family == "Lagrange":
                f = coefficients.evaluate(fn, x, s.n_comp)
                parts.append(dict(
                    kind="h1", dim=dim, n_phi=ref.n_phi, n_comp=s.n_comp, n_q=nq, n_loc=ref.n_phi * s.n_comp,
                    tab0=up(ref.tabulate(0, points)[0, :, :, 0]), w=up(weights), cell_nodes=cell_nodes, f=up(f)))
            elif s.family in ("RT", "BDM"):
                if s.n_comp != 1:
                    raise NotImplementedError("L2ProjectorWeakForm: multi-component H(div) fields")
                # the exact function returns [n_comp, 3, n_points] (R/...:54): component 0, first dim coordinates
                flat = x.reshape(-1, 3)
                val = np.asarray(fn(flat[:, 0], flat[:, 1], flat[:, 2]), dtype=np.float64)
                f_u = np.ascontiguousarray(val.reshape(-1, 3, n * nq)[0, :dim, :].T).reshape(n, nq, dim)
                tab = ref.tabulate(1, points)
                nfd = len(ref.entity_dofs[dim - 1][0])
                cases = take_rows(torch.as_tensor(_all_cases(s)).to(dev.device), items)
                parts.append(dict(
                    kind="hdiv", dim=dim, n_u=ref.n_phi, n_q=nq, nfd=nfd, n_loc=ref.n_phi, u_val=up(tab[0]),
                    u_div=up(sum(tab[1 + d][:, :, d] for d in range(dim))), p_val=up(np.zeros((nq, 1))), w=up(weights),
                    tcases=up(case_matrices(ref)), cases=cases, cell_nodes=cell_nodes, one=up(np.ones(1)), f_u=up(f_u)))
            elif s.family in ("N1E", "N2E"):
                # H(curl): covariant Piola map + edge-orientation signs (csrc/nfa_hcurl.cu); the exact function returns
                # [n_comp, 3, n_points] like the H(div) ones
                if s.n_comp != 1 or s.k_order != 1:
                    raise NotImplementedError("L2ProjectorWeakForm: N1E / N2E fields are implemented for degree 1, one component")
                flat = x.reshape(-1, 3)
                val = np.asarray(fn(flat[:, 0], flat[:, 1], flat[:, 2]), dtype=np.float64)
                f_u = np.ascontiguousarray(val.reshape(-1, 3, n * nq)[0].T).reshape(n, nq, 3)
                sel = s.mesh_topology.selected_local if items is None else s.mesh_topology.selected_local[np.asarray(items)]
                parts.append(dict(
                    kind="hcurl", dim=dim, n_u=ref.n_phi, n_q=nq, n_loc=ref.n_phi, u_val=up(ref.tabulate(0, points)[0]),
                    w=up(weights), signs=up(edge_dof_signs(s.mesh, dim, sel, ref), np.int8), cell_nodes=cell_nodes, f_u=up(f_u)))
            else:
                raise NotImplementedError("L2ProjectorWeakForm: family %r" % s.family)
        return FormData(n_items=int(parts[0]["cell_nodes"].shape[0]), items=items, parts=parts)

    def launch_local(self, dev, data, alpha_dev, dest_dev, Ke, re):
        n = data.n_items
        if Ke is not None:
            Ke.zero_()
        off = 0
        for p in data.parts:
            nl = p["n_loc"]
            kb = torch.empty((n, nl, nl), dtype=torch.float64, device=dev.device)
            rb = torch.empty((n, nl), dtype=torch.float64, device=dev.device)
            if p["kind"] == "h1":
                _lib.call(
                    "nfa_local_h1_facet", p["dim"], p["n_phi"], p["n_q"], p["n_comp"], _lib.ptr(p["tab0"]),
                    _lib.ptr(p["w"]), _lib.ptr(dev.coords), _lib.ptr(p["cell_nodes"]), n, _lib.ptr(p["f"]), None, 1.0,
                    -1.0, 1, _lib.ptr(kb), _lib.ptr(rb), _lib.stream_ptr())
            elif p["kind"] == "hcurl":
                _lib.call(
                    "nfa_local_hcurl_mass", p["dim"], p["n_u"], p["n_q"], _lib.ptr(p["u_val"]), _lib.ptr(p["w"]),
                    _lib.ptr(p["signs"]), _lib.ptr(dev.coords), _lib.ptr(p["cell_nodes"]), n, _lib.ptr(p["f_u"]), -1.0,
                    _lib.ptr(kb), _lib.ptr(rb), _lib.stream_ptr())
            else:
                _lib.call(
                    "nfa_local_hdiv_l2", p["dim"], p["n_u"], 0, p["n_q"], p["nfd"], _lib.ptr(p["u_val"]),
                    _lib.ptr(p["u_div"]), _lib.ptr(p["p_val"]), _lib.ptr(p["w"]), _lib.ptr(p["tcases"]),
                    _lib.ptr(p["cases"]), _lib.ptr(dev.coords), _lib.ptr(p["cell_nodes"]), n, _lib.ptr(p["one"]), 0, 0,
                    None, _lib.ptr(p["f_u"]), _lib.ptr(kb), _lib.ptr(rb), _lib.stream_ptr())
            if Ke is not None:
                Ke[:, off : off + nl, off : off + nl] = kb
            re[:, off : off + nl] = rb
            off += nl

    def evaluate_form(self, element_index, alpha):
        return super().evaluate_form(element_index, None)  # the projector form does not depend on alpha
