"""Shared machinery of the H1 (Lagrange) volume and facet forms: reference tables, coefficient upload, kernel
launches.  Concrete forms (``le_primal_weak_form``, ``laplace_primal_weak_form``) only say which field, which
coefficients and which kernel."""
import numpy as np
import torch

from .. import _lib, coefficients
from .weak_form import BatchedWeakForm, FormData, take_rows


def scalar_table(ref, points):
    """``tab[1+dim, n_q, n_phi]`` of a scalar element at ``points`` (drops the value-size axis)."""
    return np.ascontiguousarray(ref.tabulate(1, points)[:, :, :, 0])


def reference_stiffness_tensor(ref, points, weights):
    """``Mref[a, b, k, l] = int_ref d_k phi_a d_l phi_b`` by the space's quadrature (exact: degree 2k+1 >= 2k-2)."""
    tab = scalar_table(ref, points)
    return np.ascontiguousarray(np.einsum("q,kqa,lqb->abkl", weights, tab[1:], tab[1:]))


class H1VolumeForm(BatchedWeakForm):
    """A volume form on a single Lagrange field (``field_name``) with ``n_comp`` components."""

    form_kind = "form"
    field_name = None
    coefficient_names = ()  # scalar coefficients, in kernel order
    load_name = "rhs"
    load_sign = 1.0
    p2tet_kernel = None

    def _field(self):
        return self.space.discrete_spaces[self.field_name]

    def n_items_total(self):
        return len(self._field().elements)

    def dest_table(self):
        return self.space.destination_table()

    def _tables(self, dev):
        s = self._field()
        dim = s.dimension
        points, weights = self.space.quadrature[dim]
        key = (self.field_name, "h1tab", s.k_order, dim, len(weights))
        tab = dev.tensor(key + ("tab",), lambda: scalar_table(s.ref_element, points))
        w = dev.tensor(key + ("w",), lambda: weights)
        return dim, points, weights, tab, w

    def prepare(self, dev, items=None, arrays=None):
        """Evaluates the coefficient callables on the host over all quadrature points of ``items`` and uploads
        them.  ``arrays`` (name -> numpy array) bypasses the callables: scalars ``[n, n_q]`` / ``[n]`` / ``[1]``,
        load ``[n, n_q, n_comp]`` or None."""
        self._check()
        s = self._field()
        if s.family != "Lagrange":
            raise _lib.NfaError("%s needs a Lagrange field %r" % (type(self).__name__, self.field_name))
        dim, points, weights, tab, w = self._tables(dev)
        cell_nodes = take_rows(dev.cell_nodes(self.field_name), items)
        n = int(cell_nodes.shape[0])
        nq = len(weights)
        coefs = {}
        if arrays is None:
            x = coefficients.physical_points(dev, dim, points, cell_nodes)
            for name in self.coefficient_names:
                coefs[name] = coefficients.classify(coefficients.evaluate(self.functions[name], x))
            f = coefficients.evaluate(self.functions[self.load_name], x, s.n_comp)
        else:
            for name in self.coefficient_names:
                v = np.asarray(arrays[name], dtype=np.float64)
                if v.ndim == 2:
                    coefs[name] = (np.ascontiguousarray(v), nq, 1, "point")
                elif v.size == 1:
                    coefs[name] = (v.reshape(1), 0, 0, "constant")
                else:
                    coefs[name] = (np.ascontiguousarray(v.reshape(n)), 1, 0, "cell")
            f = arrays.get(self.load_name)
        if f is not None and not np.any(f):
            f = None
        dev_coefs = {
            k: (torch.as_tensor(v[0]).to(dev.device), v[1], v[2], v[3]) for k, v in coefs.items()
        }
        f_dev = None if f is None else torch.as_tensor(np.ascontiguousarray(f, dtype=np.float64)).to(dev.device)
        # the row-item kernel of vector-P2 tetrahedra scales by sqrt(det mu): it needs mu > 0 in every cell
        p2tet_ok = bool(
            self.p2tet_kernel and dim == 3 and s.ref_element.n_phi == 10 and s.n_comp == 3
            and all(c[3] != "point" for c in coefs.values()) and float(np.min(coefs["mu"][0])) > 0.0
            and np.all(np.isfinite(coefs["lambda"][0]))
        )
        return FormData(
            p2tet_ok=p2tet_ok,
            n_items=n, items=items, dim=dim, n_q=nq, n_phi=s.ref_element.n_phi, n_comp=s.n_comp, tab=tab, w=w,
            cell_nodes=cell_nodes, coefs=dev_coefs, f=f_dev,
        )

    # -- fused path --------------------------------------------------------------------------------
    def fused_spec(self, data):
        if any(c[3] == "point" for c in data.coefs.values()):
            return None  # coefficients vary inside cells: quadrature-loop kernel + two-stage reduce
        spec = {"kernel": self.fused_kernel, "coefs": [data.coefs[n] for n in self.coefficient_names]}
        if data.get("p2tet_ok"):
            spec["p2tet_kernel"] = self.p2tet_kernel
        return spec

    def mref(self, dev):
        s = self._field()
        dim = s.dimension
        points, weights = self.space.quadrature[dim]
        key = (self.field_name, "mref", s.k_order, dim)
        return dev.tensor(key, lambda: reference_stiffness_tensor(s.ref_element, points, weights))

    def launch_load(self, dev, data, re, comp_pitch=None):
        """Load vector only (fused path).  ``comp_pitch``: padded node records ``re [n, n_phi, comp_pitch]``."""
        if comp_pitch is not None:
            _lib.call(
                "nfa_local_h1_load_pitched", data.dim, data.n_phi, data.n_q, data.n_comp, _lib.ptr(data.tab[0]),
                _lib.ptr(data.w), _lib.ptr(dev.coords), _lib.ptr(data.cell_nodes), data.n_items, _lib.ptr(data.f),
                float(self.load_sign), _lib.ptr(re), int(comp_pitch), _lib.stream_ptr(),
            )
            return
        _lib.call(
            "nfa_local_h1_load", data.dim, data.n_phi, data.n_q, data.n_comp, _lib.ptr(data.tab[0]), _lib.ptr(data.w),
            _lib.ptr(dev.coords), _lib.ptr(data.cell_nodes), data.n_items, _lib.ptr(data.f), float(self.load_sign),
            _lib.ptr(re), _lib.stream_ptr(),
        )


class H1FacetForm(BatchedWeakForm):
    """Facet form ``J = beta int phi phi^T`` (per component), ``r = r_sign beta int g phi`` on the boundary
    elements of one Lagrange field."""

    form_kind = "bc_form"
    field_name = None
    data_name = None
    beta = 1.0
    r_sign = 1.0
    with_matrix = True
    normal_only = False

    def _field(self):
        return self.space.discrete_spaces[self.field_name]

    def n_items_total(self):
        return len(self._field().bc_elements)

    def dest_table(self):
        return self.space.bc_destination_table()

    def prepare(self, dev, items=None, arrays=None):
        if self.space is None or self.functions is None:
            raise ValueError("weak form needs .space and .functions")
        s = self._field()
        if len(self.space.discrete_spaces) != 1 and not getattr(self, "_multi", False):
            bc_fields = [n for n, sp in self.space.discrete_spaces.items() if len(sp.bc_elements) != 0]
            if bc_fields != [self.field_name]:
                raise _lib.NfaError("H1 facet form: only one field may carry boundary elements")
        fdim = s.dimension - 1
        ref = s.bc_ref_element
        points, weights = self.space.bc_quadrature[fdim]
        points = np.asarray(points, dtype=np.float64).reshape(-1, max(fdim, 1))
        key = (self.field_name, "h1facet", ref.degree, fdim, len(weights))
        tab0 = dev.tensor(key + ("tab0",), lambda: ref.tabulate(0, points)[0, :, :, 0])
        w = dev.tensor(key + ("w",), lambda: weights)
        facet_nodes = take_rows(dev.bc_cell_nodes(self.field_name), items)
        n = int(facet_nodes.shape[0])
        if arrays is None:
            x = coefficients.physical_points(dev, fdim, points, facet_nodes)
            g = coefficients.evaluate(self.functions[self.data_name], x, s.n_comp)
        else:
            g = arrays[self.data_name]
        comp_scale = None
        if self.normal_only:
            # LEPrimalWeakFormBCNormalDirichlet: phi n_c on both sides of the penalty, load (u_D . n) n_c
            rows, _ = s.mesh_topology.first_neighbours(s.bc_element_ids if items is None else s.bc_element_ids[np.asarray(items)])
            from .laplace_dual_weak_form import _facet_normals

            fn_host = facet_nodes.cpu().numpy().astype(np.int64)
            normals = _facet_normals(s.mesh.points, fn_host, s.cell_nodes()[rows])
            nc = s.n_comp
            un = np.einsum("nqc,nc->nq", g[:, :, :nc], normals[:, :nc]) if nc == 3 else np.einsum(
                "nqc,nc->nq", g, normals[:, :nc])
            g = un[:, :, None] * normals[:, None, :nc]
            comp_scale = torch.as_tensor(np.ascontiguousarray(normals[:, :nc] ** 2)).to(dev.device)
        g_dev = torch.as_tensor(np.ascontiguousarray(g, dtype=np.float64)).to(dev.device)
        return FormData(
            n_items=n, items=items, fdim=fdim, n_q=len(weights), n_phi=ref.n_phi, n_comp=s.n_comp, tab0=tab0, w=w,
            facet_nodes=facet_nodes, g=g_dev, comp_scale=comp_scale,
        )

    def launch_local(self, dev, data, alpha_dev, dest_dev, Ke, re):
        _lib.call(
            "nfa_local_h1_facet", data.fdim, data.n_phi, data.n_q, data.n_comp, _lib.ptr(data.tab0), _lib.ptr(data.w),
            _lib.ptr(dev.coords), _lib.ptr(data.facet_nodes), data.n_items, _lib.ptr(data.g),
            _lib.ptr(data.comp_scale), float(self.beta),
            float(self.r_sign), int(self.with_matrix), _lib.ptr(Ke), _lib.ptr(re), _lib.stream_ptr(),
        )
