"""Two-flux mixed transport with reaction (config 5) -- GPU versions of the example-local forms of
R/examples/parabolic_diffusive_problem/SundusWeakForm.py (same class names, same ``functions`` keys).

* ``SundusDualWeakForm``            R/...:11-177   fields mp, mc (H(div)), p, c (L2); implicit Euler in time; reaction
  ``R = 1 + eta c^2`` makes the form nonlinear in c, so residual AND tangent depend on the current iterate.
  ``evaluate_form(i, alpha_n_p_1, alpha_n, t)`` differentiates with respect to its FIRST vector.
* ``SundusDualWeakFormBCDirichlet`` R/...:180-269  natural conditions ``int p_D (psi . n)`` on the mp facet dofs and
  ``int c_D (psi . n)`` on the mc facet dofs, no matrix.

``functions``: ``rhs_f(x,y,z,t)``, ``rhs_r(x,y,z,t)``, ``kappa``, ``delta``, ``eta`` ``(x,y,z)``, ``delta_t`` (a number);
boundary: ``p(x,y,z,t)``, ``c(x,y,z,t)``.
"""
import numpy as np
import torch

from .. import _lib, coefficients
from ..basis.orientation import case_matrices
from .laplace_dual_weak_form import LaplaceDualWeakFormBCDirichlet, _all_cases
from .weak_form import BatchedWeakForm, FormData, take_rows

_FIELDS = ["mp", "mc", "p", "c"]


def _tables(space):
    if list(space.discrete_spaces.keys()) != _FIELDS:
        raise _lib.NfaError("Sundus forms expect the fields mp, mc, p, c in this order")
    mp, mc, p, c = (space.discrete_spaces[n] for n in _FIELDS)
    if (mp.family, mp.k_order) != (mc.family, mc.k_order) or (p.family, p.k_order) != (c.family, c.k_order):
        raise _lib.NfaError("Sundus forms: mp/mc and p/c must use the same elements")
    if mp.family not in ("RT", "BDM") or p.family != "Lagrange" or any(s.n_comp != 1 for s in (mp, mc, p, c)):
        raise _lib.NfaError("Sundus forms need scalar-component RT/BDM fluxes and Lagrange scalars")
    ref = mp.ref_element
    dim = mp.dimension
    nfd = len(ref.entity_dofs[dim - 1][0])
    return mp, p, ref, dim, nfd


class SundusDualWeakForm(BatchedWeakForm):
    form_kind = "time_dependent_form"

    def n_items_total(self):
        return len(self.space.discrete_spaces["mp"].elements)

    def dest_table(self):
        return self.space.destination_table()

    supports_reuse = True

    def prepare(self, dev, items=None, arrays=None, t=0.0, reuse=None):
        """``reuse``: the ``FormData`` of the same form / sequence / functions at another time -- tables, connectivity, the
        mapped quadrature points and the time-independent coefficients are taken from it and only the loads
        ``rhs_f(x, y, z, t)``, ``rhs_r(x, y, z, t)`` are evaluated again (the time loop of
        R/examples/parabolic_diffusive_problem: one ``scatter_forms`` per Newton iteration, t fixed within a step)."""
        self._check()
        up = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev.device)

        def loads_at(x):
            out = {}
            for name in ("rhs_f", "rhs_r"):
                f = coefficients.evaluate(self.functions[name], x, 1, extra=(t,))[:, :, 0]
                out[name] = up(f) if np.any(f) else None
            return out

        if reuse is not None:
            data = FormData(**reuse.__dict__)
            data.loads = loads_at(reuse.x_host)
            data.t = t
            data.delta_t = float(self.functions["delta_t"])
            return data
        mp, p, ref, dim, nfd = _tables(self.space)
        points, weights = self.space.quadrature[dim]
        key = ("sundus", mp.family, mp.k_order, p.k_order, dim, len(weights))
        tab = ref.tabulate(1, points)
        u_val = dev.tensor(key + ("uval",), lambda: tab[0])
        u_div = dev.tensor(key + ("udiv",), lambda: sum(tab[1 + d][:, :, d] for d in range(dim)))
        p_val = dev.tensor(key + ("pval",), lambda: p.ref_element.tabulate(0, points)[0, :, :, 0])
        w = dev.tensor(key + ("w",), lambda: weights)
        tcases = dev.tensor(key + ("T",), lambda: case_matrices(ref))
        cases_all = dev._get(key + ("cases",), lambda: torch.as_tensor(_all_cases(mp)).to(dev.device))
        cases = take_rows(cases_all, items)
        cell_nodes = take_rows(dev.cell_nodes("mp"), items)
        n = int(cell_nodes.shape[0])
        x = coefficients.physical_points(dev, dim, points, cell_nodes)
        coefs = {}
        for name in ("kappa", "delta", "eta"):
            arr, sc, sq, kind = coefficients.classify(coefficients.evaluate(self.functions[name], x))
            coefs[name] = (up(arr), sc, sq, kind)
        return FormData(
            n_items=n, items=items, dim=dim, n_q=len(weights), n_u=ref.n_phi, n_p=p.ref_element.n_phi, nfd=nfd,
            u_val=u_val, u_div=u_div, p_val=p_val, w=w, tcases=tcases, cases=cases, cell_nodes=cell_nodes, coefs=coefs,
            loads=loads_at(x), delta_t=float(self.functions["delta_t"]), t=t, x_host=x,
        )

    def launch_local(self, dev, data, alpha_dev, dest_dev, Ke, re, alpha_prev_dev=None):
        if dest_dev is None:
            raise _lib.NfaError("SundusDualWeakForm needs the destination table on the device")
        k, d, e = data.coefs["kappa"], data.coefs["delta"], data.coefs["eta"]
        _lib.call(
            "nfa_local_sundus", data.dim, data.n_u, data.n_p, data.n_q, data.nfd, _lib.ptr(data.u_val),
            _lib.ptr(data.u_div), _lib.ptr(data.p_val), _lib.ptr(data.w), _lib.ptr(data.tcases), _lib.ptr(data.cases),
            _lib.ptr(dev.coords), _lib.ptr(data.cell_nodes), data.n_items, _lib.ptr(k[0]), k[1], k[2], _lib.ptr(d[0]),
            d[1], d[2], _lib.ptr(e[0]), e[1], e[2], data.delta_t, _lib.ptr(data.loads["rhs_f"]),
            _lib.ptr(data.loads["rhs_r"]), _lib.ptr(alpha_dev), _lib.ptr(alpha_prev_dev), _lib.ptr(dest_dev),
            _lib.ptr(Ke), _lib.ptr(re), _lib.stream_ptr(),
        )

    def row_spec(self, data):
        """Row-owner path (``assembly/hdiv_rows.py``, ``csrc/nfa_hdivrows.cu``) for the lowest-order pair RT_1 x P_0 --
        BASELINE config 5's element: record layout and entry maps, or ``None`` for every other element."""
        from ..assembly import hdiv_rows

        mp = self.space.discrete_spaces["mp"]
        if mp.family != "RT" or data.nfd != 1 or data.n_u != data.dim + 1 or data.n_p != 1 or data.n_q > 64:
            return None
        emap, ridx, stride = hdiv_rows.sundus_maps(data.n_u)
        return {"emap": emap, "ridx": ridx, "stride": stride}

    def launch_records(self, dev, data, alpha_dev, alpha_prev_dev, dest_dev, rec, stride):
        k, d, e = data.coefs["kappa"], data.coefs["delta"], data.coefs["eta"]
        _lib.call(
            "nfa_hdiv_rt1_records", data.dim, 1, data.n_q, _lib.ptr(data.u_val), _lib.ptr(data.u_div), _lib.ptr(data.p_val),
            _lib.ptr(data.w), _lib.ptr(data.tcases), _lib.ptr(data.cases), _lib.ptr(dev.coords), _lib.ptr(data.cell_nodes),
            data.n_items, _lib.ptr(k[0]), k[1], k[2], _lib.ptr(d[0]), d[1], d[2], _lib.ptr(e[0]), e[1], e[2], data.delta_t,
            _lib.ptr(data.loads["rhs_f"]), _lib.ptr(data.loads["rhs_r"]), _lib.ptr(alpha_dev), _lib.ptr(alpha_prev_dev),
            _lib.ptr(dest_dev), stride, _lib.ptr(rec), _lib.stream_ptr(),
        )

    def evaluate_form(self, element_index, alpha_n_p_1, alpha_n, t):
        """Single-cell evaluation through the same kernel (local vectors in, ``(r_el, j_el)`` out)."""
        from ..device import space_device

        self._check()
        dev = space_device(self.space)
        data = self.prepare(dev, np.array([int(element_index)], dtype=np.int64), t=t)
        nl = self.n_loc()
        up = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64).reshape(-1)).to(dev.device)
        dest = torch.arange(nl, dtype=torch.int32, device=dev.device).reshape(1, nl)
        Ke = torch.empty((1, nl, nl), dtype=torch.float64, device=dev.device)
        re = torch.empty((1, nl), dtype=torch.float64, device=dev.device)
        self.launch_local(dev, data, up(alpha_n_p_1), dest, Ke, re, up(alpha_n))
        return re[0].cpu().numpy(), Ke[0].cpu().numpy()


class SundusDualWeakFormBCDirichlet(BatchedWeakForm):
    form_kind = "time_dependent_bc_form"

    def n_items_total(self):
        return len(self.space.discrete_spaces["mp"].bc_elements)

    def dest_table(self):
        return self.space.bc_destination_table()

    supports_reuse = True

    def prepare(self, dev, items=None, arrays=None, t=0.0, reuse=None):
        """``reuse``: the ``FormData`` of the same form / sequence / functions at another time -- the facet geometry, trace
        tables and mapped boundary points are taken from it and only the boundary data ``p(x, y, z, t)``, ``c(x, y, z, t)``
        are evaluated again (the facet set-up walks every boundary facet on the host)."""
        if self.space is None or self.functions is None:
            raise ValueError("weak form needs .space and .functions")
        if reuse is not None:
            parts = []
            for key, old in zip(("p", "c"), reuse.parts):
                part = FormData(**old.__dict__)
                pd = coefficients.evaluate(self.functions[key], old.x, 1, extra=(t,))[:, :, 0]
                part.pd = torch.as_tensor(np.ascontiguousarray(pd, dtype=np.float64)).to(dev.device)
                parts.append(part)
            return FormData(n_items=reuse.n_items, items=items, parts=parts, inner=reuse.inner, nfd=reuse.nfd)
        _tables(self.space)
        mp, mc = self.space.discrete_spaces["mp"], self.space.discrete_spaces["mc"]
        if not np.array_equal(mp.bc_element_ids, mc.bc_element_ids):
            raise _lib.NfaError("Sundus boundary form: mp and mc must carry the same boundary elements")
        # the geometry / trace tables are those of the natural condition of one H(div) field; evaluate them once
        # through the LaplaceDual boundary form on a view of the space that names the flux field "u"
        view = _FluxView(self.space)
        inner = LaplaceDualWeakFormBCDirichlet(view)
        datas = []
        for key in ("p", "c"):
            fn = self.functions[key]
            inner.functions = {"p": (lambda x, y, z, fn=fn: fn(x, y, z, t))}
            datas.append(inner.prepare(dev, items))
        return FormData(n_items=datas[0].n_items, items=items, parts=datas, inner=inner, nfd=datas[0].nfd)

    def launch_local(self, dev, data, alpha_dev, dest_dev, Ke, re, alpha_prev_dev=None):
        if Ke is not None:
            Ke.zero_()
        nfd = data.nfd
        d = data.parts[0]
        common = (_lib.ptr(d.trace_tab), _lib.ptr(d.tidx), _lib.ptr(d.w), _lib.ptr(d.tcases), _lib.ptr(d.cases),
                  _lib.ptr(dev.coords), _lib.ptr(d.facet_nodes), _lib.ptr(d.neigh_nodes), _lib.ptr(d.normals))
        if nfd == 1:
            # lowest order: local order [mp_f, mc_f] = one facet function with the two boundary data as components -> one
            # launch on the interleaved data instead of two launches + two strided copies
            if data.get("pd2") is None:
                data.pd2 = torch.stack([p.pd for p in data.parts], dim=2).contiguous()
            _lib.call("nfa_local_hdiv_facet_vec", d.dim, 1, d.n_q, 2, *common, _lib.ptr(data.pd2), data.n_items, 1.0, 0, 2,
                      _lib.ptr(re), _lib.stream_ptr())
            return
        for k, part in enumerate(data.parts):  # the field's block of the local vector written in place
            _lib.call("nfa_local_hdiv_facet_vec", d.dim, nfd, d.n_q, 1, *common, _lib.ptr(part.pd), data.n_items, 1.0, k * nfd,
                      2 * nfd, _lib.ptr(re), _lib.stream_ptr())


class _FluxView:
    """Presents the (mp, p) pair of a Sundus space under the names (u, p) the LaplaceDual boundary form reads."""

    def __init__(self, space):
        self._space = space
        self.discrete_spaces = {"u": space.discrete_spaces["mp"], "p": space.discrete_spaces["p"]}
        self.quadrature = space.quadrature
        self.bc_quadrature = space.bc_quadrature

    def __getattr__(self, name):
        return getattr(self._space, name)
