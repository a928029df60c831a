"""Weakly-symmetric mixed elasticity / Cosserat forms on row-wise H(div) stresses -- GPU versions of

* ``LEDualWeakForm``            R/src/weak_forms/le_dual_weak_form.py:12-510 (fields s, u, t; functions rhs, lambda, mu)
* ``LEDualWeakFormBCDirichlet`` R/src/weak_forms/le_dual_weak_form.py:513-578 (functions u)
* ``LCEDualWeakForm``           R/src/weak_forms/lce_dual_weak_form.py:11-648 (fields s, m, u, t; functions rhs,
  lambda_s, mu_s, kappa_s, lambda_o, mu_o, kappa_o, l)
* ``LCEDualWeakFormBCDirichlet`` R/src/weak_forms/lce_dual_weak_form.py:651-756 (functions u, t)
* ``LCEScaledDualWeakForm``     R/src/weak_forms/lce_scaled_dual_weak_form.py:12-699 (+ functions grad_l)
* ``LCEScaledDualWeakFormBCDirichlet`` R/src/weak_forms/lce_scaled_dual_weak_form.py:702-811
* ``LERieszMapWeakForm``        R/src/weak_forms/le_riesz_map_weak_form.py:9-360
* ``LCERieszMapWeakForm``       R/src/weak_forms/lce_riesz_map_weak_form.py:9-546
* ``LCEScaledRieszMapWeakForm`` R/src/weak_forms/lce_scaled_riesz_map_weak_form.py:9-598

All of them share one block structure (see ``nfa_local_hdiv_tensor`` in include/nfa.h); a class only says which
per-point coefficient arrays multiply which block.  The Jacobian follows the reference's closed-form
``evaluate_form_vectorized`` bodies (which return the load vector only); ``evaluate_form`` / whole-mesh assembly add
``J alpha`` as the reference's AD version does for these linear forms.
"""
import ctypes

import numpy as np
import torch

from .. import _lib, coefficients
from ..basis.orientation import case_matrices
from ..device import space_device
from .laplace_dual_weak_form import _all_cases, hdiv_facet_setup
from .weak_form import BatchedWeakForm, FormData, take_rows

_vp, _i32 = ctypes.c_void_p, ctypes.c_int32


class TensorFormStruct(ctypes.Structure):
    """Mirror of ``nfa_tensor_form`` (include/nfa.h)."""

    _fields_ = [
        ("dim", _i32), ("n_q", _i32), ("n_hdiv", _i32), ("n_l2", _i32), ("weights", _vp),
        ("h_n_phi", _i32 * 2), ("h_n_comp", _i32 * 2), ("h_nfd", _i32 * 2),
        ("h_val", _vp * 2), ("h_div", _vp * 2), ("h_Tcases", _vp * 2), ("h_cases", _vp * 2),
        ("h_a_id", _vp * 2), ("h_a_tr", _vp * 2), ("h_a_vol", _vp * 2), ("h_a_dd", _vp * 2),
        ("h_div_cd", _vp * 2), ("h_div_g", _vp * 2),
        ("l_n_phi", _i32 * 2), ("l_n_comp", _i32 * 2),
        ("l_val", _vp * 2), ("l_a_mass", _vp * 2), ("l_load", _vp * 2),
        ("div_coupling", (_i32 * 2) * 2), ("asym_coupling", (_i32 * 2) * 2), ("h_dd_general", _i32 * 2),
    ]


def _stress_coefficients(lam, mu, kappa, dim, scale=None):
    """``a_id, a_tr, a_vol`` of the compliance  A s : t = 1/(2 mu) sym s : sym t + 1/(2 kappa) skw s : skw t - vf tr s tr t
    (R/src/weak_forms/lce_dual_weak_form.py:506-529): symm = inner + outer, skew = inner - outer, each halved."""
    cs, ck = 1.0 / (2.0 * mu), 1.0 / (2.0 * kappa)
    vf = cs * (lam / (2.0 * mu + dim * lam))
    s = 1.0 if scale is None else scale
    return 0.5 * (cs + ck) * s, 0.5 * (cs - ck) * s, -vf * s


class HdivTensorMixedForm(BatchedWeakForm):
    form_kind = "form"
    hdiv_names = ("s",)
    l2_names = ("u", "t")

    def n_items_total(self):
        return len(self.space.discrete_spaces[self.hdiv_names[0]].elements)

    def dest_table(self):
        return self.space.destination_table()

    def point_data(self, x, dim, spaces):
        """-> dict: per H(div) field name {a_id, a_tr, a_vol, a_dd, cd, g}, per L2 field name {a_mass, load},
        ``div`` / ``asym``: lists of (l2 name, hdiv name) couplings."""
        raise NotImplementedError

    def _fields(self):
        names = list(self.space.discrete_spaces.keys())
        want = list(self.hdiv_names) + list(self.l2_names)
        if names != want:
            raise _lib.NfaError("%s expects the fields %s in this order" % (type(self).__name__, want))
        sp = self.space.discrete_spaces
        for n in self.hdiv_names:
            if sp[n].family not in ("RT", "BDM"):
                raise _lib.NfaError("%s: field %r must be RT or BDM" % (type(self).__name__, n))
        for n in self.l2_names:
            if sp[n].family != "Lagrange":
                raise _lib.NfaError("%s: field %r must be (discontinuous) Lagrange" % (type(self).__name__, n))
        return sp

    def prepare(self, dev, items=None, arrays=None):
        self._check()
        sp = self._fields()
        first = sp[self.hdiv_names[0]]
        dim = first.dimension
        points, weights = self.space.quadrature[dim]
        nq = len(weights)
        up = lambda a, dt=np.float64: torch.as_tensor(np.ascontiguousarray(a, dtype=dt)).to(dev.device)
        cell_nodes = take_rows(dev.cell_nodes(self.hdiv_names[0]), items)
        n = int(cell_nodes.shape[0])
        x = coefficients.physical_points(dev, dim, points, cell_nodes)
        pd = self.point_data(x, dim, sp)
        st = TensorFormStruct()
        keep = [cell_nodes]
        st.dim, st.n_q, st.n_hdiv, st.n_l2 = dim, nq, len(self.hdiv_names), len(self.l2_names)
        w = up(weights)
        keep.append(w)
        st.weights = w.data_ptr()

        def arr(v, shape):
            if v is None:
                return None
            v = np.array(np.broadcast_to(np.asarray(v, dtype=np.float64), shape))
            t = up(v)
            keep.append(t)
            return t.data_ptr()

        for h, name in enumerate(self.hdiv_names):
            s = sp[name]
            ref = s.ref_element
            nfd = len(ref.entity_dofs[dim - 1][0])
            for f, dofs in enumerate(ref.entity_dofs[dim - 1]):
                if list(dofs) != list(range(f * nfd, (f + 1) * nfd)):
                    raise _lib.NfaError("H(div) element: facet dofs are not facet-major")
            tab = ref.tabulate(1, points)
            val, div = up(tab[0]), up(sum(tab[1 + d][:, :, d] for d in range(dim)))
            tc = up(case_matrices(ref))
            cases = take_rows(torch.as_tensor(_all_cases(s)).to(dev.device), items).contiguous()
            keep += [val, div, tc, cases]
            st.h_n_phi[h], st.h_n_comp[h], st.h_nfd[h] = ref.n_phi, s.n_comp, nfd
            st.h_val[h], st.h_div[h], st.h_Tcases[h], st.h_cases[h] = (
                val.data_ptr(), div.data_ptr(), tc.data_ptr(), cases.data_ptr())
            c = pd.get(name, {})
            st.h_a_id[h], st.h_a_tr[h] = arr(c.get("a_id"), (n, nq)), arr(c.get("a_tr"), (n, nq))
            st.h_a_vol[h], st.h_a_dd[h] = arr(c.get("a_vol"), (n, nq)), arr(c.get("a_dd"), (n, nq))
            st.h_div_cd[h], st.h_div_g[h] = arr(c.get("cd"), (n, nq)), arr(c.get("g"), (n, nq, dim))
            st.h_dd_general[h] = 1 if c.get("dd_general") else 0
        for l, name in enumerate(self.l2_names):
            s = sp[name]
            lv = up(s.ref_element.tabulate(0, points)[0, :, :, 0])
            keep.append(lv)
            st.l_n_phi[l], st.l_n_comp[l], st.l_val[l] = s.ref_element.n_phi, s.n_comp, lv.data_ptr()
            c = pd.get(name, {})
            st.l_a_mass[l] = arr(c.get("a_mass"), (n, nq))
            load = c.get("load")
            if load is not None and not np.any(load):
                load = None
            st.l_load[l] = arr(load, (n, nq, s.n_comp))
        for key, table in (("div", st.div_coupling), ("asym", st.asym_coupling)):
            for (ln, hn) in pd.get(key, []):
                table[self.l2_names.index(ln)][self.hdiv_names.index(hn)] = 1
        return FormData(n_items=n, items=items, dim=dim, n_q=nq, struct=st, keep=keep, cell_nodes=cell_nodes)

    def launch_local(self, dev, data, alpha_dev, dest_dev, Ke, re):
        if alpha_dev is not None and dest_dev is None:
            raise _lib.NfaError("%s: alpha given without the destination table" % type(self).__name__)
        _lib.call(
            "nfa_local_hdiv_tensor", ctypes.byref(data.struct), _lib.ptr(dev.coords), _lib.ptr(data.cell_nodes),
            data.n_items, _lib.ptr(alpha_dev), _lib.ptr(dest_dev), _lib.ptr(Ke), _lib.ptr(re), _lib.stream_ptr(),
        )

    def evaluate_form_vectorized(self, element_index, alpha):
        """The reference's closed-form variant: the Jacobian blocks and the load vector only (``alpha`` is ignored, as
        in R/src/weak_forms/lce_dual_weak_form.py:410-648)."""
        return BatchedWeakForm.evaluate_form(self, element_index, None)

    def _split_rhs(self, x, dim, sp):
        """``rhs`` returns the force components followed by the couple components (R/...lce_dual_weak_form.py:487-501)."""
        nu, nt = sp["u"].n_comp, sp["t"].n_comp
        f = coefficients.evaluate(self.functions["rhs"], x, nu + nt)
        return f[:, :, :nu], f[:, :, nu : nu + nt]


class LEDualWeakForm(HdivTensorMixedForm):
    hdiv_names, l2_names = ("s",), ("u", "t")

    def point_data(self, x, dim, sp):
        lam = coefficients.evaluate(self.functions["lambda"], x)
        mu = coefficients.evaluate(self.functions["mu"], x)
        fu, ft = self._split_rhs(x, dim, sp)
        cs = 1.0 / (2.0 * mu)
        return {
            "s": {"a_id": cs, "a_vol": -cs * (lam / (2.0 * mu + dim * lam))}, "u": {"load": fu}, "t": {"load": ft},
            "div": [("u", "s")], "asym": [("t", "s")],
        }


class LCEDualWeakForm(HdivTensorMixedForm):
    hdiv_names, l2_names = ("s", "m"), ("u", "t")
    scaled = False

    def point_data(self, x, dim, sp):
        ev = lambda k: coefficients.evaluate(self.functions[k], x)
        a_id, a_tr, a_vol = _stress_coefficients(ev("lambda_s"), ev("mu_s"), ev("kappa_s"), dim)
        lc = ev("l")
        m_scale = None if self.scaled else 1.0 / lc
        if dim == 3:
            m_id, m_tr, m_vol = _stress_coefficients(ev("lambda_o"), ev("mu_o"), ev("kappa_o"), dim, m_scale)
            m = {"a_id": m_id, "a_tr": m_tr, "a_vol": m_vol}
        else:  # the couple stress of a 2-D problem is a vector: (1 / (mu_o + kappa_o)) m . v
            gamma = 1.0 / (ev("mu_o") + ev("kappa_o"))
            m = {"a_id": gamma if m_scale is None else gamma * m_scale}
        if self.scaled:  # div (l m) = grad l . m + l div m  (R/src/weak_forms/lce_scaled_dual_weak_form.py:669-687)
            m["cd"] = lc
            m["g"] = coefficients.evaluate(self.functions["grad_l"], x, dim)
        fu, ft = self._split_rhs(x, dim, sp)
        return {
            "s": {"a_id": a_id, "a_tr": a_tr, "a_vol": a_vol}, "m": m, "u": {"load": fu}, "t": {"load": ft},
            "div": [("u", "s"), ("t", "m")], "asym": [("t", "s")],
        }


class LCEScaledDualWeakForm(LCEDualWeakForm):
    scaled = True


class LERieszMapWeakForm(HdivTensorMixedForm):
    """Block-diagonal norm matrix: H(div) inner product of the stress (compliance-weighted) + L2 masses."""

    hdiv_names, l2_names = ("s",), ("u", "t")

    def point_data(self, x, dim, sp):
        lam = coefficients.evaluate(self.functions["lambda"], x)
        mu = coefficients.evaluate(self.functions["mu"], x)
        cs = 1.0 / (2.0 * mu)
        one = np.ones_like(mu)
        return {"s": {"a_id": cs, "a_vol": -cs * (lam / (2.0 * mu + dim * lam)), "a_dd": one},
                "u": {"a_mass": one}, "t": {"a_mass": one}}


class LCERieszMapWeakForm(HdivTensorMixedForm):
    hdiv_names, l2_names = ("s", "m"), ("u", "t")
    scaled = False

    def point_data(self, x, dim, sp):
        ev = lambda k: coefficients.evaluate(self.functions[k], x)
        a_id, a_tr, a_vol = _stress_coefficients(ev("lambda_s"), ev("mu_s"), ev("kappa_s"), dim)
        lc = ev("l")
        one = np.ones_like(lc)
        m_scale = None if self.scaled else 1.0 / lc
        if dim == 3:
            m_id, m_tr, m_vol = _stress_coefficients(ev("lambda_o"), ev("mu_o"), ev("kappa_o"), dim, m_scale)
            m = {"a_id": m_id, "a_tr": m_tr, "a_vol": m_vol}
        else:
            gamma = 1.0 / (ev("mu_o") + ev("kappa_o"))
            m = {"a_id": gamma if m_scale is None else gamma * m_scale}
        m["a_dd"] = one
        if self.scaled:  # norm of div (l m) = grad l . m + l div m  (lce_scaled_riesz_map_weak_form.py:546-566)
            m["cd"], m["g"], m["dd_general"] = lc, coefficients.evaluate(self.functions["grad_l"], x, dim), True
        return {"s": {"a_id": a_id, "a_tr": a_tr, "a_vol": a_vol, "a_dd": one}, "m": m,
                "u": {"a_mass": one}, "t": {"a_mass": one}}


class LCEScaledRieszMapWeakForm(LCERieszMapWeakForm):
    scaled = True


class HdivTensorNaturalBC(BatchedWeakForm):
    """``r = -int_F g_c (psi_a . n)`` on the facet dofs of the first neighbour for every H(div) field listed in
    ``field_data`` (field name -> functions key), no matrix."""

    form_kind = "bc_form"
    field_data = (("s", "u"),)
    sign = -1.0

    def n_items_total(self):
        return len(self.space.discrete_spaces[self.field_data[0][0]].bc_elements)

    def dest_table(self):
        return self.space.bc_destination_table()

    def boundary_values(self, key, x, n_comp, field):
        return coefficients.evaluate(self.functions[key], x, n_comp)

    def prepare(self, dev, items=None, arrays=None):
        if self.space is None or self.functions is None:
            raise ValueError("weak form needs .space and .functions")
        parts, off = [], 0
        for name, key in self.field_data:
            s = self.space.discrete_spaces[name]
            d = hdiv_facet_setup(self.space, name, dev, items)
            g = self.boundary_values(key, d.x, s.n_comp, name)
            d.g = torch.as_tensor(np.ascontiguousarray(g, dtype=np.float64)).to(dev.device)
            d.n_comp, d.off = s.n_comp, off
            off += d.nfd * s.n_comp
            parts.append(d)
        return FormData(n_items=parts[0].n_items, items=items, parts=parts, n_loc=off)

    def launch_local(self, dev, data, alpha_dev, dest_dev, Ke, re):
        if Ke is not None:
            Ke.zero_()  # j_el = 0: the entries enter the pattern as explicit zeros
        for d in data.parts:
            _lib.call(
                "nfa_local_hdiv_facet_vec", d.dim, d.nfd, d.n_q, d.n_comp, _lib.ptr(d.trace_tab), _lib.ptr(d.tidx),
                _lib.ptr(d.w), _lib.ptr(d.tcases), _lib.ptr(d.cases), _lib.ptr(dev.coords), _lib.ptr(d.facet_nodes),
                _lib.ptr(d.neigh_nodes), _lib.ptr(d.normals), _lib.ptr(d.g), d.n_items, float(self.sign), int(d.off),
                int(data.n_loc), _lib.ptr(re), _lib.stream_ptr(),
            )


class LEDualWeakFormBCDirichlet(HdivTensorNaturalBC):
    field_data = (("s", "u"),)


class LCEDualWeakFormBCDirichlet(HdivTensorNaturalBC):
    field_data = (("s", "u"), ("m", "t"))


class LCEScaledDualWeakFormBCDirichlet(HdivTensorNaturalBC):
    """The couple-stress data is scaled by the characteristic length: ``-int_F l t_D (psi . n)``
    (R/src/weak_forms/lce_scaled_dual_weak_form.py:790-808)."""

    field_data = (("s", "u"), ("m", "t"))

    def boundary_values(self, key, x, n_comp, field):
        g = coefficients.evaluate(self.functions[key], x, n_comp)
        if field == "m":
            g = g * (coefficients.evaluate(self.functions["l"], x) + 1.0e-16)[:, :, None]  # :743
        return g


class HdivTensorPenaltyBC(HdivTensorNaturalBC):
    """Penalised normal trace on the boundary facets: per H(div) field and row component ``J = beta int (psi.n)(psi.n)^T``,
    ``r = -beta int g_c (psi.n)``, beta = 1e12 -- the reference's ``LCEDualWeakFormBCNeumann``
    (R/src/weak_forms/lce_dual_weak_form.py:759-871).  Blocks between different fields / components are stored zeros."""

    beta = 1.0e12

    def launch_local(self, dev, data, alpha_dev, dest_dev, Ke, re):
        Ke.zero_()
        for d in data.parts:
            _lib.call(
                "nfa_local_hdiv_facet_penalty", d.dim, d.nfd, d.n_q, d.n_comp, _lib.ptr(d.trace_tab), _lib.ptr(d.tidx),
                _lib.ptr(d.w), _lib.ptr(d.tcases), _lib.ptr(d.cases), _lib.ptr(dev.coords), _lib.ptr(d.facet_nodes),
                _lib.ptr(d.neigh_nodes), _lib.ptr(d.normals), _lib.ptr(d.g), d.n_items, float(self.beta), int(d.off),
                int(data.n_loc), _lib.ptr(Ke), _lib.ptr(re), _lib.stream_ptr(),
            )


class LEDualWeakFormBCNeumann(HdivTensorPenaltyBC):
    """Stress penalty of the three-field form (the class the reference names ``LCEDualWeakFormBCNeumann`` inside
    le_dual_weak_form.py:581-644: field ``s`` only), functions key ``s``."""

    field_data = (("s", "s"),)


class LCEDualWeakFormBCNeumann(HdivTensorPenaltyBC):
    """Stress and couple-stress penalty, functions keys ``s`` and ``m`` (R/src/weak_forms/lce_dual_weak_form.py:759-871;
    the copy in lce_scaled_dual_weak_form.py:814-917 is identical)."""

    field_data = (("s", "s"), ("m", "m"))
