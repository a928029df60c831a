"""Primal forms on several Lagrange fields -- GPU versions of

* ``LCEPrimalWeakForm`` / ``LCEPrimalWeakFormBCDirichlet``   R/src/weak_forms/lce_primal_weak_form.py:9-220, :223-297
* ``LETaylorHoodWeakForm`` / ``LETaylorHoodWeakFormBCDirichlet``  R/src/weak_forms/le_taylor_hood_weak_form.py:9-186, :189-237
* ``LEPrimalStressConstraintWeakForm``  R/src/weak_forms/le_primal_stress_constraint_weak_form.py:10-126

(same class names, ``functions`` keys and sign conventions).  Each form is the reference's pattern
``el_form += det w (grad_phi @ S.T, phi @ s)``: a generalised stress ``Sigma[(field, comp), op]`` (op 0 = tested with
the value, op 1 + d = tested with d/dx_d) that is linear in the generalised strain ``E[(field, comp), op]`` of the
discrete fields, with scalar coefficient functions in front: ``Sigma = sum_m coef_m(x) B_m E``.  A form states its
constitutive map once (``stress``), the constant matrices ``B_m`` are read off by probing it with unit strains, and one
CUDA kernel (``nfa_local_h1_multifield``, csrc/nfa_multi.cu) contracts their non-zeros with the mapped basis.
"""
import ctypes

import numpy as np
import torch

from .. import _lib, coefficients
from .h1_common import H1FacetForm, scalar_table
from .weak_form import BatchedWeakForm, FormData, take_rows


class H1MultiFieldForm(BatchedWeakForm):
    form_kind = "form"
    field_names = ()  # in the order of the product space (= the order of the local dof blocks)
    n_coefficients = 0

    # -- provided by the concrete form ---------------------------------------------------------------------------
    def coefficient_values(self, x):
        """list of ``[n, n_q]`` arrays, one per coefficient of ``stress``."""
        raise NotImplementedError

    def stress(self, dim, E, coef):
        """``Sigma[n_crow, 1 + dim]`` for the strain ``E[n_crow, 1 + dim]`` and scalar coefficients ``coef``."""
        raise NotImplementedError

    def load_values(self, dim, x):
        """list of ``((comp_row, op), values[n, n_q])`` entering the residual as ``+ int values D^op phi``."""
        return []

    # -------------------------------------------------------------------------------------------------------------
    def term_lists(self, dim, n_crow):
        """The non-zeros of ``B_m[(cr, s), (cr', t)]``, read off by probing ``stress`` with unit strains and unit
        coefficients, as the CSR over component-row pairs the kernel walks: ``(ptr[n_crow^2 + 1], code, val)`` with
        ``code = s | t << 4 | m << 8``."""
        nop, M = 1 + dim, self.n_coefficients
        B = np.zeros((M, n_crow, nop, n_crow, nop))
        for m in range(M):
            coef = [1.0 if k == m else 0.0 for k in range(M)]
            for cr in range(n_crow):
                for t in range(nop):
                    E = np.zeros((n_crow, nop))
                    E[cr, t] = 1.0
                    B[m, :, :, cr, t] = np.asarray(self.stress(dim, E, coef), dtype=np.float64)
        ptr, code, val = [0], [], []
        for cr in range(n_crow):
            for cr2 in range(n_crow):
                for m in range(M):
                    for s_ in range(nop):
                        for t in range(nop):
                            v = B[m, cr, s_, cr2, t]
                            if v != 0.0:
                                code.append(s_ | (t << 4) | (m << 8))
                                val.append(v)
                ptr.append(len(code))
        return ptr, code, val

    def _fields(self):
        names = list(self.space.discrete_spaces.keys())
        if names[: len(self.field_names)] != list(self.field_names) or len(names) != len(self.field_names):
            raise _lib.NfaError(
                "%s expects the product space fields %r (in this order), found %r" % (type(self).__name__, self.field_names, names))
        fields = [self.space.discrete_spaces[n] for n in self.field_names]
        for s in fields:
            if s.family != "Lagrange":
                raise _lib.NfaError("%s needs Lagrange fields" % type(self).__name__)
        return fields

    def n_items_total(self):
        return len(self.space.discrete_spaces[self.field_names[0]].elements)

    def dest_table(self):
        return self.space.destination_table()

    def _layout(self, dev):
        """Device tables that depend on the spaces only: concatenated basis table, local dof maps, term lists."""
        fields = self._fields()
        dim = fields[0].dimension
        points, weights = self.space.quadrature[dim]
        key = (type(self).__name__, "mf", dim, len(weights), tuple((s.k_order, s.n_comp) for s in fields))

        def build():
            tabs = [scalar_table(s.ref_element, points) for s in fields]
            tab = np.ascontiguousarray(np.concatenate(tabs, axis=2))
            dof_a, dof_cr, a0, cr0 = [], [], 0, 0
            for s, t in zip(fields, tabs):
                for a in range(t.shape[2]):
                    for c in range(s.n_comp):
                        dof_a.append(a0 + a)
                        dof_cr.append(cr0 + c)
                a0 += t.shape[2]
                cr0 += s.n_comp
            n_crow = cr0
            ptr, code, val = self.term_lists(dim, n_crow)
            i32 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.int32)).to(dev.device)
            return {
                "tab": torch.as_tensor(tab).to(dev.device), "w": torch.as_tensor(np.asarray(weights, dtype=np.float64)).to(dev.device),
                "dof_a": i32(dof_a), "dof_cr": i32(dof_cr), "term_ptr": i32(ptr), "term_code": i32(code if code else [0]),
                "term_val": torch.as_tensor(np.asarray(val if val else [0.0], dtype=np.float64)).to(dev.device),
                "n_phi": int(tab.shape[2]), "n_loc": len(dof_a), "n_crow": n_crow, "n_q": len(weights), "dim": dim,
            }

        return dev._get(key, build), points

    def prepare(self, dev, items=None):
        self._check()
        lay, points = self._layout(dev)
        dim = lay["dim"]
        cell_nodes = take_rows(dev.cell_nodes(self.field_names[0]), items)
        n = int(cell_nodes.shape[0])
        x = coefficients.physical_points(dev, dim, points, cell_nodes)
        coefs = []
        for v in self.coefficient_values(x):
            arr, sc, sq, _ = coefficients.classify(np.asarray(v, dtype=np.float64) * np.ones((n, lay["n_q"])))
            coefs.append((torch.as_tensor(arr).to(dev.device), sc, sq))
        loads = [(k, np.asarray(v, dtype=np.float64)) for k, v in self.load_values(dim, x)]
        loads = [(k, v) for k, v in loads if np.any(v)]
        load_dev = load_code = None
        if loads:
            load_dev = torch.as_tensor(np.ascontiguousarray(np.stack([v for _, v in loads], axis=2))).to(dev.device)
            load_code = torch.as_tensor(np.asarray([cr | (op << 8) for (cr, op), _ in loads], dtype=np.int32)).to(dev.device)
        return FormData(n_items=n, items=items, lay=lay, cell_nodes=cell_nodes, coefs=coefs, load=load_dev,
                        load_code=load_code, n_load=len(loads))

    def launch_local(self, dev, data, alpha_dev, dest_dev, Ke, re):
        lay = data.lay
        M = len(data.coefs)
        ptrs = (ctypes.c_void_p * M)(*[c[0].data_ptr() for c in data.coefs])
        sc = (ctypes.c_int64 * M)(*[int(c[1]) for c in data.coefs])
        sq = (ctypes.c_int64 * M)(*[int(c[2]) for c in data.coefs])
        _lib.call(
            "nfa_local_h1_multifield", lay["dim"], lay["n_q"], lay["n_phi"], lay["n_loc"], lay["n_crow"],
            _lib.ptr(lay["tab"]), _lib.ptr(lay["w"]), _lib.ptr(lay["dof_a"]), _lib.ptr(lay["dof_cr"]),
            _lib.ptr(lay["term_ptr"]), _lib.ptr(lay["term_code"]), _lib.ptr(lay["term_val"]), M, ptrs, sc, sq,
            data.n_load, _lib.ptr(data.load_code), _lib.ptr(data.load), _lib.ptr(dev.coords), _lib.ptr(data.cell_nodes),
            data.n_items, _lib.ptr(alpha_dev), _lib.ptr(dest_dev), _lib.ptr(Ke), _lib.ptr(re), _lib.stream_ptr(),
        )


def _levi_civita():
    eps = np.zeros((3, 3, 3))
    eps[0, 1, 2] = eps[1, 2, 0] = eps[2, 0, 1] = 1.0
    eps[0, 2, 1] = eps[2, 1, 0] = eps[1, 0, 2] = -1.0
    return eps


class LCEPrimalWeakForm(H1MultiFieldForm):
    """Primal Cosserat elasticity, displacement ``u`` and rotation ``t`` (3 components in 3-D, 1 in 2-D), functions
    rhs / lambda / mu / kappa / gamma:  e = grad u + Theta(t),  s = 2 mu sym e + 2 kappa skw e + lambda tr e I,
    k = gamma grad t;  r = +int f . (phi_u, phi_t) + int s : grad phi_u + int k : grad phi_t + int axl(skw s) . phi_t
    (R/src/weak_forms/lce_primal_weak_form.py:52-216)."""

    field_names = ("u", "t")
    n_coefficients = 4  # lambda, mu, kappa, gamma

    def coefficient_values(self, x):
        return [coefficients.evaluate(self.functions[k], x) for k in ("lambda", "mu", "kappa", "gamma")]

    def stress(self, dim, E, coef):
        lam, mu, kap, gam = coef
        nt = 3 if dim == 3 else 1
        grad_u = E[:dim, 1:]  # [c, d] = d u_c / d x_d
        th, grad_t = E[dim : dim + nt, 0], E[dim : dim + nt, 1:]
        if dim == 3:
            theta = -np.einsum("cde,e->cd", _levi_civita(), th)  # [[0, -t2, t1], [t2, 0, -t0], [-t1, t0, 0]]
        else:
            theta = th[0] * np.array([[0.0, -1.0], [1.0, 0.0]])
        eh = grad_u + theta
        sh = 2.0 * mu * 0.5 * (eh + eh.T) + 2.0 * kap * 0.5 * (eh - eh.T) + lam * np.trace(eh) * np.eye(dim)
        skew = 0.5 * (sh - sh.T)
        if dim == 3:
            s_cross = np.array([skew[2, 1] - skew[1, 2], skew[0, 2] - skew[2, 0], skew[1, 0] - skew[0, 1]])
        else:
            s_cross = np.array([skew[1, 0] - skew[0, 1]])
        S = np.zeros_like(E)
        S[:dim, 1:] = sh            # (grad_phi_u @ sh.T)[a, c] = sum_d d_d phi_a sh[c, d]
        S[dim:, 1:] = gam * grad_t  # grad_phi_t @ k.T
        S[dim:, 0] = s_cross        # phi_t @ S_cross
        return S

    def load_values(self, dim, x):
        nt = 3 if dim == 3 else 1
        f = coefficients.evaluate(self.functions["rhs"], x, dim + nt)
        return [((c, 0), f[:, :, c]) for c in range(dim + nt)]


class LETaylorHoodWeakForm(H1MultiFieldForm):
    """Displacement-pressure (P2-P1) elasticity as the reference codes it in 2-D
    (R/src/weak_forms/le_taylor_hood_weak_form.py:72-129; its 3-D branch starts with ``assert False``):
    s_dev = s - tr(s)/dim I with s = 2 mu e + lambda tr e I;  r_u = +int f . phi + int (s_dev + p I) : grad phi,
    r_p = int psi div u - int psi p / (lambda + mu)."""

    field_names = ("u", "p")
    n_coefficients = 4  # lambda, mu, 1, 1 / (lambda + mu)

    def coefficient_values(self, x):
        lam = coefficients.evaluate(self.functions["lambda"], x)
        mu = coefficients.evaluate(self.functions["mu"], x)
        return [lam, mu, np.ones_like(lam), 1.0 / (lam + mu)]

    def stress(self, dim, E, coef):
        if dim != 2:
            raise NotImplementedError("LETaylorHoodWeakForm: the reference implements dimension 2 only (its 3-D branch asserts False)")
        lam, mu, one, inv_lm = coef
        grad_u, ph = E[:dim, 1:], E[dim, 0]
        eh = 0.5 * (grad_u + grad_u.T)
        sh = 2.0 * mu * eh + lam * np.trace(eh) * np.eye(dim)
        sh_dev = sh - (1.0 / dim) * np.trace(sh) * np.eye(dim)
        S = np.zeros_like(E)
        S[:dim, 1:] = sh_dev + one * ph * np.eye(dim)
        S[dim, 0] = one * np.trace(grad_u) - inv_lm * ph
        return S

    def load_values(self, dim, x):
        f = coefficients.evaluate(self.functions["rhs"], x, dim)
        return [((c, 0), f[:, :, c]) for c in range(dim)]


class LEPrimalStressConstraintWeakForm(H1MultiFieldForm):
    """Stress-target penalty on a displacement field: r = beta int (s(u) - s_target) : grad phi, beta = 1e6
    (R/src/weak_forms/le_primal_stress_constraint_weak_form.py:26, :82-122); functions s_target / lambda / mu."""

    field_names = ("u",)
    n_coefficients = 2
    beta = 1.0e6

    def coefficient_values(self, x):
        return [self.beta * coefficients.evaluate(self.functions[k], x) for k in ("lambda", "mu")]

    def stress(self, dim, E, coef):
        lam, mu = coef
        grad_u = E[:dim, 1:]
        eh = 0.5 * (grad_u + grad_u.T)
        S = np.zeros_like(E)
        S[:dim, 1:] = 2.0 * mu * eh + lam * np.trace(eh) * np.eye(dim)
        return S

    def load_values(self, dim, x):
        n, nq, _ = x.shape
        flat = x.reshape(-1, 3)
        st = np.asarray(self.functions["s_target"](flat[:, 0], flat[:, 1], flat[:, 2]), dtype=np.float64)
        st = st * np.ones((dim, dim, n * nq)) if st.ndim == 3 else st[:, :, None] * np.ones((dim, dim, n * nq))
        return [((c, 1 + d), -self.beta * st[c, d].reshape(n, nq)) for c in range(dim) for d in range(dim)]


class LETaylorHoodWeakFormBCDirichlet(H1FacetForm):
    """Displacement penalty on the boundary facets (only ``u`` carries boundary elements),
    R/src/weak_forms/le_taylor_hood_weak_form.py:189-237."""

    field_name = "u"
    data_name = "u"
    beta = 1.0e12
    r_sign = -1.0
    with_matrix = True


class LCEPrimalWeakFormBCDirichlet(BatchedWeakForm):
    """Penalty on both fields of the primal Cosserat space, R/src/weak_forms/lce_primal_weak_form.py:223-297: local
    dofs [u | t] of the facet, J = beta int phi phi^T per field and component (no coupling blocks: stored zeros),
    r = -beta int (u_D, t_D) phi.  The per-field blocks come from the facet kernel of the single-field forms."""

    form_kind = "bc_form"
    field_names = ("u", "t")
    beta = 1.0e12

    def _parts(self):
        parts = []
        for name in self.field_names:
            part = type("_LCEPart_" + name, (H1FacetForm,), {
                "field_name": name, "data_name": name, "beta": self.beta, "r_sign": -1.0, "with_matrix": True,
                "_multi": True})(self.space)
            part.functions = self.functions
            parts.append(part)
        return parts

    def n_items_total(self):
        return len(self.space.discrete_spaces[self.field_names[0]].bc_elements)

    def dest_table(self):
        return self.space.bc_destination_table()

    def prepare(self, dev, items=None):
        self._check()
        for name in self.field_names:
            if len(self.space.discrete_spaces[name].bc_elements) != self.n_items_total():
                raise _lib.NfaError("LCEPrimalWeakFormBCDirichlet: u and t must carry the same boundary elements")
        datas = [p.prepare(dev, items) for p in self._parts()]
        return FormData(n_items=datas[0].n_items, items=items, parts=datas)

    def launch_local(self, dev, data, alpha_dev, dest_dev, Ke, re):
        Ke.zero_()
        off = 0
        for part, d in zip(self._parts(), data.parts):
            nl = d.n_phi * d.n_comp
            Kp = torch.empty((d.n_items, nl, nl), dtype=torch.float64, device=dev.device)
            rp = torch.empty((d.n_items, nl), dtype=torch.float64, device=dev.device)
            part.launch_local(dev, d, None, None, Kp, rp)
            Ke[:, off : off + nl, off : off + nl] = Kp
            re[:, off : off + nl] = rp
            off += nl
