"""Shared set-up of the parity tests: the same problem described once for the GPU path and once for the oracle."""
import numpy as np

from nfempy_b200.mesh import box_tet_mesh, square_tri_mesh
from nfempy_b200.spaces import ProductSpace
from oracle import ref_assembly, ref_forms
from oracle.ref_mesh import RefMesh
from oracle.ref_spaces import RefProductSpace


def compare_csr(got, ref, rtol=1e-12, label=""):
    """Pattern bit-exact; values within ``rtol`` of max(|ref|, largest |ref| of the row); rhs likewise vs its max."""
    indptr_g, indices_g, data_g, rhs_g = got
    indptr_r, indices_r, data_r, rhs_r = ref
    assert indptr_g.dtype == np.int32 and indices_g.dtype == np.int32, "CSR index arrays must be int32"
    assert np.array_equal(indptr_g, indptr_r), label + " indptr differs"
    assert np.array_equal(indices_g, indices_r), label + " indices differ"
    rows = np.repeat(np.arange(len(indptr_r) - 1), np.diff(indptr_r))
    row_max = np.zeros(len(indptr_r) - 1)
    np.maximum.at(row_max, rows, np.abs(data_r))
    scale = np.maximum(np.abs(data_r), row_max[rows])
    err = np.abs(data_g - data_r) / np.where(scale > 0, scale, 1.0)
    assert err.max(initial=0.0) <= rtol, "%s matrix values: max scaled error %.3e" % (label, err.max())
    rscale = max(np.abs(rhs_r).max(initial=0.0), 1e-300)
    rerr = np.abs(rhs_g - rhs_r).max(initial=0.0) / rscale
    assert rerr <= rtol, "%s rhs: max scaled error %.3e" % (label, rerr)
    return err.max(initial=0.0), rerr


def _f_rhs(x, y, z):
    return np.array([np.sin(np.pi * x) * y, np.cos(np.pi * y) * z + 1.0, x * y * z - 0.5])


def _u_exact(x, y, z):
    return np.array([(1 - x) * x * y, np.sin(x) + z, 0.25 * z * z + y])


class Case:
    def __init__(self, mesh, data_fn, tags, btags, gpu_forms, ref_forms_fn, disc=None, alpha_fn=None):
        self.mesh, self.data_fn, self.tags, self.btags = mesh, data_fn, tags, btags
        self.gpu_forms, self.ref_forms_fn, self.disc, self.alpha_fn = gpu_forms, ref_forms_fn, disc or {}, alpha_fn

    def space(self):
        s = ProductSpace(self.data_fn(self.mesh))
        s.make_subspaces_discontinuous(self.disc)
        s.build_structures(self.tags, self.btags)
        return s

    def ref_space(self):
        r = RefMesh(self.mesh.dimension, self.mesh.points, self.mesh._blocks)
        r.build_conformal_mesh()
        s = RefProductSpace(self.data_fn(r))
        s.make_subspaces_discontinuous(self.disc)
        s.build_structures(self.tags, self.btags)
        return s

    def run_gpu(self, allow_fused=True, **policy):
        from nfempy_b200.assembly import GPUAssembler

        space = self.space()
        alpha = np.zeros(space.n_dof) if self.alpha_fn is None else self.alpha_fn(space.n_dof)
        asm = GPUAssembler(space, None, np.zeros(space.n_dof), allow_fused=allow_fused, **policy)
        asm.form_to_input_list = self.gpu_forms(space, alpha)
        return asm.assemble()

    def run_oracle(self, upper=None, explicit_diagonal=False):
        space = self.ref_space()
        alpha = np.zeros(space.n_dof) if self.alpha_fn is None else self.alpha_fn(space.n_dof)
        if upper is not None:
            return ref_assembly.assemble_filtered(space, self.ref_forms_fn(space), alpha, upper, explicit_diagonal)
        return ref_assembly.assemble(space, self.ref_forms_fn(space), alpha)


def elasticity_case(n=2, with_bc=True, varying=False, alpha=False, k=2, dim=3, mesh=None, neumann=False,
                    normal_bc=False):
    from nfempy_b200.weak_forms import (LEPrimalWeakForm, LEPrimalWeakFormBCDirichlet, LEPrimalWeakFormBCNeumann,
                                        LEPrimalWeakFormBCNormalDirichlet)

    if mesh is None:
        mesh = box_tet_mesh(n, jitter=0.2, shuffle_local_nodes=True) if dim == 3 else square_tri_mesh(n, jitter=0.2)
    btag_list = [2, 3, 4, 5, 6, 7] if dim == 3 else [2, 3, 4, 5]
    if varying:
        f_lam = lambda x, y, z: 1.0 + x + 2.0 * y * y
        f_mu = lambda x, y, z: 2.0 + np.sin(z + x)
    else:
        f_lam = lambda x, y, z: 3.0 * np.ones_like(x)
        f_mu = lambda x, y, z: 0.5 * np.ones_like(x)
    rhs = _f_rhs if dim == 3 else (lambda x, y, z: _f_rhs(x, y, z)[:2])
    uex = _u_exact if dim == 3 else (lambda x, y, z: _u_exact(x, y, z)[:2])
    functions = {"rhs": rhs, "lambda": f_lam, "mu": f_mu}
    bc_functions = {"u": uex}
    t_functions = {"t": rhs}
    dir_tags = btag_list

    def gpu_forms(space, a):
        wf = LEPrimalWeakForm(space)
        wf.functions = functions
        out = {"volume": ["form", range(len(space.discrete_spaces["u"].elements)), wf, a]}
        if with_bc:
            n_bc = len(space.discrete_spaces["u"].bc_elements)
            if neumann or normal_bc:
                # split the boundary elements between a Dirichlet part and the other condition
                bf = LEPrimalWeakFormBCDirichlet(space)
                bf.functions = bc_functions
                out["bc"] = ["bc_form", list(range(0, n_bc, 2)), bf, a]
                of = LEPrimalWeakFormBCNeumann(space) if neumann else LEPrimalWeakFormBCNormalDirichlet(space)
                of.functions = t_functions if neumann else bc_functions
                out["bc2"] = ["bc_form", list(range(1, n_bc, 2)), of, a]
            else:
                bf = LEPrimalWeakFormBCDirichlet(space)
                bf.functions = bc_functions
                out["bc"] = ["bc_form", range(n_bc), bf, a]
        return out

    def ref_forms_fn(space):
        out = [("form", ref_forms.le_primal, functions, len(space.discrete_spaces["u"].element_ids))]
        if with_bc:
            n_bc = len(space.discrete_spaces["u"].bc_element_ids)
            if neumann or normal_bc:
                out.append(("bc_form", ref_forms.le_primal_bc_dirichlet, bc_functions, list(range(0, n_bc, 2))))
                other = ref_forms.le_primal_bc_neumann if neumann else ref_forms.le_primal_bc_normal_dirichlet
                out.append(("bc_form", other, t_functions if neumann else bc_functions, list(range(1, n_bc, 2))))
            else:
                out.append(("bc_form", ref_forms.le_primal_bc_dirichlet, bc_functions, n_bc))
        return out

    alpha_fn = (lambda nd: np.cos(np.arange(nd) * 0.37) + 0.1) if alpha else None
    return Case(
        mesh, lambda g: {"u": (dim, dim, "Lagrange", k, g)}, {"u": [1]}, {"u": dir_tags if with_bc else []},
        gpu_forms, ref_forms_fn, alpha_fn=alpha_fn,
    )


def laplace_case(n=4, dim=2, k=1, with_bc=True, varying=False, alpha=False, mesh=None):
    from nfempy_b200.weak_forms import LaplacePrimalWeakForm, LaplacePrimalWeakFormBCDirichlet

    if mesh is None:
        mesh = square_tri_mesh(n, jitter=0.2) if dim == 2 else box_tet_mesh(n, jitter=0.2, shuffle_local_nodes=True)
    btag_list = [2, 3, 4, 5] if dim == 2 else [2, 3, 4, 5, 6, 7]
    kappa = (lambda x, y, z: 1.0 + x * x + y) if varying else (lambda x, y, z: 2.0)
    functions = {
        "rhs": lambda x, y, z: np.array([2.0 * np.pi**2 * np.sin(np.pi * x) * np.sin(np.pi * y) + z]),
        "kappa": kappa,
    }
    bc_functions = {"p": lambda x, y, z: np.array([np.sin(np.pi * x) + y])}

    def gpu_forms(space, a):
        wf = LaplacePrimalWeakForm(space)
        wf.functions = functions
        out = {"volume": ["form", range(len(space.discrete_spaces["p"].elements)), wf, a]}
        if with_bc:
            bf = LaplacePrimalWeakFormBCDirichlet(space)
            bf.functions = bc_functions
            out["bc"] = ["bc_form", range(len(space.discrete_spaces["p"].bc_elements)), bf, a]
        return out

    def ref_forms_fn(space):
        out = [("form", ref_forms.laplace_primal, functions, len(space.discrete_spaces["p"].element_ids))]
        if with_bc:
            out.append(("bc_form", ref_forms.laplace_primal_bc_dirichlet, bc_functions, len(space.discrete_spaces["p"].bc_element_ids)))
        return out

    alpha_fn = (lambda nd: np.sin(np.arange(nd) * 0.61) + 0.3) if alpha else None
    return Case(
        mesh, lambda g: {"p": (dim, 1, "Lagrange", k, g)}, {"p": [1]}, {"p": btag_list if with_bc else []},
        gpu_forms, ref_forms_fn, alpha_fn=alpha_fn,
    )


def mixed_poisson_case(dim=3, k=1, n=2, family="RT", with_bc=True, varying=False, mesh=None):
    from nfempy_b200.weak_forms import LaplaceDualWeakForm, LaplaceDualWeakFormBCDirichlet

    if mesh is None:
        mesh = box_tet_mesh(n, jitter=0.2, shuffle_local_nodes=True) if dim == 3 else square_tri_mesh(n, jitter=0.2)
    btag_list = [2, 3, 4, 5, 6, 7] if dim == 3 else [2, 3, 4, 5]
    kappa = (lambda x, y, z: 1.0 + x * x + 0.5 * y) if varying else (lambda x, y, z: 2.0)
    functions = {"rhs": lambda x, y, z: np.array([np.sin(np.pi * x) * (1.0 + y) + z]), "kappa": kappa}
    bc_functions = {"p": lambda x, y, z: np.array([x + 2.0 * y - z + np.sin(x)])}
    pk = k - 1 if family == "RT" else k - 1

    def gpu_forms(space, a):
        wf = LaplaceDualWeakForm(space)
        wf.functions = functions
        out = {"volume": ["form", range(len(space.discrete_spaces["u"].elements)), wf, a]}
        if with_bc:
            bf = LaplaceDualWeakFormBCDirichlet(space)
            bf.functions = bc_functions
            out["bc"] = ["bc_form", range(len(space.discrete_spaces["u"].bc_elements)), bf, a]
        return out

    def ref_forms_fn(space):
        out = [("form", ref_forms.laplace_dual, functions, len(space.discrete_spaces["u"].element_ids))]
        if with_bc:
            out.append(("bc_form", ref_forms.laplace_dual_bc_dirichlet, bc_functions, len(space.discrete_spaces["u"].bc_element_ids)))
        return out

    return Case(
        mesh, lambda g: {"u": (dim, 1, family, k, g), "p": (dim, 1, "Lagrange", pk, g)}, {"u": [1], "p": [1]},
        {"u": btag_list if with_bc else []}, gpu_forms, ref_forms_fn, disc={"p": True},
    )


class SundusCase:
    """Config 5 style: mp, mc in RT_k, p, c in P_{k-1}^disc, implicit Euler step with a nonlinear reaction."""

    def __init__(self, dim=3, k=1, n=2, mesh=None, t=0.3, delta_t=0.1):
        self.dim, self.k, self.t = dim, k, t
        self.mesh = mesh if mesh is not None else (
            box_tet_mesh(n, jitter=0.2, shuffle_local_nodes=True) if dim == 3 else square_tri_mesh(n, jitter=0.2))
        self.btags = [2, 3, 4, 5, 6, 7] if dim == 3 else [2, 3, 4, 5]
        self.functions = {
            "rhs_f": lambda x, y, z, t: np.array([np.sin(x + t) + y]), "rhs_r": lambda x, y, z, t: np.array([x * y + t * z]),
            "kappa": lambda x, y, z: 2.0 + x, "delta": lambda x, y, z: 1.5, "eta": lambda x, y, z: 0.7 + 0.1 * y,
            "delta_t": delta_t,
        }
        self.bc_functions = {"p": lambda x, y, z, t: np.array([x + t]), "c": lambda x, y, z, t: np.array([y * z - t])}

    def _data(self, g):
        d, k = self.dim, self.k
        return {"mp": (d, 1, "RT", k, g), "mc": (d, 1, "RT", k, g), "p": (d, 1, "Lagrange", k - 1, g),
                "c": (d, 1, "Lagrange", k - 1, g)}

    def _build(self, space):
        space.make_subspaces_discontinuous({"p": True, "c": True})
        space.build_structures({n: [1] for n in ("mp", "mc", "p", "c")}, {"mp": self.btags, "mc": self.btags})
        return space

    def space(self):
        return self._build(ProductSpace(self._data(self.mesh)))

    def ref_space(self):
        r = RefMesh(self.mesh.dimension, self.mesh.points, self.mesh._blocks)
        r.build_conformal_mesh()
        return self._build(RefProductSpace(self._data(r)))

    @staticmethod
    def alphas(n_dof):
        return 0.3 * np.cos(np.arange(n_dof) * 0.37) + 0.05, 0.2 * np.sin(np.arange(n_dof) * 0.11)

    def gpu_forms(self, space, alphas):
        from nfempy_b200.weak_forms import SundusDualWeakForm, SundusDualWeakFormBCDirichlet

        wf, bf = SundusDualWeakForm(space), SundusDualWeakFormBCDirichlet(space)
        wf.functions, bf.functions = self.functions, self.bc_functions
        mp = space.discrete_spaces["mp"]
        return {
            "volume": ["time_dependent_form", range(len(mp.elements)), wf, alphas, self.t],
            "bc": ["time_dependent_bc_form", range(len(mp.bc_elements)), bf, alphas, self.t],
        }

    def run_gpu(self):
        from nfempy_b200.assembly import GPUAssembler

        space = self.space()
        asm = GPUAssembler(space, None, np.zeros(space.n_dof))
        asm.form_to_input_list = self.gpu_forms(space, self.alphas(space.n_dof))
        return asm.assemble()

    def run_oracle(self):
        space = self.ref_space()
        mp = space.discrete_spaces["mp"]
        forms = [
            ("time_dependent_form", ref_forms.sundus_dual, self.functions, len(mp.element_ids)),
            ("time_dependent_bc_form", ref_forms.sundus_dual_bc_dirichlet, self.bc_functions, len(mp.bc_element_ids)),
        ]
        return ref_assembly.assemble_time_dependent(space, forms, self.alphas(space.n_dof), self.t)


def degenerate_case(kind="degenerate", dim=3, k=1, n=2, with_bc=True, alpha=True, mesh=None):
    """Arbogast-Taicher family (config 3's other form classes): porosity vanishes on part of the domain so that the
    per-point ``isclose(phi, 0)`` branches of the reference are exercised.  kind: degenerate | unscaled | projection."""
    from nfempy_b200 import weak_forms as wf_mod

    if mesh is None:
        mesh = box_tet_mesh(n, jitter=0.2, shuffle_local_nodes=True) if dim == 3 else square_tri_mesh(n, jitter=0.2)
    btag_list = [2, 3, 4, 5, 6, 7] if dim == 3 else [2, 3, 4, 5]
    porosity = lambda x, y, z: np.maximum(x - 0.4, 0.0) * (1.0 + 0.5 * y)
    d_phi = lambda x, y, z: np.maximum(x - 0.4, 0.0) ** 1.5 + 0.0 * y
    grad_d_phi = lambda x, y, z: np.array([1.5 * np.maximum(x - 0.4, 0.0) ** 0.5, 0.1 * y, 0.0 * z])
    functions = {
        "rhs": lambda x, y, z: np.array([np.sin(np.pi * x) * (1.0 + y) + z]), "porosity": porosity, "d_phi": d_phi,
        "grad_d_phi": grad_d_phi,
    }
    flux, scalar = ("u", "p") if kind == "unscaled" else ("v", "q")
    bc_functions = {scalar: lambda x, y, z: np.array([x + 2.0 * y - z + np.sin(x)]), "porosity": porosity, "d_phi": d_phi}
    scaled_fn = lambda nd: 0.4 * np.sin(np.arange(nd) * 0.23) - 0.1
    vol_cls, bc_cls, ref_vol, ref_bc = {
        "degenerate": (wf_mod.DegenerateEllipticWeakForm, wf_mod.DegenerateEllipticWeakFormBCDirichlet,
                       ref_forms.degenerate_elliptic, ref_forms.degenerate_elliptic_bc_dirichlet),
        "unscaled": (wf_mod.UnscaledEllipticWeakForm, wf_mod.UnscaledEllipticWeakFormBCDirichlet,
                     ref_forms.unscaled_elliptic, ref_forms.unscaled_elliptic_bc_dirichlet),
        "projection": (wf_mod.ScaledToPhysicalL2Projection, None, None, None),
    }[kind]

    def gpu_forms(space, a):
        wf = vol_cls(space)
        wf.functions = functions
        if kind == "projection":
            wf.alpha_scaled = scaled_fn(space.n_dof)
        out = {"volume": ["form", range(len(space.discrete_spaces[flux].elements)), wf, a]}
        if with_bc and bc_cls is not None:
            bf = bc_cls(space)
            bf.functions = bc_functions
            out["bc"] = ["bc_form", range(len(space.discrete_spaces[flux].bc_elements)), bf, a]
        return out

    def ref_forms_fn(space):
        n_el = len(space.discrete_spaces[flux].element_ids)
        if kind == "projection":
            scaled = scaled_fn(space.n_dof)

            def proj(sp, fns, iel, alpha_l):
                return ref_forms.scaled_to_physical_l2_projection(sp, fns, iel, alpha_l, scaled[sp.destination_indexes(iel)])

            return [("form", proj, functions, n_el)]
        out = [("form", ref_vol, functions, n_el)]
        if with_bc:
            out.append(("bc_form", ref_bc, bc_functions, len(space.discrete_spaces[flux].bc_element_ids)))
        return out

    alpha_fn = (lambda nd: np.cos(np.arange(nd) * 0.37) + 0.1) if alpha else None
    case = Case(
        mesh, lambda g: {flux: (dim, 1, "RT", k, g), scalar: (dim, 1, "Lagrange", k - 1, g)}, {flux: [1], scalar: [1]},
        {flux: btag_list if with_bc else []}, gpu_forms, ref_forms_fn, disc={scalar: True}, alpha_fn=alpha_fn,
    )
    case.scaled_fn = scaled_fn  # the second state vector of the "projection" kind (tools/make_reference_pins.py)
    return case


def tensor_mixed_case(kind="lce", dim=3, s_family="BDM", m_family="RT", k=1, n=2, with_bc=True, alpha=True, mesh=None,
                      neumann=False):
    """Config 4 style: weakly symmetric mixed elasticity (kind="le": fields s, u, t) and mixed Cosserat elasticity
    (kind="lce" / "lce_scaled": fields s, m, u, t) on row-wise H(div) stresses with discontinuous P_{k-1} multipliers."""
    from nfempy_b200 import weak_forms as wf_mod
    from oracle import ref_tensor_forms as rt

    if mesh is None:
        mesh = box_tet_mesh(n, jitter=0.2, shuffle_local_nodes=True) if dim == 3 else square_tri_mesh(n, jitter=0.2)
    btag_list = [2, 3, 4, 5, 6, 7] if dim == 3 else [2, 3, 4, 5]
    tc = 3 if dim == 3 else 1
    mc = 3 if dim == 3 else 1
    one = lambda x: np.ones_like(x)
    functions = {
        "rhs": lambda x, y, z: np.array([np.sin(x) + y, x * y + 1.0, z - x, np.cos(y), x + 0.5, y * z][: dim + tc] if dim == 3
                                        else [np.sin(x) + y, x * y + 1.0, np.cos(y) + x]),
        "lambda": lambda x, y, z: 1.0 + 0.5 * x, "mu": lambda x, y, z: 2.0 + 0.25 * y * y,
        "lambda_s": lambda x, y, z: 1.0 + 0.5 * x, "mu_s": lambda x, y, z: 2.0 + 0.25 * y * y,
        "kappa_s": lambda x, y, z: 0.3 + 0.1 * x * y, "lambda_o": lambda x, y, z: 0.7 * one(x),
        "mu_o": lambda x, y, z: 1.3 + 0.2 * x, "kappa_o": lambda x, y, z: 0.4 * one(x),
        "l": lambda x, y, z: 0.5 + 0.25 * x + 0.1 * y * y,
        "grad_l": lambda x, y, z: np.array([0.25 * one(x), 0.2 * y, 0.0 * x]),
    }
    bc_functions = {
        "u": lambda x, y, z: np.array([x + y, np.sin(x) - z, 0.5 * y * z][:dim]),
        "t": lambda x, y, z: np.array([y - x, 0.3 * z + 1.0, np.cos(x)][:tc]),
        "l": functions["l"], "grad_l": functions["grad_l"],
    }
    # prescribed normal stresses of the penalised ("Neumann") condition, one value per row of the tensor
    neumann_functions = {
        "s": lambda x, y, z: np.array([1.0 + x * y, np.cos(z) - y, 0.25 * x + z][:dim]),
        "m": lambda x, y, z: np.array([0.5 - y, x * z + 0.1, np.sin(y)][:mc]),
    }
    if kind == "le":
        names, bfields = ("s", "u", "t"), ("s",)
        vol_cls, bc_cls = wf_mod.LEDualWeakForm, wf_mod.LEDualWeakFormBCDirichlet
        ref_vol, ref_bc = rt.le_dual, rt.le_dual_bc_dirichlet
    elif kind == "lce":
        names, bfields = ("s", "m", "u", "t"), ("s", "m")
        vol_cls, bc_cls = wf_mod.LCEDualWeakForm, wf_mod.LCEDualWeakFormBCDirichlet
        ref_vol, ref_bc = rt.lce_dual, rt.lce_dual_bc_dirichlet
    elif kind == "lce_scaled":
        names, bfields = ("s", "m", "u", "t"), ("s", "m")
        vol_cls, bc_cls = wf_mod.LCEScaledDualWeakForm, wf_mod.LCEScaledDualWeakFormBCDirichlet
        ref_vol, ref_bc = rt.lce_scaled_dual, rt.lce_scaled_dual_bc_dirichlet
    else:  # Riesz maps (preconditioner matrices): no boundary form
        with_bc = False
        names = ("s", "u", "t") if kind == "le_riesz" else ("s", "m", "u", "t")
        bfields = ("s",) if kind == "le_riesz" else ("s", "m")
        vol_cls = {"le_riesz": wf_mod.LERieszMapWeakForm, "lce_riesz": wf_mod.LCERieszMapWeakForm,
                   "lce_scaled_riesz": wf_mod.LCEScaledRieszMapWeakForm}[kind]
        ref_vol = {"le_riesz": rt.le_riesz_map, "lce_riesz": rt.lce_riesz_map, "lce_scaled_riesz": rt.lce_scaled_riesz_map}[kind]
        bc_cls = ref_bc = None

    def data_fn(g):
        d = {"s": (dim, dim, s_family, k, g)}
        if "m" in names:
            d["m"] = (dim, mc, m_family, k, g)
        d["u"] = (dim, dim, "Lagrange", k - 1, g)
        d["t"] = (dim, tc, "Lagrange", k - 1, g)
        return d

    def gpu_forms(space, a):
        wf = vol_cls(space)
        wf.functions = functions
        out = {"volume": ["form", range(len(space.discrete_spaces["s"].elements)), wf, a]}
        if with_bc:
            bf = bc_cls(space)
            bf.functions = bc_functions
            n_bc = len(space.discrete_spaces["s"].bc_elements)
            if neumann:  # natural condition on the even boundary facets, penalised normal stresses on the odd ones
                out["bc"] = ["bc_form", list(range(0, n_bc, 2)), bf, a]
                nf = (wf_mod.LEDualWeakFormBCNeumann if kind == "le" else wf_mod.LCEDualWeakFormBCNeumann)(space)
                nf.functions = neumann_functions
                out["bc_n"] = ["bc_form", list(range(1, n_bc, 2)), nf, a]
            else:
                out["bc"] = ["bc_form", range(n_bc), bf, a]
        return out

    def ref_forms_fn(space):
        out = [("form", ref_vol, functions, len(space.discrete_spaces["s"].element_ids))]
        if with_bc:
            n_bc = len(space.discrete_spaces["s"].bc_element_ids)
            if neumann:
                out.append(("bc_form", ref_bc, bc_functions, list(range(0, n_bc, 2))))
                ref_n = rt.le_dual_bc_neumann if kind == "le" else rt.lce_dual_bc_neumann
                out.append(("bc_form", ref_n, neumann_functions, list(range(1, n_bc, 2))))
            else:
                out.append(("bc_form", ref_bc, bc_functions, n_bc))
        return out

    alpha_fn = (lambda nd: np.cos(np.arange(nd) * 0.37) + 0.1) if alpha else None
    return Case(
        mesh, data_fn, {nm: [1] for nm in names}, {nm: (btag_list if with_bc else []) for nm in bfields},
        gpu_forms, ref_forms_fn, disc={"u": True, "t": True}, alpha_fn=alpha_fn,
    )


# ---- primal forms on several Lagrange fields (W8): LCE primal, Taylor-Hood, stress constraint ------------------
def _t_exact(x, y, z):
    return np.array([np.cos(y) * x, 0.5 * z - x * y, np.sin(z) + 0.1])


def lce_primal_case(dim=3, k=2, n=2, varying=False, alpha=False, with_bc=True, kt=None, mesh=None):
    from nfempy_b200.weak_forms import LCEPrimalWeakForm, LCEPrimalWeakFormBCDirichlet
    from oracle import ref_h1_multi

    if mesh is None:
        mesh = box_tet_mesh(n, jitter=0.2, shuffle_local_nodes=True) if dim == 3 else square_tri_mesh(n, jitter=0.2)
    btags = [2, 3, 4, 5, 6, 7] if dim == 3 else [2, 3, 4, 5]
    nt = 3 if dim == 3 else 1
    kt = k if kt is None else kt
    if varying:
        mats = {"lambda": lambda x, y, z: 1.0 + x + 2.0 * y * y, "mu": lambda x, y, z: 2.0 + np.sin(z + x),
                "kappa": lambda x, y, z: 0.3 + 0.1 * x * y, "gamma": lambda x, y, z: 1.5 + z * z}
    else:
        mats = {"lambda": lambda x, y, z: 3.0 + 0.0 * x, "mu": lambda x, y, z: 0.5 + 0.0 * x,
                "kappa": lambda x, y, z: 0.1 + 0.0 * x, "gamma": lambda x, y, z: 2.0 + 0.0 * x}

    def rhs(x, y, z):
        f, t = _f_rhs(x, y, z), _t_exact(x, y, z)
        return np.array(list(f[:dim]) + list(t[:nt]))

    functions = dict(mats, rhs=rhs)
    bc_functions = {"u": lambda x, y, z: _u_exact(x, y, z)[:dim], "t": lambda x, y, z: _t_exact(x, y, z)[:nt]}

    def gpu_forms(space, a):
        wf = LCEPrimalWeakForm(space)
        wf.functions = functions
        out = {"volume": ["form", range(len(space.discrete_spaces["u"].elements)), wf, a]}
        if with_bc:
            bf = LCEPrimalWeakFormBCDirichlet(space)
            bf.functions = bc_functions
            out["bc"] = ["bc_form", range(len(space.discrete_spaces["u"].bc_elements)), bf, a]
        return out

    def ref_forms_fn(space):
        out = [("form", ref_h1_multi.lce_primal, functions, len(space.discrete_spaces["u"].element_ids))]
        if with_bc:
            out.append(("bc_form", ref_h1_multi.lce_primal_bc_dirichlet, bc_functions,
                        len(space.discrete_spaces["u"].bc_element_ids)))
        return out

    alpha_fn = (lambda nd: np.cos(np.arange(nd) * 0.37) + 0.1) if alpha else None
    return Case(
        mesh, lambda g: {"u": (dim, dim, "Lagrange", k, g), "t": (dim, nt, "Lagrange", kt, g)},
        {"u": [1], "t": [1]}, {"u": btags if with_bc else [], "t": btags if with_bc else []}, gpu_forms, ref_forms_fn,
        alpha_fn=alpha_fn,
    )


def taylor_hood_case(n=3, varying=False, alpha=False, with_bc=True, mesh=None):
    from nfempy_b200.weak_forms import LETaylorHoodWeakForm, LETaylorHoodWeakFormBCDirichlet
    from oracle import ref_h1_multi

    if mesh is None:
        mesh = square_tri_mesh(n, jitter=0.2)
    if varying:
        mats = {"lambda": lambda x, y, z: 10.0 + x + 2.0 * y * y, "mu": lambda x, y, z: 2.0 + np.sin(x)}
    else:
        mats = {"lambda": lambda x, y, z: 1.0e3 + 0.0 * x, "mu": lambda x, y, z: 0.5 + 0.0 * x}
    functions = dict(mats, rhs=lambda x, y, z: _f_rhs(x, y, z)[:2])
    bc_functions = {"u": lambda x, y, z: _u_exact(x, y, z)[:2]}

    def gpu_forms(space, a):
        wf = LETaylorHoodWeakForm(space)
        wf.functions = functions
        out = {"volume": ["form", range(len(space.discrete_spaces["u"].elements)), wf, a]}
        if with_bc:
            bf = LETaylorHoodWeakFormBCDirichlet(space)
            bf.functions = bc_functions
            out["bc"] = ["bc_form", range(len(space.discrete_spaces["u"].bc_elements)), bf, a]
        return out

    def ref_forms_fn(space):
        out = [("form", ref_h1_multi.le_taylor_hood, functions, len(space.discrete_spaces["u"].element_ids))]
        if with_bc:
            out.append(("bc_form", ref_h1_multi.le_taylor_hood_bc_dirichlet, bc_functions,
                        len(space.discrete_spaces["u"].bc_element_ids)))
        return out

    alpha_fn = (lambda nd: np.cos(np.arange(nd) * 0.37) + 0.1) if alpha else None
    return Case(
        mesh, lambda g: {"u": (2, 2, "Lagrange", 2, g), "p": (2, 1, "Lagrange", 1, g)},
        {"u": [1], "p": [1]}, {"u": [2, 3, 4, 5] if with_bc else []}, gpu_forms, ref_forms_fn, alpha_fn=alpha_fn,
    )


def stress_constraint_case(dim=3, k=2, n=2, alpha=False):
    from nfempy_b200.weak_forms import LEPrimalStressConstraintWeakForm
    from oracle import ref_h1_multi

    mesh = box_tet_mesh(n, jitter=0.2, shuffle_local_nodes=True) if dim == 3 else square_tri_mesh(n, jitter=0.2)

    def s_target(x, y, z):
        rows = [[1.0 + x, 0.5 * y, z * x], [0.5 * y, 2.0 - y * y, 0.1 + 0.0 * x], [z * x, 0.1 + 0.0 * x, np.sin(z)]]
        return np.array([[rows[c][d] * np.ones_like(x) for d in range(dim)] for c in range(dim)])

    functions = {"s_target": s_target, "lambda": lambda x, y, z: 1.0 + x + 2.0 * y * y,
                 "mu": lambda x, y, z: 2.0 + np.sin(x)}

    def gpu_forms(space, a):
        wf = LEPrimalStressConstraintWeakForm(space)
        wf.functions = functions
        return {"volume": ["form", range(len(space.discrete_spaces["u"].elements)), wf, a]}

    def ref_forms_fn(space):
        return [("form", ref_h1_multi.le_stress_constraint, functions, len(space.discrete_spaces["u"].element_ids))]

    alpha_fn = (lambda nd: np.cos(np.arange(nd) * 0.37) + 0.1) if alpha else None
    return Case(mesh, lambda g: {"u": (dim, dim, "Lagrange", k, g)}, {"u": [1]}, {"u": []}, gpu_forms, ref_forms_fn,
                alpha_fn=alpha_fn)
