#!/usr/bin/env python
"""Runs the REFERENCE'S OWN SOURCE (/root/reference/src) on the parity cases of tests/parity_helpers.py and commits what it
produced as fixtures ``tests/golden/refcode_<case>.npz``.  Works only in the build container (/root/reference is not on the
GPU box); the fixtures travel.

    python tools/make_reference_pins.py            # all cases
    python tools/make_reference_pins.py le_p2_3d   # some

What runs is the reference's ``Mesh.build_conformal_mesh`` (mesh.py:239-339), ``MeshTopology``, ``DoFMap``,
``DiscreteSpace`` / ``ProductSpace.build_structures`` and ``destination_indexes``, ``FiniteElement.evaluate_mapping /
evaluate_basis`` with ``permute_and_transform`` and ``transform_lower_to_higher``, each weak form's ``evaluate_form`` and
``SequentialAssembler.scatter_forms`` (ScatterFormData.py:5-56) -- line for line, unmodified, imported from where it lies.
What does NOT run is the third-party code underneath, absent in this image: meshio, basix, auto_diff (and PETSc's Mat,
replaced by a dict of keys with ADD_VALUES).  tools/ref_shims/ holds the stand-ins and says what each answers with; the
basix stand-in answers with the oracle's element tables, so these fixtures pin every restated line ABOVE the basix boundary
(numbering, dof maps, geometry map, orientation handling, trace maps, the forms, the scatter) and leave basix's own
conventions unpinned.

Each fixture: ``indptr, indices, data, rhs`` of the assembled system, ``dest`` / ``bc_dest`` (every cell's and boundary
cell's destination indexes, ragged rows padded with -1), and the mesh numbering (``cell_dim, cell_material, node_tags,
sub_cells`` per cell id) the reference's ``Mesh`` built."""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference/src"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools", "ref_shims"))
sys.path.insert(0, REFERENCE)

import meshio  # noqa: E402  (the stand-in)

from tests import parity_helpers as ph  # noqa: E402,F401
from tests.refcode_cases import cases  # noqa: E402

# oracle restatement -> the reference class it restates
REF_CLASS = {
    "le_primal": ("le_primal_weak_form", "LEPrimalWeakForm"),
    "le_primal_bc_dirichlet": ("le_primal_weak_form", "LEPrimalWeakFormBCDirichlet"),
    "le_primal_bc_neumann": ("le_primal_weak_form", "LEPrimalWeakFormBCNeumann"),
    "le_primal_bc_normal_dirichlet": ("le_primal_weak_form", "LEPrimalWeakFormBCNormalDirichlet"),
    "laplace_dual": ("laplace_dual_weak_form", "LaplaceDualWeakForm"),
    "laplace_dual_bc_dirichlet": ("laplace_dual_weak_form", "LaplaceDualWeakFormBCDirichlet"),
    "l2_projector": ("l2_projector_weak_form", "L2ProjectorWeakForm"),
    "degenerate_elliptic": ("degenerate_elliptic_weak_form", "DegenerateEllipticWeakForm"),
    "degenerate_elliptic_bc_dirichlet": ("degenerate_elliptic_weak_form", "DegenerateEllipticWeakFormBCDirichlet"),
    "unscaled_elliptic": ("unscaled_elliptic_weak_form", "UnscaledEllipticWeakForm"),
    "unscaled_elliptic_bc_dirichlet": ("unscaled_elliptic_weak_form", "UnscaledEllipticWeakFormBCDirichlet"),
    "le_dual": ("le_dual_weak_form", "LEDualWeakForm"),
    "le_dual_bc_dirichlet": ("le_dual_weak_form", "LEDualWeakFormBCDirichlet"),
    "lce_dual": ("lce_dual_weak_form", "LCEDualWeakForm"),
    "lce_dual_bc_dirichlet": ("lce_dual_weak_form", "LCEDualWeakFormBCDirichlet"),
    "lce_scaled_dual": ("lce_scaled_dual_weak_form", "LCEScaledDualWeakForm"),
    "lce_scaled_dual_bc_dirichlet": ("lce_scaled_dual_weak_form", "LCEScaledDualWeakFormBCDirichlet"),
    "le_riesz_map": ("le_riesz_map_weak_form", "LERieszMapWeakForm"),
    "lce_riesz_map": ("lce_riesz_map_weak_form", "LCERieszMapWeakForm"),
    "lce_scaled_riesz_map": ("lce_scaled_riesz_map_weak_form", "LCEScaledRieszMapWeakForm"),
    "lce_primal": ("lce_primal_weak_form", "LCEPrimalWeakForm"),
    "lce_primal_bc_dirichlet": ("lce_primal_weak_form", "LCEPrimalWeakFormBCDirichlet"),
    "le_taylor_hood": ("le_taylor_hood_weak_form", "LETaylorHoodWeakForm"),
    "le_taylor_hood_bc_dirichlet": ("le_taylor_hood_weak_form", "LETaylorHoodWeakFormBCDirichlet"),
    "le_stress_constraint": ("le_primal_stress_constraint_weak_form", "LEPrimalStressConstraintWeakForm"),
    # (le_dual_weak_form.py's own LCEDualWeakFormBCNeumann, :580-, is stale: it reads a stored mapping its elements no
    # longer keep -- AttributeError 'MappingData' object has no attribute 'x' -- so the oracle's le_dual_bc_neumann has no
    # runnable counterpart)
    "lce_dual_bc_neumann": ("lce_dual_weak_form", "LCEDualWeakFormBCNeumann"),
}

# forms whose example drivers call ``evaluate_form_vectorized`` from their own scatter closure
# (R/examples/cosserat_elasticity/mfe_cosserat_elasticity_example_3.py:150-243); the oracle restates those bodies
VECTORIZED = {"le_dual", "lce_dual", "lce_scaled_dual", "le_riesz_map", "lce_riesz_map", "lce_scaled_riesz_map"}


class DictMat:
    """PETSc ``Mat`` stand-in for ``jac_g.setValue(row=, col=, value=, addv=True)`` (ScatterFormData.py:11)."""

    def __init__(self, n):
        self.n, self.entries = n, {}

    def setValue(self, row, col, value, addv=True):
        key = (int(row), int(col))
        self.entries[key] = (self.entries.get(key, 0.0) + float(value)) if addv else float(value)

    def csr(self):
        keys = sorted(self.entries)
        rows = np.array([k[0] for k in keys], dtype=np.int64)
        indices = np.array([k[1] for k in keys], dtype=np.int64)
        data = np.array([self.entries[k] for k in keys])
        indptr = np.zeros(self.n + 1, dtype=np.int64)
        np.add.at(indptr, rows + 1, 1)
        return np.cumsum(indptr), indices, data


def reference_system(case, vectorized=True):
    from assembly.SequentialAssembler import SequentialAssembler
    from mesh.mesh import Mesh
    from spaces.product_space import ProductSpace

    name = "memory://%d" % id(case)
    meshio.register(name, case.mesh.points, case.mesh._blocks)
    gmesh = Mesh(dimension=case.mesh.dimension, file_name=name)
    gmesh.build_conformal_mesh()
    data = case.data_fn(gmesh)
    space = ProductSpace(data)
    space.make_subspaces_discontinuous({k: bool(case.disc.get(k, False)) for k in data})
    space.build_structures(case.tags, case.btags)
    n = space.n_dof
    alpha = np.zeros(n) if case.alpha_fn is None else case.alpha_fn(n)
    oracle_space = case.ref_space()  # only for the form list (functions, counts) of the case
    assert oracle_space.n_dof == n, (oracle_space.n_dof, n)
    jac, res = DictMat(n), np.zeros(n)
    assembler = SequentialAssembler(space, jac, res)
    form_list = {}
    for j, (kind, oracle_fn, functions, n_items) in enumerate(case.ref_forms_fn(oracle_space)):
        if oracle_fn.__name__ == "proj":
            # ScaledToPhysicalL2Projection takes two local vectors and is scattered by its driver's own closure
            # (R/examples/degenerate_elliptic_scalar/degenerate_mixed_flow_model.py:172-189): restated here
            from weak_forms.scaled_to_physical_l2_projection import ScaledToPhysicalL2Projection

            form = ScaledToPhysicalL2Projection(space)
            form.functions = functions
            scaled = case.scaled_fn(n)
            for i in range(n_items):
                dest = space.destination_indexes(i)
                r_el, j_el = form.evaluate_form(i, alpha[dest], scaled[dest])
                res[dest] += r_el
                data, row, col = j_el.ravel(), np.repeat(dest, len(dest)), np.tile(dest, len(dest))
                for v, r, c in zip(data, row, col):
                    jac.setValue(row=r, col=c, value=v, addv=True)
            continue
        module, cls = REF_CLASS[oracle_fn.__name__]
        form = getattr(importlib.import_module("weak_forms." + module), cls)(space)
        form.functions = functions
        if vectorized and oracle_fn.__name__ in VECTORIZED:
            form.evaluate_form = form.evaluate_form_vectorized
        sequence = range(n_items) if isinstance(n_items, int) else list(n_items)
        form_list["f%d" % j] = [kind, sequence, form, alpha]
    assembler.form_to_input_list = form_list
    assembler.scatter_forms(measure_time_q=False)
    indptr, indices, values = jac.csr()
    # destination tables of every cell / boundary cell (all fields concatenated, as the scatter sees them)
    first = next(iter(space.discrete_spaces.values()))
    dest = [np.asarray(space.destination_indexes(i)) for i in range(len(first.elements))]
    bc_dest = [np.asarray(space.bc_destination_indexes(i)) for i in range(len(first.bc_elements))] if case.btags and any(case.btags.values()) else []
    return dict(indptr=indptr, indices=indices, data=values, rhs=res, dest=_ragged(dest), bc_dest=_ragged(bc_dest),
                **mesh_numbering(gmesh))


def reference_system_sundus(case):
    """Config 5's form classes (example-local: R/examples/parabolic_diffusive_problem/SundusWeakForm.py) through the
    reference's time-dependent scatter (ScatterFormData.py:28-33, 43-49): residual AND tangent at a non-zero iterate."""
    sys.path.insert(0, "/root/reference/examples/parabolic_diffusive_problem")
    from assembly.SequentialAssembler import SequentialAssembler
    from mesh.mesh import Mesh
    from spaces.product_space import ProductSpace
    from SundusWeakForm import SundusDualWeakForm, SundusDualWeakFormBCDirichlet

    name = "memory://%d" % id(case)
    meshio.register(name, case.mesh.points, case.mesh._blocks)
    gmesh = Mesh(dimension=case.mesh.dimension, file_name=name)
    gmesh.build_conformal_mesh()
    space = case._build(ProductSpace(case._data(gmesh)))  # the case's own calls, on the reference's ProductSpace
    n = space.n_dof
    alphas = case.alphas(n)
    jac, res = DictMat(n), np.zeros(n)
    wf, bf = SundusDualWeakForm(space), SundusDualWeakFormBCDirichlet(space)
    wf.functions, bf.functions = case.functions, case.bc_functions
    mp = space.discrete_spaces["mp"]
    assembler = SequentialAssembler(space, jac, res)
    assembler.form_to_input_list = {
        "volume": ["time_dependent_form", range(len(mp.elements)), wf, alphas, case.t],
        "bc": ["time_dependent_bc_form", range(len(mp.bc_elements)), bf, alphas, case.t],
    }
    assembler.scatter_forms(measure_time_q=False)
    indptr, indices, values = jac.csr()
    dest = [np.asarray(space.destination_indexes(i)) for i in range(len(mp.elements))]
    bc_dest = [np.asarray(space.bc_destination_indexes(i)) for i in range(len(mp.bc_elements))]
    return dict(indptr=indptr, indices=indices, data=values, rhs=res, dest=_ragged(dest), bc_dest=_ragged(bc_dest),
                **mesh_numbering(gmesh))


def reference_system_md(k, alpha_seed):
    """The mixed-dimensional system of R/examples/md_elliptic_scalar/md_elliptic_scalar.py:205-342 on the synthetic
    fractured box: the library's mixed Poisson forms on each dimension (through SequentialAssembler) and the example's
    ``CouplingWeakForm`` over the interface triples, scattered the way the driver's closure does (:284-302)."""
    sys.path.insert(0, "/root/reference/examples/md_elliptic_scalar")
    from assembly.SequentialAssembler import SequentialAssembler
    from CouplingWeakForm import CouplingWeakForm
    from mesh.mesh import Mesh
    from spaces.product_space import ProductSpace
    from weak_forms.laplace_dual_weak_form import LaplaceDualWeakForm, LaplaceDualWeakFormBCDirichlet

    from tests import md_helpers as mh

    mesh, spaces, _, triples, outer = mh.md_problem(2, k)
    name = "memory://md%d" % k
    meshio.register(name, mesh.points, mesh._blocks)
    gmesh = Mesh(dimension=3, file_name=name)
    gmesh.build_conformal_mesh()
    md_space = mh.md_spaces(gmesh, ProductSpace, k)
    n = md_space[1].dof_shift + md_space[1].n_dof
    assert n == spaces[1].dof_shift + spaces[1].n_dof
    alpha = np.zeros(n) if alpha_seed is None else np.random.default_rng(alpha_seed).standard_normal(n)
    jac, res = DictMat(n), np.zeros(n)
    bc = {"p": mh.P_EXACT}
    for space, fun, bc_items in ((md_space[0], {"rhs": mh.RHS, "kappa": mh.KAPPA_C0}, outer),
                                 (md_space[1], {"rhs": mh.RHS, "kappa": mh.KAPPA_C1}, None)):
        vol, bnd = LaplaceDualWeakForm(space), LaplaceDualWeakFormBCDirichlet(space)
        vol.functions, bnd.functions = fun, bc
        u = space.discrete_spaces["u"]
        assembler = SequentialAssembler(space, jac, res)
        assembler.form_to_input_list = {
            "vol": ["form", range(len(u.elements)), vol, alpha],
            "bc": ["bc_form", list(bc_items) if bc_items is not None else range(len(u.bc_elements)), bnd, alpha],
        }
        assembler.scatter_forms(measure_time_q=False)
    coupling = CouplingWeakForm(md_space)
    coupling.functions = mh.COUPLING
    for c1_idx, c0_p_idx, c0_n_idx in np.asarray(triples).tolist():
        dest = np.concatenate([md_space[0].bc_destination_indexes(c0_p_idx, "u"),
                               md_space[0].bc_destination_indexes(c0_n_idx, "u"),
                               md_space[1].destination_indexes(c1_idx, "p")])
        r_el, j_el = coupling.evaluate_form(c1_idx, c0_p_idx, c0_n_idx, alpha[dest])
        res[dest] += r_el
        data, row, col = j_el.ravel(), np.repeat(dest, len(dest)), np.tile(dest, len(dest))
        for v, r, c in zip(data, row, col):
            jac.setValue(row=r, col=c, value=v, addv=True)
    indptr, indices, values = jac.csr()
    return dict(indptr=indptr, indices=indices, data=values, rhs=res, triples=np.asarray(triples, dtype=np.int64),
                **mesh_numbering(gmesh))


MD_CASES = {"md_elliptic_k1": (1, 5), "md_elliptic_k2": (2, None)}  # name -> (k, alpha seed), as tests/test_gpu_md_coupling.py


def _ragged(rows):
    if not rows:
        return np.zeros((0, 0), dtype=np.int64)
    width = max(len(r) for r in rows)
    out = np.full((len(rows), width), -1, dtype=np.int64)
    for i, r in enumerate(rows):
        out[i, : len(r)] = r
    return out


def mesh_numbering(gmesh):
    """The reference Mesh's cells as arrays: dimension, material id (-1 = None), node tags and sub-cell ids of every
    dimension (padded with -1), indexed by cell id."""
    cells = gmesh.cells
    n = len(cells)
    cell_dim = np.array([c.dimension for c in cells], dtype=np.int64)
    material = np.array([-1 if c.material_id is None else int(c.material_id) for c in cells], dtype=np.int64)
    node_tags = _ragged([np.asarray(c.node_tags) for c in cells])
    sub = np.full((n, 4, 6), -1, dtype=np.int64)
    for i, c in enumerate(cells):
        assert c.id == i
        for d in range(c.dimension + 1):
            ids = np.asarray(c.sub_cells_ids[d])
            sub[i, d, : len(ids)] = ids
    return dict(cell_dim=cell_dim, cell_material=material, node_tags=node_tags, sub_cells=sub)


def compare_with_oracle(case, got):
    want = case.run_oracle()
    same_pattern = np.array_equal(want[0], got["indptr"]) and np.array_equal(want[1], got["indices"])
    if not same_pattern:
        return {"pattern_equal": False}
    scale = max(np.abs(want[2]).max(), 1e-300)
    rscale = max(np.abs(want[3]).max(), 1e-300)
    return {"pattern_equal": True, "max_rel_values": float(np.abs(want[2] - got["data"]).max() / scale),
            "max_rel_rhs": float(np.abs(want[3] - got["rhs"]).max() / rscale)}


def verify_golden():
    """The committed golden fixtures of tools/make_golden.py are ORACLE output.  This runs the reference's own source on the
    reference's own gmsh files (through its ``Mesh(dimension, file_name)`` constructor) and on the same form lists, compares
    with the committed arrays and records the outcome -- with the sha256 of each array it was compared with -- in
    ``tests/golden/refcode_verified.json`` (tests/test_cpu_reference_pins.py checks the hashes against the files)."""
    import contextlib
    import hashlib
    import io
    import json

    from assembly.SequentialAssembler import SequentialAssembler
    from mesh.mesh import Mesh
    from spaces.product_space import ProductSpace
    from weak_forms.le_primal_weak_form import LEPrimalWeakForm, LEPrimalWeakFormBCDirichlet

    from tests import golden_cases as gc

    msh_dir = "/root/reference/examples/cosserat_elasticity/gmsh_files/ex_3"
    out = {}

    def digest(a):
        return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()

    def record(key, got, want):
        same = np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
        entry = {"pattern_equal": bool(same), "sha256_data": digest(want[2]), "sha256_rhs": digest(want[3])}
        if same:
            entry["values_bit_identical"] = bool(np.array_equal(got[2], want[2]))
            entry["rhs_bit_identical"] = bool(np.array_equal(got[3], want[3]))
            entry["max_rel_values"] = float(np.abs(got[2] - want[2]).max() / np.abs(want[2]).max())
            entry["max_rel_rhs"] = float(np.abs(got[3] - want[3]).max() / np.abs(want[3]).max())
        out[key] = entry
        print(key, entry, flush=True)

    # vector-P2 elasticity + Dirichlet penalty on the reference's 3-D fixture, levels 0 and 1 (config 2's form list)
    for level in (0, 1, 2):
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            gmesh = Mesh(dimension=3, file_name=os.path.join(msh_dir, "example_3_3d_l_%d.msh" % level))
            gmesh.build_conformal_mesh()
            space = ProductSpace({"u": (3, 3, "Lagrange", 2, gmesh)})
            space.make_subspaces_discontinuous({"u": False})
            space.build_structures({"u": [1]}, {"u": [2, 3, 4, 5, 6, 7]})
            n = space.n_dof
            jac, res = DictMat(n), np.zeros(n)
            wf, bf = LEPrimalWeakForm(space), LEPrimalWeakFormBCDirichlet(space)
            wf.functions, bf.functions = gc.LE_FUNCTIONS, gc.LE_BC_FUNCTIONS
            us = space.discrete_spaces["u"]
            assembler = SequentialAssembler(space, jac, res)
            alpha = np.zeros(n)
            assembler.form_to_input_list = {"volume": ["form", range(len(us.elements)), wf, alpha],
                                            "bc": ["bc_form", range(len(us.bc_elements)), bf, alpha]}
            assembler.scatter_forms(measure_time_q=False)
        if level == 2:
            # 4 261 tets: the committed fixture is a fingerprint of the oracle's result (pattern, rhs, diagonal, K V for four
            # seeded vectors, row sums of |K|), not the 12 MB value array
            import scipy.sparse as sp

            z = np.load(os.path.join(gc.GOLDEN, "le_p2_ex3_3d_l2_probe.npz"))
            indptr, indices, values = jac.csr()
            same = np.array_equal(indptr, z["indptr"]) and np.array_equal(indices, z["indices"])
            K = sp.csr_matrix((values, indices, indptr), shape=(n, n))
            V = np.random.default_rng(int(z["seed"])).standard_normal((n, 4))
            entry = {"pattern_equal": bool(same), "sha256_rhs": digest(z["rhs"]), "sha256_KV": digest(z["KV"]),
                     "rhs_bit_identical": bool(np.array_equal(res, z["rhs"])),
                     "max_rel_KV": float(np.abs(K @ V - z["KV"]).max() / np.abs(z["KV"]).max()),
                     "max_rel_diag": float(np.abs(K.diagonal() - z["diag"]).max() / np.abs(z["diag"]).max()),
                     "max_rel_abs_row_sum": float(np.abs(np.asarray(abs(K).sum(axis=1)).ravel() - z["abs_row_sum"]).max()
                                                  / np.abs(z["abs_row_sum"]).max())}
            out["le_p2_ex3_3d_l2_probe.npz"] = entry
            print("le_p2_ex3_3d_l2_probe.npz", entry, flush=True)
        else:
            z = np.load(os.path.join(gc.GOLDEN, "le_p2_ex3_3d_l%d.npz" % level))
            record("le_p2_ex3_3d_l%d.npz" % level, jac.csr() + (res,), (z["indptr"], z["indices"], z["data"], z["rhs"]))
        print("  (%d tets, %.0f s)" % (len(us.elements), time.perf_counter() - t0), flush=True)
    # the four systems of multi_ex3_2d_l0.npz on the reference's 2-D fixture
    z = np.load(os.path.join(gc.GOLDEN, "multi_ex3_2d_l0.npz"))
    for key, (case, _) in gc.multi_cases().items():
        with contextlib.redirect_stdout(io.StringIO()):
            got = reference_system(case, vectorized=False)
        record("multi_ex3_2d_l0.npz:" + key, (got["indptr"], got["indices"], got["data"], got["rhs"]),
               tuple(z["%s_%s" % (key, part)] for part in ("indptr", "indices", "data", "rhs")))
    with open(os.path.join(gc.GOLDEN, "refcode_verified.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


POSTPROCESS_CASES = [(3, 1, "BDM"), (2, 2, "RT"), (3, 2, "RT")]  # as tests/test_gpu_postprocess.py


def postprocess_pins():
    """The reference's error / solution norms (R/src/postprocess/l2_error_post_processor.py:7-117, 185-305,
    solution_norms_post_processor.py:7-78) on the mixed Cosserat spaces and fields of tests/test_gpu_postprocess.py, for a
    seeded coefficient vector -> tests/golden/refcode_postprocess.json."""
    import contextlib
    import io
    import json

    from mesh.mesh import Mesh
    from postprocess.l2_error_post_processor import div_error, div_scaled_error, l2_error, l2_error_projected
    from postprocess.solution_norms_post_processor import div_norm, l2_norm
    from spaces.product_space import ProductSpace

    from tests.test_gpu_postprocess import _fields

    out = {}
    for dim, k, fam in POSTPROCESS_CASES:
        case = ph.tensor_mixed_case(kind="lce", dim=dim, s_family=fam, m_family="RT", k=k, n=2 if dim == 2 else 1)
        with contextlib.redirect_stdout(io.StringIO()):
            name = "memory://pp%d%d%s" % (dim, k, fam)
            meshio.register(name, case.mesh.points, case.mesh._blocks)
            gmesh = Mesh(dimension=dim, file_name=name)
            gmesh.build_conformal_mesh()
            data = case.data_fn(gmesh)
            space = ProductSpace(data)
            space.make_subspaces_discontinuous({f: bool(case.disc.get(f, False)) for f in data})
            space.build_structures(case.tags, case.btags)
        alpha = np.cos(np.arange(space.n_dof) * 0.31) + 0.2
        fn = _fields(dim)
        fn_l = dict(fn, l=lambda x, y, z: 0.5 + 0.25 * x + 0.1 * y * y,
                    grad_l=lambda x, y, z: np.array([0.25 * np.ones_like(x), 0.2 * y, 0.0 * x]))
        shifted = np.concatenate([np.full(5, 7.0), alpha])
        entry = {
            "n_dof": int(space.n_dof),
            "l2_error": [float(v) for v in l2_error(dim, space, fn, alpha)],
            "l2_error_skip_m_t": [float(v) for v in l2_error(dim, space, fn, alpha, skip_fields=["m", "t"])],
            "div_error": [float(v) for v in div_error(dim, space, fn, alpha)],
            "l2_error_projected": [float(v) for v in l2_error_projected(dim, space, alpha)],
            "l2_error_projected_shift5": [float(v) for v in l2_error_projected(dim, space, shifted, dof_shift=5)],
            "l2_norm": [float(v) for v in l2_norm(dim, space, fn)],
            "div_norm": [float(v) for v in div_norm(dim, space, fn)],
            "div_scaled_error_without_scale": [float(v) for v in div_scaled_error(dim, space, fn, alpha)],
            "div_scaled_error": [float(v) for v in div_scaled_error(dim, space, fn_l, alpha)],
        }
        out["%dd_k%d_%s" % (dim, k, fam)] = entry
        print(dim, k, fam, {key: len(v) if isinstance(v, list) else v for key, v in entry.items()}, flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "refcode_postprocess.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


def driver_pin():
    """Runs the reference's EXAMPLE DRIVER itself -- ``primal_approximation`` and ``primal_postprocessing`` of
    R/examples/cauchy_elasticity/cg_elasticity_example_1.py, unmodified -- over the stand-ins (plus tools/ref_shims/petsc4py)
    and keeps what its own scatter closure inserted into the "sbaij" matrix: the filtered, upper-triangular CSR with explicit
    diagonal zeros, the solution vector and the L2 error it reports -> tests/golden/refcode_driver_cg_elasticity.npz.  The
    example's strong-solution module is replaced by tests/refcode_cases.py's functions so that the tests need no reference."""
    import contextlib
    import io
    import types

    from petsc4py import PETSc

    from tests import refcode_cases as rc

    stub = types.ModuleType("strong_solution_elasticity_example_1")
    stub.displacement, stub.rhs = rc.driver_displacement, rc.driver_rhs
    sys.modules["strong_solution_elasticity_example_1"] = stub
    sys.path.insert(0, "/root/reference/examples/cauchy_elasticity")
    import cg_elasticity_example_1 as driver

    case = rc.driver_case()
    meshio.register("memory://driver", case.mesh.points, case.mesh._blocks)
    method = ("FEM", {"u": ("Lagrange", 2)})
    with contextlib.redirect_stdout(io.StringIO()):
        gmesh = driver.create_mesh_from_file("memory://driver", 3)
        PETSc.ASSEMBLED.clear()
        alpha, _ = driver.primal_approximation(rc.DRIVER_MATERIAL, method, gmesh)
        n_dof, errors = driver.primal_postprocessing(rc.DRIVER_MATERIAL, method, gmesh, alpha)
    indptr, indices, data = PETSc.ASSEMBLED[-1].getValuesCSR()
    want = case.run_oracle(upper=True, explicit_diagonal=True)
    same = np.array_equal(want[0], indptr) and np.array_equal(want[1], indices)
    print("driver cg_elasticity_example_1: n_dof %d nnz %d pattern_equal %s max_rel_values %s l2_error %.6e" % (
        n_dof, len(data), same, float(np.abs(want[2] - data).max() / np.abs(want[2]).max()) if same else None, errors[0]))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "refcode_driver_cg_elasticity.npz"), indptr=indptr,
                        indices=indices, data=data, alpha=np.asarray(alpha), l2_error=np.asarray(errors), **mesh_numbering(gmesh))


def taylor_hood_driver_pin():
    """``mixed_approximation`` of R/examples/cauchy_elasticity/th_elasticity_example_1.py, unmodified, method at k_order = 2:
    the matrix its closures fill ("filtered" policy with explicit diagonal zeros, full storage) and its solution ->
    tests/golden/refcode_driver_th_elasticity.npz."""
    import contextlib
    import io
    import types

    from petsc4py import PETSc

    from tests import refcode_cases as rc

    stub = types.ModuleType("strong_solution_elasticity_example_1")
    stub.displacement, stub.rhs, stub.pressure = rc.driver_displacement, rc.driver_rhs, rc.driver_pressure
    sys.modules["strong_solution_elasticity_example_1"] = stub
    sys.path.insert(0, "/root/reference/examples/cauchy_elasticity")
    import th_elasticity_example_1 as driver

    case = rc.taylor_hood_driver_case()
    meshio.register("memory://th", case.mesh.points, case.mesh._blocks)
    method = next(iter(driver.method_definition(2)))
    with contextlib.redirect_stdout(io.StringIO()):
        gmesh = driver.create_mesh_from_file("memory://th", 2)
        PETSc.ASSEMBLED.clear()
        alpha = driver.mixed_approximation(rc.DRIVER_MATERIAL, method, gmesh)
    alpha = alpha[0] if isinstance(alpha, tuple) else alpha
    indptr, indices, data = PETSc.ASSEMBLED[-1].getValuesCSR()
    want = case.run_oracle(upper=False, explicit_diagonal=True)
    same = np.array_equal(want[0], indptr) and np.array_equal(want[1], indices)
    print("driver th_elasticity_example_1: n_dof %d nnz %d pattern_equal %s max_rel_values %s" % (
        len(want[3]), len(data), same, float(np.abs(want[2] - data).max() / np.abs(want[2]).max()) if same else None))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "refcode_driver_th_elasticity.npz"), indptr=indptr,
                        indices=indices, data=data, alpha=np.asarray(alpha))


def mixed_elasticity_driver_pin():
    """``three_field_approximation`` of R/examples/cauchy_elasticity/mfe_elasticity_example_1.py (:68-294, unmodified, method
    at k_order = 0, symmetric storage): ``A`` (LEDualWeakForm + the natural boundary form) and ``P`` (LERieszMapWeakForm) as
    its closures fill them, and its solution -> tests/golden/refcode_driver_mfe_elasticity.npz."""
    import contextlib
    import io
    import types

    from petsc4py import PETSc

    from tests import refcode_cases as rc

    case, riesz = rc.mixed_elasticity_driver_cases()
    forms = case.ref_forms_fn(case.ref_space())
    functions, bc_functions = forms[0][2], forms[1][2]
    stub = types.ModuleType("strong_solution_elasticity_example_1")
    stub.displacement = lambda m_lambda, m_mu, dim: bc_functions["u"]
    stub.rotation = lambda m_lambda, m_mu, dim: bc_functions["t"]
    stub.stress = lambda m_lambda, m_mu, dim: (lambda x, y, z: np.zeros((dim, dim)))
    stub.stress_divergence = stub.stress
    stub.rhs = lambda m_lambda, m_mu, dim: functions["rhs"]
    sys.modules["strong_solution_elasticity_example_1"] = stub
    sys.path.insert(0, "/root/reference/examples/cauchy_elasticity")
    import mfe_elasticity_example_1 as driver

    meshio.register("memory://mfe", case.mesh.points, case.mesh._blocks)
    method = next(iter(driver.method_definition(0)))
    with contextlib.redirect_stdout(io.StringIO()):
        gmesh = driver.create_mesh_from_file("memory://mfe", 3)
        PETSc.ASSEMBLED.clear()
        result = driver.three_field_approximation(rc.DRIVER_MATERIAL, method, gmesh, symmetric_solver_q=True)
    alpha = result[0] if isinstance(result, tuple) else result
    (a_ptr, a_idx, a_val), (p_ptr, p_idx, p_val) = [m.getValuesCSR() for m in PETSc.ASSEMBLED[-2:]]
    want_a = case.run_oracle(upper=True, explicit_diagonal=True)
    want_p = riesz.run_oracle(upper=True, explicit_diagonal=False)
    for label, want, got in (("A", want_a, (a_ptr, a_idx, a_val)), ("P", want_p, (p_ptr, p_idx, p_val))):
        same = np.array_equal(want[0], got[0]) and np.array_equal(want[1], got[1])
        print("driver mfe_elasticity %s: nnz %d pattern_equal %s max_rel_values %s" % (
            label, len(got[2]), same, float(np.abs(want[2] - got[2]).max() / np.abs(want[2]).max()) if same else None))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "refcode_driver_mfe_elasticity.npz"), a_indptr=a_ptr,
                        a_indices=a_idx, a_data=a_val, p_indptr=p_ptr, p_indices=p_idx, p_data=p_val, alpha=np.asarray(alpha))


def cosserat_driver_pin():
    """BASELINE config 4's driver itself: ``four_field_scaled_approximation`` of
    R/examples/cosserat_elasticity/mfe_cosserat_elasticity_example_3.py (:93-316, unmodified, "wc_rt" method at k_order = 0,
    symmetric storage) over the stand-ins: the saddle-point matrix ``A`` and the Riesz-map preconditioner ``P`` exactly as its
    scatter closures insert them, and the solution of its (here: direct) solve -> refcode_driver_mfe_cosserat.npz.  The
    example's strong-solution module is replaced by the coefficient / boundary functions of tests/parity_helpers.py."""
    import contextlib
    import io
    import types

    from petsc4py import PETSc

    from tests import refcode_cases as rc

    case, riesz = rc.cosserat_driver_cases()
    space0 = case.ref_space()
    (_, _, functions, _), (_, _, bc_functions, _) = case.ref_forms_fn(space0)
    stub = types.ModuleType("strong_solution_cosserat_elasticity_example_3")
    stub.displacement = lambda material, dim: bc_functions["u"]
    stub.rotation = lambda material, dim: bc_functions["t"]
    unused = lambda material, dim: (lambda x, y, z: np.zeros((dim, dim)))  # noqa: E731  (post-processing only)
    stub.stress = stub.couple_stress_scaled = stub.stress_divergence = stub.couple_stress_divergence_scaled = unused
    stub.rhs_scaled = lambda material, dim: functions["rhs"]
    stub.get_material_functions = lambda material, dim: {k: v for k, v in functions.items() if k != "rhs"}
    sys.modules["strong_solution_cosserat_elasticity_example_3"] = stub
    sys.path.insert(0, "/root/reference/examples/cosserat_elasticity")
    import mfe_cosserat_elasticity_example_3 as driver

    meshio.register("memory://cosserat", case.mesh.points, case.mesh._blocks)
    method = next(iter(driver.method_definition(0)))
    assert method[0] == "wc_rt"
    with contextlib.redirect_stdout(io.StringIO()):
        gmesh = driver.create_mesh_from_file("memory://cosserat", 3)
        PETSc.ASSEMBLED.clear()
        alpha, _ = driver.four_field_scaled_approximation(method, gmesh, symmetric_solver_q=True)
    (a_ptr, a_idx, a_val), (p_ptr, p_idx, p_val) = [m.getValuesCSR() for m in PETSc.ASSEMBLED[-2:]]
    want_a = case.run_oracle(upper=True, explicit_diagonal=True)
    want_p = riesz.run_oracle(upper=True, explicit_diagonal=False)
    for label, want, got in (("A", want_a, (a_ptr, a_idx, a_val)), ("P", want_p, (p_ptr, p_idx, p_val))):
        same = np.array_equal(want[0], got[0]) and np.array_equal(want[1], got[1])
        print("driver mfe_cosserat %s: nnz %d pattern_equal %s max_rel_values %s" % (
            label, len(got[2]), same, float(np.abs(want[2] - got[2]).max() / np.abs(want[2]).max()) if same else None))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "refcode_driver_mfe_cosserat.npz"), a_indptr=a_ptr,
                        a_indices=a_idx, a_data=a_val, p_indptr=p_ptr, p_indices=p_idx, p_data=p_val, alpha=np.asarray(alpha))


def transport_driver_pin():
    """BASELINE config 5's driver itself: ``four_fields_formulation`` of
    R/examples/parabolic_diffusive_problem/mixed_transport_model.py (:104-387, unmodified; method "mixed_rt" at k_order = 0 =
    RT_1 x RT_1 x P_0 x P_0): ten implicit-Euler steps, each a Newton loop that re-assembles the Sundus forms into the SAME
    matrix (``scale(0.0)`` + ADD_VALUES) until the residual norm drops below 1e-10, then the L2 errors at t = 1 -- on the
    jittered Kuhn box of six tetrahedra.  Kept: the final coefficient vector, the four L2 errors, the returned number ->
    tests/golden/refcode_driver_transport.npz.  (``gmsh`` is an empty stand-in: the driver's mesher is not used.)"""
    import contextlib
    import io
    import types

    from nfempy_b200.mesh import box_tet_mesh

    sys.modules.setdefault("gmsh", types.ModuleType("gmsh"))
    sys.path.insert(0, "/root/reference/examples/parabolic_diffusive_problem")
    import mixed_transport_model as driver

    mesh = box_tet_mesh(1, jitter=0.2, shuffle_local_nodes=True)
    meshio.register("memory://transport", mesh.points, mesh._blocks)
    kept = {}
    true_l2_error = driver.l2_error

    def l2_error(dim, fe_space, functions, alpha):
        kept["alpha"], kept["errors"] = np.array(alpha), np.array(true_l2_error(dim, fe_space, functions, alpha))
        return kept["errors"]

    driver.l2_error = l2_error
    method = next(iter(driver.method_definition(0)))
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        gmesh = driver.Mesh(dimension=3, file_name="memory://transport")
        gmesh.build_conformal_mesh()
        returned = driver.four_fields_formulation(method, gmesh)
    print("driver mixed_transport_model: n_dof %d, L2 errors %s, returned %.12e (%.0f s)" % (
        len(kept["alpha"]), kept["errors"], returned, time.perf_counter() - t0))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "refcode_driver_transport.npz"), alpha=kept["alpha"],
                        l2_errors=kept["errors"], returned=np.array(returned))


def verify_level3():
    """The largest reference fixture, example_3_3d_l_3.msh (28 736 tets): the reference's source assembles the vector-P2
    elasticity VOLUME form (config 2's matrix) cell by cell -- about four minutes -- and the result is compared, entry by
    entry, with oracle/strong_cpu.py (the vectorised CPU assembly bench.py checks the timed GPU step against), fed with the
    numbering the reference's own spaces produced.  Outcome appended to tests/golden/refcode_verified.json (a record: the test
    suite cannot repeat it without /root/reference, this mode can)."""
    import contextlib
    import io
    import json

    from assembly.SequentialAssembler import SequentialAssembler
    from mesh.mesh import Mesh
    from spaces.product_space import ProductSpace
    from weak_forms.le_primal_weak_form import LEPrimalWeakForm

    from oracle import strong_cpu
    from tests import golden_cases as gc

    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        gmesh = Mesh(dimension=3, file_name="/root/reference/examples/cosserat_elasticity/gmsh_files/ex_3/example_3_3d_l_3.msh")
        gmesh.build_conformal_mesh()
        space = ProductSpace({"u": (3, 3, "Lagrange", 2, gmesh)})
        space.make_subspaces_discontinuous({"u": False})
        space.build_structures({"u": [1]}, {"u": [2, 3, 4, 5, 6, 7]})
        n = space.n_dof
        jac, res = DictMat(n), np.zeros(n)
        wf = LEPrimalWeakForm(space)
        wf.functions = gc.LE_FUNCTIONS
        us = space.discrete_spaces["u"]
        assembler = SequentialAssembler(space, jac, res)
        assembler.form_to_input_list = {"volume": ["form", range(len(us.elements)), wf, np.zeros(n)]}
        assembler.scatter_forms(measure_time_q=False)
    t1 = time.perf_counter()
    indptr, indices, values = jac.csr()
    dest = np.array([space.destination_indexes(i) for i in range(len(us.elements))], dtype=np.int64)
    cell_nodes = np.array([el.data.cell.node_tags for el in us.elements], dtype=np.int64)
    s_ptr, s_idx, s_val = strong_cpu.assemble_elasticity(np.asarray(gmesh.points), cell_nodes, dest, n, 1.0, 1.0)
    same = np.array_equal(indptr, s_ptr) and np.array_equal(indices, s_idx)
    row = np.repeat(np.arange(n), np.diff(indptr))
    row_max = np.zeros(n)
    np.maximum.at(row_max, row, np.abs(values))
    entry = {"tets": int(len(us.elements)), "n_dof": int(n), "nnz": int(len(values)), "pattern_equal": bool(same),
             "max_rel_values_vs_strong_cpu": float((np.abs(values - s_val) / np.maximum(np.abs(values), row_max[row])).max())
             if same else None, "reference_source_seconds": round(t1 - t0, 1)}
    print("le_p2 volume form on example_3_3d_l_3.msh:", entry, flush=True)
    path = os.path.join(gc.GOLDEN, "refcode_verified_l3.json")
    with open(path, "w") as f:
        json.dump(entry, f, indent=1, sort_keys=True)


def main():
    if sys.argv[1:] == ["--verify-l3"]:
        return verify_level3()
    if sys.argv[1:] == ["--mfe-driver"]:
        return mixed_elasticity_driver_pin()
    if sys.argv[1:] == ["--th-driver"]:
        return taylor_hood_driver_pin()
    if sys.argv[1:] == ["--transport-driver"]:
        return transport_driver_pin()
    if sys.argv[1:] == ["--cosserat-driver"]:
        return cosserat_driver_pin()
    if sys.argv[1:] == ["--driver"]:
        return driver_pin()
    if sys.argv[1:] == ["--verify-golden"]:
        return verify_golden()
    if sys.argv[1:] == ["--postprocess"]:
        return postprocess_pins()
    wanted = sys.argv[1:]
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, (k, seed) in MD_CASES.items():
        if wanted and name not in wanted:
            continue
        import contextlib
        import io

        from oracle import ref_forms, ref_md
        from tests import md_helpers as mh

        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            got = reference_system_md(k, seed)
        _, _, ref_spaces, triples, outer = mh.md_problem(2, k)
        n = len(got["rhs"])
        alpha = np.zeros(n) if seed is None else np.random.default_rng(seed).standard_normal(n)
        u0, u1 = ref_spaces[0].discrete_spaces["u"], ref_spaces[1].discrete_spaces["u"]
        bc = {"p": mh.P_EXACT}
        want = ref_md.assemble_md(
            ref_spaces,
            [[("form", ref_forms.laplace_dual, {"rhs": mh.RHS, "kappa": mh.KAPPA_C0}, len(u0.element_ids)),
              ("bc_form", ref_forms.laplace_dual_bc_dirichlet, bc, outer)],
             [("form", ref_forms.laplace_dual, {"rhs": mh.RHS, "kappa": mh.KAPPA_C1}, len(u1.element_ids)),
              ("bc_form", ref_forms.laplace_dual_bc_dirichlet, bc, len(u1.bc_element_ids))]],
            triples, mh.COUPLING, alpha)
        same = np.array_equal(want[0], got["indptr"]) and np.array_equal(want[1], got["indices"])
        report = {"pattern_equal": same}
        if same:
            report["max_rel_values"] = float(np.abs(want[2] - got["data"]).max() / np.abs(want[2]).max())
            report["max_rel_rhs"] = float(np.abs(want[3] - got["rhs"]).max() / np.abs(want[3]).max())
        np.savez_compressed(os.path.join(out_dir, "refcode_%s.npz" % name), **got)
        print("%-28s n_dof %6d nnz %8d  vs oracle: %s  (%.1f s)" % (name, n, len(got["data"]), report,
                                                                   time.perf_counter() - t0), flush=True)
    for name, make in cases().items():
        if wanted and name not in wanted:
            continue
        t0 = time.perf_counter()
        import contextlib
        import io

        case = make()
        try:
            with contextlib.redirect_stdout(io.StringIO()):  # the reference prints timing lines while it builds spaces
                if isinstance(case, ph.SundusCase):
                    got = reference_system_sundus(case)
                else:
                    got = reference_system(case, vectorized=not name.endswith("_ad"))
        except Exception as exc:  # keep going: one stale reference class must not hide the others
            print("%-28s FAILED in the reference code: %s: %s" % (name, type(exc).__name__, exc), flush=True)
            continue
        report = compare_with_oracle(case, got)
        np.savez_compressed(os.path.join(out_dir, "refcode_%s.npz" % name), **got)
        print("%-28s n_dof %6d nnz %8d  vs oracle: %s  (%.1f s)" % (name, len(got["rhs"]), len(got["data"]), report,
                                                                   time.perf_counter() - t0), flush=True)


if __name__ == "__main__":
    main()
