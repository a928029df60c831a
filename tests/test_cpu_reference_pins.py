"""The oracle (and the product's host layer) against fixtures produced by the REFERENCE'S OWN SOURCE.

``tests/golden/refcode_<case>.npz`` were written by ``tools/make_reference_pins.py``, which imports /root/reference/src
unmodified -- ``Mesh.build_conformal_mesh``, ``DoFMap``, ``DiscreteSpace`` / ``ProductSpace``, ``FiniteElement``,
``permute_and_transform``, ``transform_lower_to_higher``, the weak forms' ``evaluate_form`` (quadrature loop + forward AD) or
``evaluate_form_vectorized``, ``SequentialAssembler.scatter_forms`` -- over stand-ins for the third-party packages this
image lacks (``tools/ref_shims``: meshio, basix, auto_diff).  They pin every restated line ABOVE the basix boundary; basix's
own element conventions stay unpinned (the stand-in answers with the oracle's tables).  Lagrange P1 / P2 bases are unique,
so for configs 1-2 nothing is left unpinned but summation order."""
import os

import numpy as np
import pytest

from tests.parity_helpers import compare_csr
from tests.refcode_cases import cases

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = cases()
SKIP_ORACLE = set()
MD_CASES = {"md_elliptic_k1": (1, 5), "md_elliptic_k2": (2, None)}  # name -> (k, alpha seed); tools/make_reference_pins.py


def _fixture(name):
    path = os.path.join(GOLDEN, "refcode_%s.npz" % name)
    if not os.path.exists(path):
        pytest.fail("fixture missing: run tools/make_reference_pins.py %s in the build container" % name)
    return np.load(path)


def test_every_case_has_its_fixture_and_nothing_else():
    have = {f[len("refcode_"):-len(".npz")] for f in os.listdir(GOLDEN) if f.startswith("refcode_") and f.endswith(".npz")}
    assert have == set(CASES) | set(MD_CASES) | {"driver_cg_elasticity", "driver_mfe_cosserat", "driver_transport",
                                                       "driver_th_elasticity", "driver_mfe_elasticity"}


@pytest.mark.parametrize("name", sorted(set(CASES) - SKIP_ORACLE))
def test_oracle_assembles_what_the_reference_code_assembled(name):
    fx = _fixture(name)
    case = CASES[name]()
    got = case.run_oracle()
    compare_csr(got, (fx["indptr"], fx["indices"], fx["data"], fx["rhs"]), 1e-12)


def _rows(table):
    return [row[row >= 0] for row in table]


@pytest.mark.parametrize("name", sorted(CASES))
def test_numbering_and_destination_tables_equal_the_reference_codes(name):
    """Mesh numbering (cell ids by dimension, node tags, sub-cell ids, material ids) and the destination indexes of every
    cell and boundary cell: product (vectorised tables) and oracle (dict-based) against the reference's ``Mesh`` /
    ``DoFMap`` / ``ProductSpace``."""
    fx = _fixture(name)
    case = CASES[name]()
    # --- product mesh
    space = case.space()
    mesh = case.mesh
    n = len(fx["cell_dim"])
    assert mesh.max_cell_id() == n
    assert np.array_equal(mesh.cell_dimension(np.arange(n)), fx["cell_dim"])
    for d in range(mesh.dimension + 1):
        ids = np.flatnonzero(fx["cell_dim"] == d)
        assert np.array_equal(mesh.ids_of_dimension(d), ids)
        assert np.array_equal(mesh.node_tags_of(d), fx["node_tags"][ids, : d + 1])
        for e in range(d + 1):
            width = len(mesh.sub_cells_of(d, e)[0])
            assert np.array_equal(mesh.sub_cells_of(d, e), fx["sub_cells"][ids, e, :width])
            assert np.all(fx["sub_cells"][ids, e, width:] == -1)
    for cid in range(0, n, max(1, n // 97)):  # material ids through the cell objects (None <-> -1)
        material = mesh.cells[cid].material_id
        assert (-1 if material is None else int(material)) == int(fx["cell_material"][cid])
    # --- oracle mesh
    rspace = case.ref_space()
    rmesh = next(iter(rspace.discrete_spaces.values())).mesh
    for cid in range(n):
        cell = rmesh.cells[cid]
        assert cell.dimension == fx["cell_dim"][cid]
        assert np.array_equal(cell.node_tags, fx["node_tags"][cid, : cell.dimension + 1])
        for e in range(cell.dimension + 1):
            want = fx["sub_cells"][cid, e]
            assert np.array_equal(cell.sub_cells_ids[e], want[want >= 0])
    # --- destination tables
    assert space.n_dof == rspace.n_dof == len(fx["rhs"])
    # continuous cubic Lagrange: the product lists the two dofs of a mis-oriented edge in exchanged order instead of
    # exchanging the basis functions (DiscreteSpace.destination_table), so its rows equal the reference's as SETS there
    exchanged = any(s.family == "Lagrange" and s.k_order == 3 and not s.discontinuous for s in space.discrete_spaces.values())
    dest = _rows(fx["dest"])
    for i, want in enumerate(dest):
        got = space.destination_indexes(i)
        assert np.array_equal(np.sort(got), np.sort(want)) if exchanged else np.array_equal(got, want)
        assert np.array_equal(rspace.destination_indexes(i), want)
    for i, want in enumerate(_rows(fx["bc_dest"])):
        got = space.bc_destination_indexes(i)
        assert np.array_equal(np.sort(got), np.sort(want)) if exchanged else np.array_equal(got, want)
        assert np.array_equal(rspace.bc_destination_indexes(i), want)


@pytest.mark.parametrize("name", sorted(MD_CASES))
def test_oracle_md_system_equals_the_reference_codes(name):
    """Mixed-dimensional system (R/examples/md_elliptic_scalar/md_elliptic_scalar.py:205-342): the reference's mixed Poisson
    forms on the fractured box and on the fracture plane (stacked with ``dof_shift``) and its ``CouplingWeakForm`` over the
    interface triples, against the oracle's ``ref_md.assemble_md``; the product's triples and stacked destination indexes
    against the ones the reference's spaces gave."""
    from oracle import ref_forms, ref_md
    from tests import md_helpers as mh

    k, seed = MD_CASES[name]
    fx = _fixture(name)
    mesh, spaces, ref_spaces, triples, outer = mh.md_problem(2, k)
    assert np.array_equal(np.asarray(triples, dtype=np.int64), fx["triples"])
    n = len(fx["rhs"])
    assert n == spaces[1].dof_shift + spaces[1].n_dof
    alpha = np.zeros(n) if seed is None else np.random.default_rng(seed).standard_normal(n)
    u0, u1 = ref_spaces[0].discrete_spaces["u"], ref_spaces[1].discrete_spaces["u"]
    bc = {"p": mh.P_EXACT}
    got = ref_md.assemble_md(
        ref_spaces,
        [[("form", ref_forms.laplace_dual, {"rhs": mh.RHS, "kappa": mh.KAPPA_C0}, len(u0.element_ids)),
          ("bc_form", ref_forms.laplace_dual_bc_dirichlet, bc, outer)],
         [("form", ref_forms.laplace_dual, {"rhs": mh.RHS, "kappa": mh.KAPPA_C1}, len(u1.element_ids)),
          ("bc_form", ref_forms.laplace_dual_bc_dirichlet, bc, len(u1.bc_element_ids))]],
        triples, mh.COUPLING, alpha)
    compare_csr(got, (fx["indptr"], fx["indices"], fx["data"], fx["rhs"]), 1e-12)
    # the synthetic post-cut mesh as the reference's Mesh numbers it
    nc = len(fx["cell_dim"])
    assert mesh.max_cell_id() == nc and np.array_equal(mesh.cell_dimension(np.arange(nc)), fx["cell_dim"])
    for d in range(4):
        ids = np.flatnonzero(fx["cell_dim"] == d)
        assert np.array_equal(mesh.node_tags_of(d), fx["node_tags"][ids, : d + 1])


def test_committed_golden_fixtures_are_the_ones_the_reference_code_reproduced():
    """``tools/make_reference_pins.py --verify-golden`` ran the reference's source on the reference's own gmsh files
    (vector-P2 elasticity on example_3_3d_l_0 / l_1 / l_2.msh: 298, 1 325 and 4 261 tets of config 2's form list; four more systems on
    example_3_2d_l_0.msh) and compared with the ORACLE-made fixtures of tools/make_golden.py: same pattern, values and
    right-hand sides bit for bit (to round-off where the reference was run through its AD entry point and the oracle
    restates the vectorized one).  The recorded hashes tie that outcome to the files committed here."""
    import hashlib
    import json

    with open(os.path.join(GOLDEN, "refcode_verified.json")) as f:
        verified = json.load(f)
    assert len(verified) == 7
    probe = verified.pop("le_p2_ex3_3d_l2_probe.npz")  # 4 261 tets: fingerprint of the oracle's result
    z = np.load(os.path.join(GOLDEN, "le_p2_ex3_3d_l2_probe.npz"))
    assert hashlib.sha256(np.ascontiguousarray(z["rhs"]).tobytes()).hexdigest() == probe["sha256_rhs"]
    assert hashlib.sha256(np.ascontiguousarray(z["KV"]).tobytes()).hexdigest() == probe["sha256_KV"]
    assert probe["pattern_equal"] and probe["rhs_bit_identical"]
    assert max(probe["max_rel_KV"], probe["max_rel_diag"], probe["max_rel_abs_row_sum"]) <= 1e-13
    for key, entry in verified.items():
        file_name, _, prefix = key.partition(":")
        z = np.load(os.path.join(GOLDEN, file_name))
        data, rhs = (z[prefix + "_data"], z[prefix + "_rhs"]) if prefix else (z["data"], z["rhs"])
        assert hashlib.sha256(np.ascontiguousarray(data).tobytes()).hexdigest() == entry["sha256_data"], key
        assert hashlib.sha256(np.ascontiguousarray(rhs).tobytes()).hexdigest() == entry["sha256_rhs"], key
        assert entry["pattern_equal"] and entry["max_rel_values"] <= 1e-13 and entry["max_rel_rhs"] <= 1e-13, key
    assert verified["le_p2_ex3_3d_l0.npz"]["values_bit_identical"] and verified["le_p2_ex3_3d_l1.npz"]["values_bit_identical"]


@pytest.mark.parametrize("dim,k,fam", [(3, 1, "BDM"), (2, 2, "RT"), (3, 2, "RT")])
def test_oracle_error_norms_equal_the_reference_codes(dim, k, fam):
    """``l2_error``, ``div_error``, ``l2_error_projected`` (+ ``dof_shift``), ``l2_norm``, ``div_norm``, ``div_scaled_error``
    of R/src/postprocess, run by tools/make_reference_pins.py --postprocess on the spaces / fields of
    tests/test_gpu_postprocess.py (which compares the CUDA reductions with the same oracle functions)."""
    import json

    from oracle import ref_postprocess as rp
    from tests.parity_helpers import tensor_mixed_case
    from tests.test_gpu_postprocess import _fields

    with open(os.path.join(GOLDEN, "refcode_postprocess.json")) as f:
        want = json.load(f)["%dd_k%d_%s" % (dim, k, fam)]
    case = tensor_mixed_case(kind="lce", dim=dim, s_family=fam, m_family="RT", k=k, n=2 if dim == 2 else 1)
    rspace = case.ref_space()
    assert rspace.n_dof == want["n_dof"]
    alpha = np.cos(np.arange(rspace.n_dof) * 0.31) + 0.2
    fn = _fields(dim)
    fn_l = dict(fn, l=lambda x, y, z: 0.5 + 0.25 * x + 0.1 * y * y,
                grad_l=lambda x, y, z: np.array([0.25 * np.ones_like(x), 0.2 * y, 0.0 * x]))
    shifted = np.concatenate([np.full(5, 7.0), alpha])
    got = {
        "l2_error": rp.l2_error(dim, rspace, fn, alpha),
        "l2_error_skip_m_t": rp.l2_error(dim, rspace, fn, alpha, skip_fields=["m", "t"]),
        "div_error": rp.div_error(dim, rspace, fn, alpha),
        "l2_error_projected": rp.l2_error_projected(dim, rspace, alpha),
        "l2_error_projected_shift5": rp.l2_error_projected(dim, rspace, shifted, dof_shift=5),
        "l2_norm": rp.l2_norm(dim, rspace, fn),
        "div_norm": rp.div_norm(dim, rspace, fn),
        "div_scaled_error_without_scale": rp.div_scaled_error(dim, rspace, fn, alpha),
        "div_scaled_error": rp.div_scaled_error(dim, rspace, fn_l, alpha),
    }
    for key, values in got.items():
        assert len(values) == len(want[key]), key
        assert np.allclose(values, want[key], rtol=1e-13, atol=0.0), (key, values, want[key])


def test_oracle_filtered_scatter_equals_what_the_reference_driver_inserted():
    """``tools/make_reference_pins.py --driver`` ran the reference's example driver itself
    (R/examples/cauchy_elasticity/cg_elasticity_example_1.py ``primal_approximation`` + ``primal_postprocessing``, over the
    stand-ins incl. a petsc4py one) and kept what its OWN scatter closure put into the "sbaij" matrix: entries with
    ``~isclose(v, 1e-16)`` and ``row <= col``, explicit diagonal zeros (:92-145) -- the ``filtered_upper`` policy of
    ``GPUAssembler``.  The oracle's ``assemble_filtered`` must give the same CSR, the driver's solution must solve the oracle's
    system, and the oracle's ``l2_error`` of that solution must be the number the driver printed."""
    import scipy.sparse as sp

    from oracle import ref_postprocess
    from tests.refcode_cases import DRIVER_MATERIAL, driver_case, driver_displacement

    fx = np.load(os.path.join(GOLDEN, "refcode_driver_cg_elasticity.npz"))
    case = driver_case()
    indptr, indices, data, rhs = case.run_oracle(upper=True, explicit_diagonal=True)
    assert np.array_equal(indptr, fx["indptr"]) and np.array_equal(indices, fx["indices"])
    assert np.abs(data - fx["data"]).max() <= 1e-13 * np.abs(fx["data"]).max()
    n = len(rhs)
    upper = sp.csr_matrix((data, indices, indptr), shape=(n, n))
    assert (sp.tril(upper, -1)).nnz == 0
    full = upper + sp.triu(upper, 1).T
    assert np.abs(full @ fx["alpha"] + rhs).max() <= 1e-9 * np.abs(rhs).max()  # the driver solved K alpha = -rg
    rspace = case.ref_space()
    exact = {"u": driver_displacement(DRIVER_MATERIAL["lambda"], DRIVER_MATERIAL["mu"], 3)}
    assert np.allclose(ref_postprocess.l2_error(3, rspace, exact, fx["alpha"]), fx["l2_error"], rtol=1e-12, atol=0.0)
    # the product's host layer on the driver's tag lists (volume [1, 2, 3, 4], boundary [5, 6, 7, 8])
    space = case.space()
    assert space.n_dof == n
    u, ru = space.discrete_spaces["u"], rspace.discrete_spaces["u"]
    assert len(u.elements) == len(ru.element_ids) == 48 and len(u.bc_elements) == len(ru.bc_element_ids) > 0
    for i in range(len(u.bc_elements)):
        assert np.array_equal(space.bc_destination_indexes(i), rspace.bc_destination_indexes(i))


def test_oracle_equals_what_the_reference_cosserat_driver_assembled():
    """BASELINE config 4's driver (R/examples/cosserat_elasticity/mfe_cosserat_elasticity_example_3.py
    ``four_field_scaled_approximation``, run unmodified by tools/make_reference_pins.py --cosserat-driver, method "wc_rt" at
    k_order = 0): the saddle-point matrix its closure fills from ``LCEScaledDualWeakForm.evaluate_form_vectorized`` (filtered,
    upper, explicit diagonal zeros; the natural boundary form feeds the right-hand side only) and the preconditioner from
    ``LCEScaledRieszMapWeakForm`` (filtered, upper), against the oracle's ``assemble_filtered``; the driver's solution solves
    the oracle's system."""
    import scipy.sparse as sp

    from tests.refcode_cases import cosserat_driver_cases

    fx = np.load(os.path.join(GOLDEN, "refcode_driver_mfe_cosserat.npz"))
    case, riesz = cosserat_driver_cases()
    a = case.run_oracle(upper=True, explicit_diagonal=True)
    p = riesz.run_oracle(upper=True, explicit_diagonal=False)
    for got, key in ((a, "a"), (p, "p")):
        assert np.array_equal(got[0], fx[key + "_indptr"]) and np.array_equal(got[1], fx[key + "_indices"])
        assert np.abs(got[2] - fx[key + "_data"]).max() <= 1e-13 * np.abs(fx[key + "_data"]).max()
    n = len(a[3])
    upper = sp.csr_matrix((a[2], a[1], a[0]), shape=(n, n))
    full = upper + sp.triu(upper, 1).T
    assert np.abs(full @ fx["alpha"] + a[3]).max() <= 1e-9 * max(np.abs(a[3]).max(), 1.0)


def test_oracle_time_loop_reproduces_the_reference_transport_driver():
    """BASELINE config 5's driver (R/examples/parabolic_diffusive_problem/mixed_transport_model.py ``four_fields_formulation``,
    run unmodified by tools/make_reference_pins.py --transport-driver; RT_1 x RT_1 x P_0 x P_0 on six tetrahedra): ten
    implicit-Euler steps of Newton iterations on the Sundus forms, then the L2 errors at t = 1.  The same loop written with
    the oracle's forms, scatter and error norms must arrive at the driver's final coefficients and at the numbers it reports.
    Manufactured solution of the driver's 3-D branch (:177-216): p = c = t b(x) b(y) b(z), b(s) = s (1 - s), fluxes
    -grad p, unit coefficients."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl

    from oracle import ref_assembly, ref_forms, ref_postprocess
    from oracle.ref_mesh import RefMesh
    from oracle.ref_spaces import RefProductSpace

    from tests.refcode_cases import transport_driver_problem, transport_space

    fx = np.load(os.path.join(GOLDEN, "refcode_driver_transport.npz"))
    mesh, functions, bc_functions, delta_t, t_end = transport_driver_problem()
    rmesh = RefMesh(3, mesh.points, mesh._blocks)
    rmesh.build_conformal_mesh()
    space = transport_space(RefProductSpace, rmesh)
    n = space.n_dof
    assert n == len(fx["alpha"])
    mp = space.discrete_spaces["mp"]
    forms = [("time_dependent_form", ref_forms.sundus_dual, functions, len(mp.element_ids)),
             ("time_dependent_bc_form", ref_forms.sundus_dual_bc_dirichlet, bc_functions, len(mp.bc_element_ids))]
    previous, newton_total = np.zeros(n), 0
    for t in np.arange(delta_t, t_end + delta_t, delta_t):
        iterate = previous.copy()
        for _ in range(20):
            indptr, indices, data, res = ref_assembly.assemble_time_dependent(space, forms, (iterate, previous), t)
            if np.linalg.norm(res) < 1.0e-10:
                break
            iterate += spl.spsolve(sp.csr_matrix((data, indices, indptr), shape=(n, n)).tocsc(), -res)
            newton_total += 1
        else:
            raise AssertionError("Newton did not converge at t = %g" % t)
        previous = iterate
    assert 10 <= newton_total <= 60
    assert np.abs(previous - fx["alpha"]).max() <= 1e-9 * np.abs(fx["alpha"]).max()
    at_end = {k: (lambda x, y, z, f=f: f(x, y, z, t_end)) for k, f in bc_functions.items()}
    errors = ref_postprocess.l2_error(3, space, at_end, previous)
    assert np.allclose(errors, fx["l2_errors"], rtol=1e-8, atol=0.0)
    assert abs(errors[2] + errors[3] - float(fx["returned"])) <= 1e-8 * float(fx["returned"])


def test_oracle_filtered_scatter_equals_what_the_reference_taylor_hood_driver_inserted():
    """R/examples/cauchy_elasticity/th_elasticity_example_1.py ``mixed_approximation`` (run unmodified by
    tools/make_reference_pins.py --th-driver, u in P_2, p in discontinuous P_0, 2-D): the "filtered" policy with explicit
    diagonal zeros in full storage, as its closures fill the matrix (:99-145), against the oracle's ``assemble_filtered``;
    the driver's solution solves the oracle's system."""
    import scipy.sparse as sp

    from tests.refcode_cases import taylor_hood_driver_case

    fx = np.load(os.path.join(GOLDEN, "refcode_driver_th_elasticity.npz"))
    case = taylor_hood_driver_case()
    indptr, indices, data, rhs = case.run_oracle(upper=False, explicit_diagonal=True)
    assert np.array_equal(indptr, fx["indptr"]) and np.array_equal(indices, fx["indices"])
    assert np.abs(data - fx["data"]).max() <= 1e-13 * np.abs(fx["data"]).max()
    n = len(rhs)
    K = sp.csr_matrix((data, indices, indptr), shape=(n, n))
    assert np.abs(K @ fx["alpha"] + rhs).max() <= 1e-9 * np.abs(rhs).max()
    # the product's host layer on the driver's spaces and tag lists
    space, rspace = case.space(), case.ref_space()
    assert space.n_dof == n
    for i in range(len(space.discrete_spaces["u"].elements)):
        assert np.array_equal(space.destination_indexes(i), rspace.destination_indexes(i))
    for i in range(len(space.discrete_spaces["u"].bc_elements)):
        assert np.array_equal(space.bc_destination_indexes(i), rspace.bc_destination_indexes(i))


def test_oracle_equals_what_the_reference_mixed_elasticity_driver_assembled():
    """R/examples/cauchy_elasticity/mfe_elasticity_example_1.py ``three_field_approximation`` (run unmodified by
    tools/make_reference_pins.py --mfe-driver; BDM_1 rows x P_0 x P_0, symmetric storage): ``A`` from
    ``LEDualWeakForm.evaluate_form_vectorized`` (filtered, upper, explicit diagonal zeros; natural boundary form on the tags
    [5, 6, 7, 8]) and ``P`` from ``LERieszMapWeakForm`` (filtered, upper), against the oracle's ``assemble_filtered``; the
    driver's solution solves the oracle's system."""
    import scipy.sparse as sp

    from tests.refcode_cases import mixed_elasticity_driver_cases

    fx = np.load(os.path.join(GOLDEN, "refcode_driver_mfe_elasticity.npz"))
    case, riesz = mixed_elasticity_driver_cases()
    a = case.run_oracle(upper=True, explicit_diagonal=True)
    p = riesz.run_oracle(upper=True, explicit_diagonal=False)
    for got, key in ((a, "a"), (p, "p")):
        assert np.array_equal(got[0], fx[key + "_indptr"]) and np.array_equal(got[1], fx[key + "_indices"])
        assert np.abs(got[2] - fx[key + "_data"]).max() <= 1e-13 * np.abs(fx[key + "_data"]).max()
    n = len(a[3])
    upper = sp.csr_matrix((a[2], a[1], a[0]), shape=(n, n))
    full = upper + sp.triu(upper, 1).T
    assert np.abs(full @ fx["alpha"] + a[3]).max() <= 1e-9 * max(np.abs(a[3]).max(), 1.0)
    space, rspace = case.space(), case.ref_space()
    for i in range(len(space.discrete_spaces["s"].bc_elements)):
        assert np.array_equal(space.bc_destination_indexes(i), rspace.bc_destination_indexes(i))


def test_recorded_outcome_on_the_largest_reference_fixture():
    """A RECORD, not a re-run (the suite has no /root/reference): ``tools/make_reference_pins.py --verify-l3`` assembled
    config 2's matrix on example_3_3d_l_3.msh (28 736 tets, 10.4 M nonzeros) with the reference's own source -- five minutes --
    and compared it entry by entry with oracle/strong_cpu.py, the vectorised CPU assembly that bench.py checks the timed GPU
    step against at 998 250 tets."""
    import json

    with open(os.path.join(GOLDEN, "refcode_verified_l3.json")) as f:
        rec = json.load(f)
    assert rec["tets"] == 28736 and rec["pattern_equal"] and rec["max_rel_values_vs_strong_cpu"] <= 1e-13
