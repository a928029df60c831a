"""Mixed-dimensional pieces on the CPU: the fractured mesh generator, the stacked destination tables of product and
oracle, and the oracle's restatement of the coupling form (symmetry, zero blocks, unit flux moments)."""
import numpy as np

from oracle import ref_md
from tests.md_helpers import COUPLING, md_problem


def test_fractured_box_is_cut_along_the_plane():
    mesh, spaces, ref_spaces, triples, outer = md_problem(2)
    u0 = spaces[0].discrete_spaces["u"]
    tets = u0.cell_nodes()
    z = mesh.points[:, 2]
    above = mesh.points[tets].mean(axis=1)[:, 2] > 0.5
    plane = np.isclose(z, 0.5)
    # no node of the plane is shared between the two sides or with the fracture cells
    up, down = np.unique(tets[above]), np.unique(tets[~above])
    frac = np.unique(spaces[1].discrete_spaces["u"].cell_nodes())
    assert plane[frac].all() and not set(up) & set(down) and not set(frac) & (set(up) | set(down))
    assert len(triples) == 8 and len(outer) == 48 and len(u0.bc_element_ids) == 64
    # the three copies of an interface cell coincide point by point, vertex order included
    pos = mesh.face_nodes[u0.bc_element_ids[triples[:, 1]] - mesh.offset[2]]
    neg = mesh.face_nodes[u0.bc_element_ids[triples[:, 2]] - mesh.offset[2]]
    c1 = spaces[1].discrete_spaces["u"].cell_nodes()[triples[:, 0]]
    assert np.allclose(mesh.points[pos], mesh.points[c1]) and np.allclose(mesh.points[neg], mesh.points[c1])
    # one side above, the other below
    n0 = u0.mesh_topology.first_neighbours(u0.bc_element_ids[triples[:, 1]])[0]
    n1 = u0.mesh_topology.first_neighbours(u0.bc_element_ids[triples[:, 2]])[0]
    assert above[n0].all() and not above[n1].any()


def test_stacked_destination_tables_match_the_oracle():
    from nfempy_b200.weak_forms import CouplingWeakForm

    mesh, spaces, ref_spaces, triples, outer = md_problem(2)
    assert [s.n_dof for s in spaces] == [s.n_dof for s in ref_spaces] and spaces[1].dof_shift == spaces[0].n_dof
    form = CouplingWeakForm(spaces)
    got = form.interface_dest(triples)
    for t, row in zip(triples, got):
        want = np.concatenate([
            ref_spaces[0].bc_destination_indexes(int(t[1]), "u"), ref_spaces[0].bc_destination_indexes(int(t[2]), "u"),
            ref_spaces[1].destination_indexes(int(t[0]), "p"),
        ])
        assert np.array_equal(row, want)
    assert got[:, -1].min() >= spaces[0].n_dof  # the fracture potential sits behind the first space


def test_oracle_coupling_block_structure():
    mesh, spaces, ref_spaces, triples, outer = md_problem(2)
    a = np.random.default_rng(0).standard_normal(3)
    r, j = ref_md.coupling(ref_spaces, COUPLING, *[int(v) for v in triples[3]], a)
    assert np.abs(j - j.T).max() <= 1e-15 and np.abs(r - j @ a).max() <= 1e-15
    assert j[0, 1] == 0.0 and j[2, 2] == 0.0
    assert np.allclose(np.abs(j[0:2, 2]), 1.0, atol=1e-13)  # int psi.n over the facet of a lowest-order RT function
    assert j[0, 0] > 0 and np.isclose(j[0, 0], j[1, 1])


def test_md_oracle_reproduces_uniform_flow_across_the_fracture():
    """Known answer for the mixed-dimensional oracle (bulk mixed Poisson on both sides + fracture + Robin coupling,
    R/examples/md_elliptic_scalar/CouplingWeakForm.py:11-151): a uniform flux q through the fracture plane z = 1/2.  The
    bulk pressure is linear in z on each side with the jump the Robin law u . n = alpha (p|_Gamma - p_f) asks for
    (alpha = kappa_n / (delta / 2)), the fracture carries no tangential flux and a constant pressure, what leaves one side
    enters the other -- everything lies in RT_1 x P_0, so the discrete solution is exact: bulk flux (0, 0, -q) at every
    point, fracture pressure p_f, zero fracture flux."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl

    from oracle import ref_forms
    from oracle.ref_forms import RefElement

    mesh, spaces, ref_spaces, triples, outer = md_problem(2)
    q, kappa, kappa_n, delta, p_f = 0.8, 2.0, 3.0, 0.2, 0.4
    alpha = kappa_n / (delta / 2.0)
    side = lambda z: np.where(np.asarray(z) > 0.5, 1.0, -1.0)
    p_bulk = lambda x, y, z: np.array([p_f + side(z) * q / alpha + (q / kappa) * (np.asarray(z) - 0.5) + 0.0 * x])
    zero = lambda x, y, z: np.array([0.0 * x])
    f0 = {"rhs": zero, "kappa": lambda x, y, z: kappa}
    f1 = {"rhs": zero, "kappa": lambda x, y, z: 0.7}
    coupling = {"kappa_normal": lambda x, y, z: kappa_n, "delta": lambda x, y, z: delta}
    u0, u1 = ref_spaces[0].discrete_spaces["u"], ref_spaces[1].discrete_spaces["u"]
    indptr, indices, data, res = ref_md.assemble_md(
        ref_spaces,
        [[("form", ref_forms.laplace_dual, f0, len(u0.element_ids)), ("bc_form", ref_forms.laplace_dual_bc_dirichlet, {"p": p_bulk}, outer)],
         [("form", ref_forms.laplace_dual, f1, len(u1.element_ids)),
          ("bc_form", ref_forms.laplace_dual_bc_dirichlet, {"p": lambda x, y, z: np.array([p_f + 0.0 * x])}, len(u1.bc_element_ids))]],
        triples, coupling)
    A = sp.csr_matrix((data, indices, indptr)).tocsc()
    sol = spl.spsolve(A, -res)
    assert np.abs(A @ sol + res).max() < 1e-11
    pts = np.array([[0.2, 0.3, 0.1], [0.55, 0.15, 0.2]])
    s0, s1 = ref_spaces
    for i in range(len(u0.element_ids)):
        el = RefElement(u0, u0.element_ids[i])
        x, jac, det_jac, inv_jac = el.evaluate_mapping(pts)
        tab = el.evaluate_basis(pts, jac, det_jac, inv_jac)
        dest = s0.destination_indexes(i)
        for k in range(len(pts)):
            assert np.abs(sol[dest[: tab.shape[2]]] @ tab[0, k, :, :3] - np.array([0.0, 0.0, -q])).max() < 1e-9
    for i in range(len(u1.element_ids)):
        dest = s1.destination_indexes(i)
        assert abs(sol[dest[-1]] - p_f) < 1e-9 and np.abs(sol[dest[:-1]]).max() < 1e-9
