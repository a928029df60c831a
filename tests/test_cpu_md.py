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
