"""The vectorised "strong CPU" baseline of bench.py (oracle/strong_cpu.py: batched numpy + scipy COO -> CSR) produces
the same CSR matrix as the cell-by-cell oracle: identical pattern, values within 1e-12."""
import numpy as np

from oracle import ref_assembly, ref_forms, strong_cpu
from tests.parity_helpers import elasticity_case


def test_strong_cpu_baseline_matches_the_cell_by_cell_oracle():
    case = elasticity_case(n=2, with_bc=False, varying=False)
    rspace = case.ref_space()
    ref = case.run_oracle()
    u = rspace.discrete_spaces["u"]
    n_cells = len(u.element_ids)
    dest = np.array([rspace.destination_indexes(i) for i in range(n_cells)])
    cell_nodes = np.array([u.mesh.cells[u.element_ids[i]].node_tags for i in range(n_cells)])
    indptr, indices, data = strong_cpu.assemble_elasticity(u.mesh.points, cell_nodes, dest, rspace.n_dof, 3.0, 0.5)
    assert np.array_equal(indptr, ref[0]) and np.array_equal(indices, ref[1])
    assert np.abs(data - ref[2]).max() <= 1e-12 * np.abs(ref[2]).max()
