"""Generates the committed fixtures under tests/golden/ (run in the build container, where /root/reference exists).

* mesh_ex3_{3d,2d}_l0.npz      -- points + element blocks of the reference's own gmsh fixtures
  (R/examples/cosserat_elasticity/gmsh_files/ex_3/example_3_{3d,2d}_l_0.msh) as read by
  nfempy_b200.mesh.msh_reader (restating meshio's gmsh-4.1 reader), so the GPU box can rebuild the meshes
  without /root/reference;
* numbering_ex3_3d_l0.npz       -- cell numbering (node_tags / sub-cell ids / material) from oracle.ref_mesh;
* le_p2_ex3_3d_l0.npz           -- oracle CSR + rhs: LEPrimalWeakForm P2 x 3 + Dirichlet penalty on the six sides;
* laplace_p1_ex3_2d_l0.npz      -- oracle CSR + rhs: LaplacePrimalWeakForm P1 + Dirichlet penalty (config 1 style);
* multi_ex3_2d_l0.npz           -- oracle CSR + rhs of four later form families on the 2-D fixture (primal Cosserat + BC,
  Taylor-Hood + BC, mixed Poisson on BDM_3 + natural BC, mixed Cosserat with natural and penalised stress BCs);
  ``python tools/make_golden.py multi`` regenerates only this file (needs no /root/reference: the mesh arrays are committed).

* mesh_ex3_3d_l{1,2}.npz, le_p2_ex3_3d_l1.npz, le_p2_ex3_3d_l2_probe.npz -- the larger levels of the 3-D fixture
  (``python tools/make_golden.py levels``), see ``main_levels``.

The reference itself cannot run here (basix / auto_diff / petsc4py / meshio absent), so these are ORACLE outputs:
they pin the oracle against regressions and give the GPU tests a fixed target; they do not pin the oracle to the
real reference ("parity unpinned").  Later: tools/make_reference_pins.py --verify-golden runs the reference's own
source on the same gmsh files and form lists and reproduces these arrays bit for bit (tests/golden/refcode_verified.json).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from nfempy_b200.mesh.msh_reader import read_msh  # noqa: E402
from oracle import ref_assembly, ref_forms  # noqa: E402
from oracle.ref_mesh import RefMesh  # noqa: E402
from oracle.ref_spaces import RefProductSpace  # noqa: E402
from tests.golden_cases import LE_FUNCTIONS, LE_BC_FUNCTIONS, LAPLACE_FUNCTIONS, LAPLACE_BC_FUNCTIONS  # noqa: E402

REF = "/root/reference/examples/cosserat_elasticity/gmsh_files/ex_3"
OUT = os.path.join(ROOT, "tests", "golden")


def save_mesh(name, path):
    points, blocks = read_msh(path)
    arrays = {"points": points, "n_blocks": np.array(len(blocks))}
    for i, (t, c, p) in enumerate(blocks):
        arrays["type_%d" % i] = np.array(t)
        arrays["conn_%d" % i] = c.astype(np.int32)
        arrays["phys_%d" % i] = p.astype(np.int32)
    np.savez_compressed(os.path.join(OUT, name), **arrays)
    return points, blocks


def main():
    os.makedirs(OUT, exist_ok=True)
    points, blocks = save_mesh("mesh_ex3_3d_l0.npz", os.path.join(REF, "example_3_3d_l_0.msh"))
    mesh = RefMesh(3, points, blocks)
    mesh.build_conformal_mesh()
    dims = np.array([c.dimension for c in mesh.cells])
    mats = np.array([-1 if c.material_id is None else c.material_id for c in mesh.cells])
    nodes = np.full((len(mesh.cells), 4), -1)
    for c in mesh.cells:
        nodes[c.id, : len(c.node_tags)] = c.node_tags
    vol = [c for c in mesh.cells if c.dimension == 3]
    np.savez_compressed(
        os.path.join(OUT, "numbering_ex3_3d_l0.npz"), dims=dims.astype(np.int8), material=mats.astype(np.int32),
        node_tags=nodes.astype(np.int32), vol_edges=np.array([c.sub_cells_ids[1] for c in vol], dtype=np.int32),
        vol_faces=np.array([c.sub_cells_ids[2] for c in vol], dtype=np.int32),
    )
    space = RefProductSpace({"u": (3, 3, "Lagrange", 2, mesh)})
    space.build_structures({"u": [1]}, {"u": [2, 3, 4, 5, 6, 7]})
    us = space.discrete_spaces["u"]
    forms = [
        ("form", ref_forms.le_primal, LE_FUNCTIONS, len(us.element_ids)),
        ("bc_form", ref_forms.le_primal_bc_dirichlet, LE_BC_FUNCTIONS, len(us.bc_element_ids)),
    ]
    indptr, indices, data, rhs = ref_assembly.assemble(space, forms)
    np.savez_compressed(os.path.join(OUT, "le_p2_ex3_3d_l0.npz"), indptr=indptr, indices=indices, data=data, rhs=rhs)
    print("le_p2_ex3_3d_l0: n_dof", space.n_dof, "nnz", len(indices))

    points, blocks = save_mesh("mesh_ex3_2d_l0.npz", os.path.join(REF, "example_3_2d_l_0.msh"))
    mesh = RefMesh(2, points, blocks)
    mesh.build_conformal_mesh()
    btags = sorted({int(p[0]) for t, c, p in blocks if t == "line"})
    dtag = sorted({int(p[0]) for t, c, p in blocks if t == "triangle"})
    space = RefProductSpace({"p": (2, 1, "Lagrange", 1, mesh)})
    space.build_structures({"p": dtag}, {"p": btags})
    ps = space.discrete_spaces["p"]
    forms = [
        ("form", ref_forms.laplace_primal, LAPLACE_FUNCTIONS, len(ps.element_ids)),
        ("bc_form", ref_forms.laplace_primal_bc_dirichlet, LAPLACE_BC_FUNCTIONS, len(ps.bc_element_ids)),
    ]
    indptr, indices, data, rhs = ref_assembly.assemble(space, forms)
    np.savez_compressed(
        os.path.join(OUT, "laplace_p1_ex3_2d_l0.npz"), indptr=indptr, indices=indices, data=data, rhs=rhs,
        tags=np.array(dtag), btags=np.array(btags),
    )
    print("laplace_p1_ex3_2d_l0: n_dof", space.n_dof, "nnz", len(indices))


def main_levels():
    """The larger levels of the reference's 3-D fixture: meshes + oracle results of
    vector-P2 elasticity with the Dirichlet penalty on the six sides.
    * level 1 (1 325 tets): the complete CSR + rhs of the cell-by-cell oracle (le_p2_ex3_3d_l1.npz);
    * level 2 (4 261 tets): a compact fingerprint of the cell-by-cell oracle's result (le_p2_ex3_3d_l2_probe.npz) -- the
      int32 pattern, the products K v for four seeded vectors, the diagonal, the row sums of |K| and the rhs; the full
      value array (12 MB) is not committed, the GPU test compares all values against oracle/strong_cpu.py at run time."""
    for level, full in ((1, True), (2, False)):
        points, blocks = save_mesh("mesh_ex3_3d_l%d.npz" % level, os.path.join(REF, "example_3_3d_l_%d.msh" % level))
        mesh = RefMesh(3, points, blocks)
        mesh.build_conformal_mesh()
        space = RefProductSpace({"u": (3, 3, "Lagrange", 2, mesh)})
        space.build_structures({"u": [1]}, {"u": [2, 3, 4, 5, 6, 7]})
        us = space.discrete_spaces["u"]
        forms = [
            ("form", ref_forms.le_primal, LE_FUNCTIONS, len(us.element_ids)),
            ("bc_form", ref_forms.le_primal_bc_dirichlet, LE_BC_FUNCTIONS, len(us.bc_element_ids)),
        ]
        indptr, indices, data, rhs = ref_assembly.assemble(space, forms)
        print("le_p2_ex3_3d_l%d: n_dof" % level, space.n_dof, "nnz", len(indices))
        if full:
            np.savez_compressed(os.path.join(OUT, "le_p2_ex3_3d_l%d.npz" % level), indptr=indptr, indices=indices, data=data, rhs=rhs)
        else:
            import scipy.sparse as sp

            n = space.n_dof
            K = sp.csr_matrix((data, indices, indptr), shape=(n, n))
            V = np.random.default_rng(2024).standard_normal((n, 4))
            np.savez_compressed(
                os.path.join(OUT, "le_p2_ex3_3d_l%d_probe.npz" % level), indptr=indptr, indices=indices, rhs=rhs,
                KV=K @ V, diag=K.diagonal(), abs_row_sum=np.asarray(abs(K).sum(axis=1)).ravel(), seed=np.array(2024),
            )


def main_multi():
    """multi_ex3_2d_l0.npz: oracle CSR + rhs of the later form families on the 2-D fixture mesh (tests/golden_cases.py)."""
    from tests.golden_cases import multi_cases

    arrays = {}
    for name, (case, _) in multi_cases().items():
        indptr, indices, data, rhs = case.run_oracle()
        arrays.update({name + "_indptr": indptr, name + "_indices": indices, name + "_data": data, name + "_rhs": rhs})
        print(name, "n_dof", len(rhs), "nnz", len(indices))
    np.savez_compressed(os.path.join(OUT, "multi_ex3_2d_l0.npz"), **arrays)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "multi":
        main_multi()
    elif len(sys.argv) > 1 and sys.argv[1] == "levels":
        main_levels()
    else:
        main()
        main_multi()
