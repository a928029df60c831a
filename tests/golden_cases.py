"""Problem data of the committed golden fixtures (shared by tools/make_golden.py and the tests)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _le_rhs(x, y, z):
    return np.array([np.sin(np.pi * x) * np.cos(np.pi * y), np.cos(np.pi * y) * np.sin(np.pi * z), x * y + z])


def _le_u(x, y, z):
    return np.array([x * (1 - x) * y, np.sin(np.pi * x) * z, 0.5 * z * z + y])


LE_FUNCTIONS = {"rhs": _le_rhs, "lambda": lambda x, y, z: np.ones_like(x), "mu": lambda x, y, z: np.ones_like(x)}
LE_BC_FUNCTIONS = {"u": _le_u}
LAPLACE_FUNCTIONS = {
    "rhs": lambda x, y, z: np.array([2.0 * np.pi**2 * np.sin(np.pi * x) * np.sin(np.pi * y)]),
    "kappa": lambda x, y, z: 1.0,
}
LAPLACE_BC_FUNCTIONS = {"p": lambda x, y, z: np.array([np.sin(np.pi * x) * np.sin(np.pi * y)])}


def load_mesh_blocks(name):
    z = np.load(os.path.join(GOLDEN, name))
    blocks = [
        (str(z["type_%d" % i]), z["conn_%d" % i].astype(np.int64), z["phys_%d" % i].astype(np.int64))
        for i in range(int(z["n_blocks"]))
    ]
    return z["points"], blocks


def fixture_mesh(dim, level=0):
    """The reference's own gmsh fixture (example_3_{2d,3d}_l_{level}.msh) as a product mesh, rebuilt from the committed
    arrays."""
    from nfempy_b200.mesh import Mesh

    points, blocks = load_mesh_blocks("mesh_ex3_%dd_l%d.npz" % (dim, level))
    mesh = Mesh.from_arrays(dim, points, blocks)
    mesh.build_conformal_mesh()
    return mesh


def multi_cases():
    """name -> (case, tolerance) of the golden fixtures in multi_ex3_2d_l0.npz: forms added after the first two fixtures,
    all on the reference's 2-D gmsh mesh (116 triangles, boundary tags 2..5)."""
    from tests.parity_helpers import lce_primal_case, mixed_poisson_case, taylor_hood_case, tensor_mixed_case

    mesh = fixture_mesh(2)
    return {
        "lce_primal": (lce_primal_case(dim=2, k=2, kt=1, varying=True, alpha=True, mesh=mesh), 1e-12),
        "taylor_hood": (taylor_hood_case(varying=True, alpha=True, mesh=mesh), 1e-12),
        "mixed_poisson_bdm3": (mixed_poisson_case(dim=2, k=3, family="BDM", varying=True, mesh=mesh), 1e-12),
        "lce_dual_neumann": (tensor_mixed_case(kind="lce", dim=2, s_family="BDM", m_family="RT", k=1, mesh=mesh,
                                               neumann=True), 1e-12),
    }
