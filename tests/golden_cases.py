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
