"""Stand-in for meshio (absent here): ``read`` answers with nfempy_b200.mesh.msh_reader.read_msh's arrays behind the surface
R/src/mesh/mesh.py:19, 247-250, 334-337 uses.  ``register(name, points, blocks)`` makes in-memory meshes readable by name."""
import numpy as np

_DIM = {"vertex": 0, "line": 1, "triangle": 2, "tetra": 3}
_MEMORY = {}


class CellBlock:
    def __init__(self, type_, data):
        self.type, self.data, self.dim = type_, np.asarray(data), _DIM[type_]


class Mesh:
    def __init__(self, points, blocks):
        self.points = np.asarray(points, dtype=float)
        self.cells = [CellBlock(t, c) for t, c, _ in blocks]
        self.cell_data = {"gmsh:physical": [np.asarray(p) for _, _, p in blocks]}

    def get_cells_type(self, name):
        sel = [b.data for b in self.cells if b.type == name]
        return np.concatenate(sel) if sel else np.empty((0, _DIM[name] + 1), dtype=int)


def register(name, points, blocks):
    _MEMORY[name] = (points, blocks)


def read(file_name):
    if file_name in _MEMORY:
        return Mesh(*_MEMORY[file_name])
    from nfempy_b200.mesh.msh_reader import read_msh

    return Mesh(*read_msh(file_name))
