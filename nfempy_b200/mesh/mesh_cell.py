"""Light-weight cell record with the attribute surface of the reference ``MeshCell``
(R/src/mesh/mesh_cell.py:53-121): ``dimension``, ``id``, ``material_id``, ``node_tags``, ``perm``,
``sub_cells_ids[d]`` and the small query helpers the weak forms and spaces use."""
import numpy as np


class MeshCell:
    __slots__ = ("dimension", "type", "id", "material_id", "physical_name", "node_tags", "perm", "sub_cells_ids")

    def __init__(self, dimension):
        self.dimension = dimension
        self.type = "%dd-cell" % dimension
        self.id = None
        self.material_id = None
        self.physical_name = "Unnamed"
        self.node_tags = np.zeros(0, dtype=int)
        self.perm = np.zeros(0, dtype=int)
        self.sub_cells_ids = [np.zeros(0, dtype=int) for _ in range(dimension + 1)]

    def set_id(self, id):
        self.id = id

    def get_id(self):
        return self.id

    def set_material_id(self, material_id):
        self.material_id = material_id

    def get_material_id(self):
        return self.material_id

    def set_node_tags(self, node_tags):
        self.node_tags = node_tags
        self.perm = np.argsort(node_tags)

    def get_node_tags(self):
        return self.node_tags

    def set_sub_cells_ids(self, dim, cells_ids):
        if dim > self.dimension:
            raise ValueError("sub-cell dimension exceeds cell dimension")
        self.sub_cells_ids[dim] = cells_ids

    def sub_cell_index(self, dim, target_id):
        hit = np.flatnonzero(self.sub_cells_ids[dim] == target_id)
        return int(hit[0]) if hit.size else None

    def node_tag_index(self, target_tag):
        hit = np.flatnonzero(self.node_tags == target_tag)
        return int(hit[0]) if hit.size else None

    def index(self):
        return (self.dimension, self.id)
