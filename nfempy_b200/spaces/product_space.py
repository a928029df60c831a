"""``ProductSpace`` with the reference's constructor, attributes and index semantics
(R/src/spaces/product_space.py:9-201), backed by whole-mesh destination tables.

``discrete_spaces_data = {field: (dim, n_components, family, k_order, mesh)}``.
"""
import numpy as np

from ..basis.quadrature import make_quadrature
from .discrete_space import DiscreteSpace


class ProductSpace:
    def __init__(self, discrete_spaces_data):
        self.names = list(discrete_spaces_data.keys())
        self._define_integration_order_and_quadrature(discrete_spaces_data)
        self.discrete_spaces = {}
        for field, (dim, components, family, k_order, gmesh) in discrete_spaces_data.items():
            space = DiscreteSpace(dim, components, family, k_order, gmesh, integration_order=self.integration_order)
            space.name = field
            self.discrete_spaces[field] = space
        self._dof_shift = 0
        self.n_dof = 0
        self.discrete_spaces_dofs = {}
        self.fields_physical_tags = None

    def _define_integration_order_and_quadrature(self, data):
        # R/src/spaces/product_space.py:23-50: one rule of degree 2 max(k) + 1 per dimension, shared by all fields
        k_orders = [d[3] for d in data.values()]
        dims = [d[0] for d in data.values()]
        self.integration_order = 2 * int(np.max(k_orders)) + 1
        self.quadrature = [None, None, None, None]
        for dim in range(int(np.max(dims)) + 1):
            if dim == 0:
                self.quadrature[dim] = (np.array([1.0]), np.array([1.0]))
            else:
                self.quadrature[dim] = make_quadrature(dim, self.integration_order)
        self.bc_quadrature = [None, None, None, None]
        for dim in dims:
            dim -= 1
            if dim < 0:
                continue
            if dim == 0:
                self.bc_quadrature[dim] = (np.array([1.0]), np.array([1.0]))
            else:
                self.bc_quadrature[dim] = make_quadrature(dim, self.integration_order)

    @property
    def dof_shift(self):
        return self._dof_shift

    @dof_shift.setter
    def dof_shift(self, value):
        if not isinstance(value, (int, np.integer)):
            raise TypeError("dof_shift must be an integer")
        if value < 0:
            raise ValueError("dof_shift must be non-negative")
        self._dof_shift = int(value)
        self._drop_tables()

    def dimension(self):
        return int(np.max([s.dimension for s in self.discrete_spaces.values()]))

    def make_subspaces_discontinuous(self, discrete_space_disc):
        for name in self.names:
            if discrete_space_disc.get(name, False):
                self.discrete_spaces[name].make_discontinuous()

    def build_structures(self, physical_tags, b_physical_tags):
        self.fields_physical_tags = physical_tags
        for name, space in self.discrete_spaces.items():
            space.build_structures(physical_tags.get(name, [None]))
            space.build_boundary_structures(b_physical_tags.get(name, []))
        self.discrete_spaces_dofs = {n: s.dof_map.dof_number() for n, s in self.discrete_spaces.items()}
        self.n_dof = int(sum(self.discrete_spaces_dofs.values()))
        self._drop_tables()

    # -- strides --------------------------------------------------------------------------------
    def field_strides(self):
        dofs = [0] + list(self.discrete_spaces_dofs.values())
        return dict(zip(self.names, np.add.accumulate(dofs)[:-1]))

    def _drop_tables(self):
        self._dest = None
        self._bc_dest = None

    # -- whole-mesh tables (what the kernels and the sparsity plan read) -------------------------
    def destination_table(self, field_name=None):
        """``int64[n_elements, n_loc]``: row ``i`` equals ``destination_indexes(i)``."""
        strides = self.field_strides()
        if field_name is not None:
            return self.discrete_spaces[field_name].destination_table() + strides[field_name] + self.dof_shift
        if self._dest is None:
            parts = [s.destination_table() + strides[n] + self.dof_shift for n, s in self.discrete_spaces.items()]
            n_el = {len(p) for p in parts}
            if len(n_el) != 1:
                raise ValueError("ProductSpace: fields are defined on different element lists")
            self._dest = np.concatenate(parts, axis=1)
        return self._dest

    def bc_destination_table(self, field_name=None):
        strides = self.field_strides()
        if field_name is not None:
            space = self.discrete_spaces[field_name]
            return space.bc_destination_table() + strides[field_name] + self.dof_shift
        if self._bc_dest is None:
            parts = [
                s.bc_destination_table() + strides[n] + self.dof_shift
                for n, s in self.discrete_spaces.items()
                if len(s.bc_elements) != 0
            ]
            self._bc_dest = np.concatenate(parts, axis=1) if parts else np.zeros((0, 0), dtype=np.int64)
        return self._bc_dest

    # -- reference per-element API ----------------------------------------------------------------
    def discrete_spaces_destination_indexes(self, cell_index):
        strides = self.field_strides()
        return {
            n: s.destination_table()[cell_index] + strides[n] + self.dof_shift
            for n, s in self.discrete_spaces.items()
        }

    def destination_indexes(self, cell_index, field_name=None):
        if field_name is None:
            return self.destination_table()[cell_index].copy()
        return self.discrete_spaces_destination_indexes(cell_index)[field_name]

    def discrete_spaces_bc_destination_indexes(self, cell_index):
        strides = self.field_strides()
        return {
            n: s.bc_destination_table()[cell_index] + strides[n] + self.dof_shift
            for n, s in self.discrete_spaces.items()
            if len(s.bc_elements) != 0
        }

    def bc_destination_indexes(self, cell_index, field_name=None):
        if field_name is None:
            return np.concatenate(list(self.discrete_spaces_bc_destination_indexes(cell_index).values()))
        return self.discrete_spaces_bc_destination_indexes(cell_index)[field_name]
