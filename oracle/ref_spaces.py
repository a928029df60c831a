"""ORACLE (test infrastructure only -- never imported by the product path).

Dictionary-based restatement of the reference DoF numbering, one cell at a time:

* ``RefDoFMap``   -- R/src/spaces/dof_map.py:38-170 (entity blocks, ``destination_indices``,
  ``bc_destination_indices``);
* ``RefSpace``    -- R/src/spaces/discrete_space.py:57-221 (element list, facet element list, trace element rule);
* ``RefProductSpace`` -- R/src/spaces/product_space.py:23-201 (quadrature choice, field strides, ``dof_shift``).

``num_entity_dofs`` comes from ``oracle.ref_basis`` (closed-form element layouts), not from the product's tables.
"""
import numpy as np

from . import ref_basis


class RefDoFMap:
    def __init__(self, mesh, dim, physical_tags, family, k_order, discontinuous, n_components):
        self.mesh, self.dim = mesh, dim
        self.selected, self.entity_ids = mesh.entity_tables(physical_tags, dim)
        self.num_entity_dofs = ref_basis.num_entity_dofs(family, dim, k_order, discontinuous)
        self.n_components = n_components
        support = [0, 0, 0, 0]
        self.e_dofs = [0, 0, 0, 0]
        for d, counts in enumerate(self.num_entity_dofs):
            self.e_dofs[d] = int(np.mean(counts))
            support[d] = len(self.entity_ids[d]) * self.e_dofs[d] * n_components
        offsets = np.add.accumulate([0] + support)
        self.maps = [dict(), dict(), dict(), dict()]
        for d in range(dim + 1):
            ids = self.entity_ids[d]
            if len(ids) == 0:
                continue
            chunks = np.split(np.arange(offsets[d], offsets[d + 1]), len(ids))
            self.maps[d] = dict(zip(ids, chunks))
        self.n_dof = int(sum(support))

    def destination_indices(self, cell_id):
        cell = self.mesh.cells[cell_id]
        out = []
        for d in range(self.dim + 1):
            out.append(np.array([self.maps[d].get(s) for s in cell.sub_cells_ids[d]], dtype=int).ravel())
        return np.concatenate(out)

    def bc_destination_indices(self, cell_id, bc_cell_id):
        cell = self.mesh.cells[cell_id]
        bc_sub = self.mesh.cells[bc_cell_id].sub_cells_ids
        out = []
        for d in range(self.dim):
            keep = [s for s in bc_sub[d] if s in set(cell.sub_cells_ids[d].tolist())]
            if self.e_dofs[d] == 0:
                keep = []
            out.append(np.array([self.maps[d].get(s) for s in keep], dtype=int).ravel())
        return np.concatenate(out)


class RefSpace:
    def __init__(self, dim, n_comp, family, k_order, mesh):
        self.dimension, self.n_comp, self.generator_family, self.k_order, self.mesh = dim, n_comp, family, k_order, mesh
        self.discontinuous = False
        self.bc_element_ids = []

    @property
    def family(self):
        if (self.dimension == 1 and self.generator_family in ("RT", "BDM", "N1E", "N2E")) or self.dimension == 0:
            return "Lagrange"
        return self.generator_family

    def build(self, physical_tags, b_physical_tags):
        self.dof_map = RefDoFMap(
            self.mesh, self.dimension, physical_tags, self.family, self.k_order, self.discontinuous, self.n_comp
        )
        self.n_dof = self.dof_map.n_dof
        self.element_ids = list(self.dof_map.entity_ids[self.dimension])
        self.id_to_element = {c: i for i, c in enumerate(self.element_ids)}
        self.bc_element_ids = []
        if len(b_physical_tags) != 0:
            facets = self.dof_map.entity_ids[self.dimension - 1]
            self.bc_element_ids = [f for f in facets if self.mesh.cells[f].material_id in b_physical_tags]
            disc, k, fam = self.discontinuous, self.k_order, self.family
            if self.dimension - 1 < 2:
                fam = "Lagrange"
                if self.dimension - 1 < 1:
                    k = 0
            if self.family == "BDM":
                k, fam, disc = self.k_order, "Lagrange", True
            if self.family == "RT":
                k, fam, disc = self.k_order - 1, "Lagrange", True
            self.bc_family, self.bc_k_order, self.bc_discontinuous = fam, k, disc

    def first_neighbour(self, facet_id):
        """First predecessor of the facet in the (cell -> facet) digraph: the first selected cell listing it."""
        for c in self.dof_map.selected:
            if facet_id in self.mesh.cells[c].sub_cells_ids[self.dimension - 1]:
                return c
        raise AssertionError("facet without neighbour")


class RefProductSpace:
    def __init__(self, data):
        self.names = list(data.keys())
        k_orders = [d[3] for d in data.values()]
        dims = [d[0] for d in data.values()]
        self.integration_order = 2 * max(k_orders) + 1
        self.quadrature = [None] * 4
        for dim in range(max(dims) + 1):
            self.quadrature[dim] = (
                (np.array([1.0]), np.array([1.0])) if dim == 0 else ref_basis.gauss_jacobi(dim, self.integration_order)
            )
        self.bc_quadrature = [None] * 4
        for dim in dims:
            dim -= 1
            if dim < 0:
                continue
            self.bc_quadrature[dim] = (
                (np.array([1.0]), np.array([1.0])) if dim == 0 else ref_basis.gauss_jacobi(dim, self.integration_order)
            )
        self.discrete_spaces = {n: RefSpace(d[0], d[1], d[2], d[3], d[4]) for n, d in data.items()}
        self.dof_shift = 0

    def make_subspaces_discontinuous(self, flags):
        for n in self.names:
            if flags.get(n, False):
                self.discrete_spaces[n].discontinuous = True

    def build_structures(self, physical_tags, b_physical_tags):
        for n, s in self.discrete_spaces.items():
            s.build(physical_tags.get(n, [None]), b_physical_tags.get(n, []))
        self.n_dof = sum(s.n_dof for s in self.discrete_spaces.values())

    def _strides(self):
        acc = np.add.accumulate([0] + [s.n_dof for s in self.discrete_spaces.values()])
        return dict(zip(self.names, acc[:-1]))

    def destination_indexes(self, cell_index, field_name=None):
        strides = self._strides()
        parts = {}
        for n, s in self.discrete_spaces.items():
            parts[n] = s.dof_map.destination_indices(s.element_ids[cell_index]) + strides[n] + self.dof_shift
        return np.concatenate(list(parts.values())) if field_name is None else parts[field_name]

    def bc_destination_indexes(self, cell_index, field_name=None):
        strides = self._strides()
        parts = {}
        for n, s in self.discrete_spaces.items():
            if len(s.bc_element_ids) == 0:
                continue
            facet = s.bc_element_ids[cell_index]
            neigh = s.first_neighbour(facet)
            parts[n] = s.dof_map.bc_destination_indices(neigh, facet) + strides[n] + self.dof_shift
        return np.concatenate(list(parts.values())) if field_name is None else parts[field_name]
