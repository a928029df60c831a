"""Array-based ``MeshTopology`` / ``DoFMap`` / ``DiscreteSpace`` with the reference's numbering.

Reference behaviour followed:

* selected cells and entity lists: R/src/mesh/mesh.py:668-681 + R/src/mesh/mesh_topology.py:22-47 -- cells of the
  field's dimension whose ``material_id`` is in the physical tags, ascending id; ``entity_ids[d]`` = sub-entities in
  order of first appearance (networkx node insertion order);
* global numbering: R/src/spaces/dof_map.py:38-129 -- blocks [vertices | edges | faces | volumes], each entity owns
  ``e_dofs * n_comp`` consecutive ids, ``e_dofs = int(mean(num_entity_dofs[d]))``;
* local -> global: R/src/spaces/dof_map.py:134-146 (cell) and :148-170 (facet through its first neighbour);
* facet ("bc") elements: R/src/spaces/discrete_space.py:163-221.

The reference answers ``destination_indices`` with networkx + dict look-ups per call; here the whole table
``int32[n_cells, n_loc]`` is built once with numpy and is what the CUDA kernels consume.
"""
import numpy as np

from ..basis.reference_element import create_element
from ..mesh import _accel
from ..basis.finite_element import FiniteElementView

_VECTOR_FAMILIES = ("RT", "BDM", "N1E", "N2E")


class MeshTopology:
    def __init__(self, mesh, dimension):
        if mesh.dimension < dimension:
            raise ValueError("MeshTopology:: max dimension available in the mesh is %r" % mesh.dimension)
        self.mesh = mesh
        self.dimension = dimension
        self.entity_ids = {}
        self.selected = np.zeros(0, dtype=np.int64)

    def build_data_on_pysical_tags(self, physical_tags=()):
        mesh, dim = self.mesh, self.dimension
        ids = mesh.ids_of_dimension(dim)
        tags = [t for t in physical_tags if t is not None]
        self.selected = ids[np.isin(mesh.material[ids], np.array(tags, dtype=np.int64))]
        self.selected_local = self.selected - mesh.offset[dim]
        self.entity_ids = {}
        self.entity_pos = {}
        for d in range(dim + 1):
            sub = mesh.sub_cells_of(dim, d)[self.selected_local].ravel()
            uniq, first = _accel.unique(sub, return_index=True)
            order = _accel.argsort(first)
            self.entity_ids[d] = uniq[order]
            pos = np.full(int(mesh.offset[d + 1] - mesh.offset[d]), -1, dtype=np.int64)
            pos[self.entity_ids[d] - mesh.offset[d]] = np.arange(len(order))
            self.entity_pos[d] = pos

    def entities_by_dimension(self, dimension):
        if self.dimension < dimension:
            raise ValueError("MeshTopology:: max dimension available is %r" % self.dimension)
        return self.entity_ids[dimension]

    def entities_by_codimension(self, codimension):
        return self.entity_ids[self.dimension - codimension]

    def first_neighbours(self, facet_ids):
        """For each facet (dimension ``dim-1`` cell id) the first selected cell containing it and the facet's
        local index there -- ``list(entity_map.predecessors(facet))[0]`` in the reference."""
        mesh, dim = self.mesh, self.dimension
        sub = mesh.sub_cells_of(dim, dim - 1)[self.selected_local]
        flat = sub.ravel()
        uniq, first = _accel.unique(flat, return_index=True)
        where = np.searchsorted(uniq, facet_ids)
        if np.any(where >= len(uniq)) or not np.array_equal(uniq[np.minimum(where, len(uniq) - 1)], facet_ids):
            raise ValueError("MeshTopology: facet without a neighbour cell in the selected sub-domain")
        hit = first[where]
        n_sub = sub.shape[1]
        return hit // n_sub, hit % n_sub


class DoFMap:
    def __init__(self, mesh_topology, ref_element, n_components):
        self.mesh_topology = mesh_topology
        self.ref_element = ref_element
        self.dimension = mesh_topology.dimension
        self.n_components = n_components
        self.n_dof = 0
        self.build_entity_maps(n_components)

    def build_entity_maps(self, n_components=1, n_dof_shift=0):
        topo, dim = self.mesh_topology, self.dimension
        self.e_dofs = np.zeros(4, dtype=np.int64)
        support = np.zeros(4, dtype=np.int64)
        for d, counts in enumerate(self.ref_element.num_entity_dofs):
            self.e_dofs[d] = int(np.mean(counts))
            support[d] = len(topo.entity_ids[d]) * self.e_dofs[d] * n_components
        self.global_indices = np.concatenate([[0], np.cumsum(support)]) + n_dof_shift
        self.n_dof = int(support.sum())

    def dof_number(self):
        return self.n_dof

    def entity_block(self, d, entity_ids):
        """Global dofs ``[n, e_dofs*ncomp]`` owned by entities of dimension ``d`` (global cell ids)."""
        width = int(self.e_dofs[d] * self.n_components)
        pos = self.mesh_topology.entity_pos[d][np.asarray(entity_ids) - self.mesh_topology.mesh.offset[d]]
        if np.any(pos < 0):
            raise ValueError("DoFMap: entity outside the field's sub-domain")
        return self.global_indices[d] + pos[..., None] * width + np.arange(width)

    def destination_table(self):
        """``int64[n_selected, n_loc_field]`` -- ``destination_indices`` of every selected cell."""
        topo, dim = self.mesh_topology, self.dimension
        parts = []
        for d in range(dim + 1):
            if self.e_dofs[d] == 0:
                continue
            sub = topo.mesh.sub_cells_of(dim, d)[topo.selected_local]
            width = sub.shape[1] * int(self.e_dofs[d] * self.n_components)
            parts.append(self.entity_block(d, sub).reshape(len(sub), width))
        if not parts:
            return np.zeros((len(topo.selected), 0), dtype=np.int64)
        return np.concatenate(parts, axis=1)

    def destination_indices(self, cell_id):
        topo = self.mesh_topology
        row = np.searchsorted(topo.selected, cell_id)
        if row >= len(topo.selected) or topo.selected[row] != cell_id:
            raise KeyError(cell_id)
        return self._dest_cache()[row]

    def _dest_cache(self):
        if not hasattr(self, "_dest"):
            self._dest = self.destination_table()
        return self._dest

    def bc_destination_table(self, bc_cell_ids):
        """``bc_destination_indices(first neighbour, facet)`` for every facet id in ``bc_cell_ids``.

        The reference filters the facet's own sub-entities by membership in the neighbour (always true for a
        conformal mesh) and keeps the facet's sub-entity order (R/src/spaces/dof_map.py:148-170)."""
        topo, dim = self.mesh_topology, self.dimension
        mesh = topo.mesh
        local = np.asarray(bc_cell_ids) - mesh.offset[dim - 1]
        parts = []
        for d in range(dim):
            if self.e_dofs[d] == 0:
                continue
            sub = mesh.sub_cells_of(dim - 1, d)[local]
            width = sub.shape[1] * int(self.e_dofs[d] * self.n_components)
            parts.append(self.entity_block(d, sub).reshape(len(sub), width))
        if not parts:
            return np.zeros((len(local), 0), dtype=np.int64)
        return np.concatenate(parts, axis=1)


class _ElementList:
    """Lazy stand-in for the reference's ``list[FiniteElement]`` (one Python object per cell there)."""

    def __init__(self, space, cell_ids, ref_element, bc):
        self._space, self._ids, self._ref, self._bc = space, cell_ids, ref_element, bc

    def __len__(self):
        return len(self._ids)

    def __getitem__(self, i):
        if i < 0:
            i += len(self._ids)
        if i < 0 or i >= len(self._ids):
            raise IndexError(i)
        return FiniteElementView(self._space, int(self._ids[i]), self._ref, self._bc)

    def __iter__(self):
        for i in range(len(self._ids)):
            yield self[i]


class DiscreteSpace:
    def __init__(self, dimension, n_components, family, k_order, mesh, integration_order=0):
        self.dimension = dimension
        self.n_comp = n_components
        self.generator_family = family
        self.k_order = k_order
        self.integration_order = integration_order
        self.name = "Unnamed"
        self.mesh_topology = MeshTopology(mesh, dimension)
        self.discontinuous = False
        self.element_ids = np.zeros(0, dtype=np.int64)
        self.bc_element_ids = np.zeros(0, dtype=np.int64)
        self.elements = []
        self.bc_elements = []
        self.id_to_element = {}
        self.id_to_bc_element = {}
        self.ref_element = None
        self.bc_ref_element = None

    @property
    def family(self):
        # R/src/spaces/discrete_space.py:39-49
        if (self.dimension == 1 and self.generator_family in _VECTOR_FAMILIES) or self.dimension == 0:
            return "Lagrange"
        return self.generator_family

    @property
    def mesh(self):
        return self.mesh_topology.mesh

    def set_name(self, name):
        self.name = name

    def make_discontinuous(self):
        self.discontinuous = True

    def build_structures(self, physical_tags=(None,)):
        self.mesh_topology.build_data_on_pysical_tags(physical_tags)
        self.ref_element = create_element(self.family, self.dimension, self.k_order, self.discontinuous)
        self.dof_map = DoFMap(self.mesh_topology, self.ref_element, self.n_comp)
        self.n_dof = self.dof_map.dof_number()
        self.element_ids = self.mesh_topology.entities_by_dimension(self.dimension)
        self.elements = _ElementList(self, self.element_ids, self.ref_element, bc=False)
        self.id_to_element = _IdToIndex(self.element_ids)

    def build_boundary_structures(self, bc_physical_tags=(None,)):
        if len(bc_physical_tags) == 0:
            return
        mesh = self.mesh
        facets = self.mesh_topology.entities_by_dimension(self.dimension - 1)
        tags = np.array([t for t in bc_physical_tags if t is not None], dtype=np.int64)
        self.bc_element_ids = facets[np.isin(mesh.material[facets], tags)]
        # trace element (R/src/spaces/discrete_space.py:177-194)
        disc, k, fam = self.discontinuous, self.k_order, self.family
        if self.dimension - 1 < 2:
            fam = "Lagrange"
            if self.dimension - 1 < 1:
                k = 0
        if self.family == "BDM":
            k, fam, disc = self.k_order, "Lagrange", True
        if self.family == "RT":
            k, fam, disc = self.k_order - 1, "Lagrange", True
        self.bc_ref_element = create_element(fam, self.dimension - 1, k, disc) if self.dimension >= 1 else None
        self.bc_elements = _ElementList(self, self.bc_element_ids, self.bc_ref_element, bc=True)
        self.id_to_bc_element = _IdToIndex(self.bc_element_ids)

    # -- arrays for the kernels -----------------------------------------------------------------
    def cell_nodes(self):
        """``int64[n_elements, dim+1]`` node tags of the field's cells (vertex order = ``node_tags``)."""
        return self.mesh.node_tags_of(self.dimension)[self.mesh_topology.selected_local]

    def bc_cell_nodes(self):
        return self.mesh.node_tags_of(self.dimension - 1)[self.bc_element_ids - self.mesh.offset[self.dimension - 1]]

    def destination_table(self):
        """``int64[n_elements, n_loc_field]``.  Continuous cubic Lagrange fields: the two interior dofs of an edge the
        reference finds mis-oriented are listed in exchanged order.  The reference exchanges the two basis functions instead
        (``permute_and_transform`` with ``transformations["interval"][0] = [[0, 1], [1, 0]]``,
        R/src/basis/permute_and_transform.py:22-35, 110-140: P_3 is the first Lagrange degree whose dof transformations
        are not the identity); scattering the untransformed local matrix through the exchanged destinations gives the same
        global system, and the kernels keep one reference table for every cell."""
        if not hasattr(self, "_dest_oriented"):
            dest = self.dof_map._dest_cache()
            if self.family == "Lagrange" and self.k_order == 3 and not self.discontinuous and self.dimension >= 2:
                from ..basis.orientation import edge_signs

                signs = edge_signs(self.mesh, self.dimension, self.mesh_topology.selected_local)  # [n, n_edges]
                nc, first = int(self.n_comp), (self.dimension + 1) * int(self.n_comp)
                dest = dest.copy()
                for e in range(signs.shape[1]):
                    rows = np.flatnonzero(signs[:, e] < 0)
                    a = first + 2 * nc * e
                    block = dest[rows, a : a + 2 * nc].copy()
                    dest[rows, a : a + nc], dest[rows, a + nc : a + 2 * nc] = block[:, nc:], block[:, :nc]
            self._dest_oriented = dest
        return self._dest_oriented

    def bc_destination_table(self):
        """``int64[n_bc_elements, n_loc_facet]``.  Continuous cubic Lagrange fields on tetrahedra: the boundary triangle's
        own nodal table lists the two interior nodes of its local edge (a, b) from a to b, the volume side orders the
        edge's two dofs along ascending vertex tags (``destination_table``), so the pair is exchanged where
        ``tag[a] > tag[b]``.  (On triangles the facet IS the edge and both sides follow the edge cell's own vertex order.)"""
        if not hasattr(self, "_bc_dest"):
            dest = self.dof_map.bc_destination_table(self.bc_element_ids)
            if self.family == "Lagrange" and self.k_order == 3 and not self.discontinuous and self.dimension == 3 and len(dest):
                from ..basis.reference_element import SUB_ENTITIES

                tags = self.bc_cell_nodes()
                nc = int(self.n_comp)
                dest = dest.copy()
                for e, (a, b) in enumerate(SUB_ENTITIES[2][1]):
                    rows = np.flatnonzero(tags[:, a] > tags[:, b])
                    c0 = (3 + 2 * e) * nc
                    block = dest[rows, c0 : c0 + 2 * nc].copy()
                    dest[rows, c0 : c0 + nc], dest[rows, c0 + nc : c0 + 2 * nc] = block[:, nc:], block[:, :nc]
            self._bc_dest = dest
        return self._bc_dest


class _IdToIndex:
    """``dict(zip(ids, range(n)))`` without materialising a dict for 10^6 cells."""

    def __init__(self, ids):
        self._order = np.argsort(ids, kind="stable")
        self._sorted = np.asarray(ids)[self._order]

    def __getitem__(self, cell_id):
        p = np.searchsorted(self._sorted, cell_id)
        if p >= len(self._sorted) or self._sorted[p] != cell_id:
            raise KeyError(cell_id)
        return int(self._order[p])

    def get(self, cell_id, default=None):
        try:
            return self[cell_id]
        except KeyError:
            return default

    def lookup(self, cell_ids):
        p = np.searchsorted(self._sorted, cell_ids)
        return self._order[p]

    def __len__(self):
        return len(self._sorted)
