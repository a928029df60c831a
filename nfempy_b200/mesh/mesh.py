"""Array-based restatement of the reference ``Mesh`` entity numbering.

Reference behaviour followed (R = /root/reference):

* global cell ids (R/src/mesh/mesh.py:239-339): vertices ``0..nv-1`` in point order; edges numbered by
  first appearance scanning tetrahedra, then triangles, then lines (tet local edges (2,3),(1,3),(1,2),(0,3),
  (0,2),(0,1); triangle local edges (1,2),(0,2),(0,1)); faces by first appearance over tetrahedra (local faces
  (1,2,3),(0,2,3),(0,1,3),(0,1,2)) then triangles; volumes last;
* cell materialisation in file block order (R/src/mesh/mesh.py:125-237, 334-337): a sub-entity's ``node_tags``
  keep the order of whichever element created it first -- a tet hands its edges/faces over in local order,
  a triangle hands its edges over sorted, a file element keeps the order in the file; only file elements
  carry a ``material_id``.

The reference keeps one Python object per cell; 10^6 tets have ~4.5*10^6 cells, so everything here is flat
numpy arrays and ``mesh.cells[id]`` builds a ``MeshCell`` view on demand.
"""
import numpy as np

from . import _accel
from .mesh_cell import MeshCell
from .msh_reader import read_msh

TET_EDGES = np.array([[2, 3], [1, 3], [1, 2], [0, 3], [0, 2], [0, 1]])
TET_FACES = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])
TRI_EDGES = np.array([[1, 2], [0, 2], [0, 1]])
_TYPE_DIM = {"vertex": 0, "line": 1, "triangle": 2, "tetra": 3}
NO_MATERIAL = -1


def _first_appearance(keys):
    """Rank of each key by order of first appearance, plus the number of distinct keys."""
    uniq, first = _accel.unique(keys, return_index=True)
    order = _accel.argsort(first)
    return None, uniq[order]


class _SortedLookup:
    """key -> id for a set of distinct int64 keys."""

    def __init__(self, keys, ids):
        o = _accel.argsort(keys)
        self.keys = keys[o]
        self.ids = ids[o]

    def __call__(self, keys):
        pos = _accel.searchsorted(self.keys, keys)
        pos = np.minimum(pos, max(len(self.keys) - 1, 0))
        if len(self.keys) == 0 or not np.array_equal(self.keys[pos], keys):
            raise KeyError("unknown mesh entity")
        return self.ids[pos]


class _CellsView:
    def __init__(self, mesh):
        self._mesh = mesh

    def __len__(self):
        return self._mesh.n_cells

    @property
    def shape(self):
        return (self._mesh.n_cells,)

    def __getitem__(self, cell_id):
        if isinstance(cell_id, (slice, list, np.ndarray)):
            ids = np.arange(self._mesh.n_cells)[cell_id]
            return [self._mesh._make_cell(int(i)) for i in np.atleast_1d(ids)]
        return self._mesh._make_cell(int(cell_id))

    def __iter__(self):
        for i in range(self._mesh.n_cells):
            yield self._mesh._make_cell(i)


class Mesh:
    def __init__(self, dimension, file_name=None):
        self.dimension = dimension
        self.points = None
        self._blocks = None
        self.n_cells = 0
        if file_name is not None:
            self.points, self._blocks = read_msh(file_name)

    @classmethod
    def from_arrays(cls, dimension, points, blocks):
        """``blocks``: list of ``(type, connectivity[n,k], physical[n])`` in (would-be) file order."""
        mesh = cls(dimension)
        mesh.points = np.ascontiguousarray(points, dtype=np.float64)
        mesh._blocks = [
            (t, np.asarray(c, dtype=np.int64), np.asarray(p, dtype=np.int64)) for t, c, p in blocks
        ]
        return mesh

    # -- reference API -------------------------------------------------------------------------
    def get_cells_type(self, name):
        k = _TYPE_DIM[name] + 1
        sel = [c for t, c, _ in self._blocks if t == name]
        if len(sel) == 0:
            return np.zeros((0, k), dtype=np.int64)
        return np.concatenate(sel, axis=0)

    def max_node_tag(self):
        return self.points.shape[0]

    def max_cell_id(self):
        return self.n_cells

    @property
    def cells(self):
        return _CellsView(self)

    def build_conformal_mesh(self):
        npnt = self.points.shape[0]
        if npnt >= 2_000_000:
            raise ValueError("Mesh: face keys are packed in int64; more than 2e6 points overflow")
        P = np.int64(npnt)
        lines = self.get_cells_type("line")
        tris = self.get_cells_type("triangle")
        tets = self.get_cells_type("tetra")

        def edge_key(pairs):
            a = np.minimum(pairs[..., 0], pairs[..., 1])
            b = np.maximum(pairs[..., 0], pairs[..., 1])
            return a * P + b

        def face_key(trip):
            s = _accel.sort_last(trip)
            return (s[..., 0] * P + s[..., 1]) * P + s[..., 2]

        # ---- numbering scan (tets, triangles, lines) R/src/mesh/mesh.py:252-324
        e_scan = np.concatenate(
            [edge_key(tets[:, TET_EDGES]).ravel(), edge_key(tris[:, TRI_EDGES]).ravel(), edge_key(lines)]
        )
        _, e_keys = _first_appearance(e_scan)
        f_scan = np.concatenate([face_key(tets[:, TET_FACES]).ravel(), face_key(tris)])
        _, f_keys = _first_appearance(f_scan)
        # volumes: distinct sorted node 4-tuples in first-appearance order; pack via (face key of 3 smallest, largest)
        ts = _accel.sort_last(tets)
        t_face = (ts[:, 0] * P + ts[:, 1]) * P + ts[:, 2]
        if len(tets):
            t_first, t_inv = _accel.unique_rows2(t_face, ts[:, 3])
            t_order = _accel.argsort(t_first)
            t_rank = np.empty(len(t_first), dtype=np.int64)
            t_rank[t_order] = np.arange(len(t_first))
            tet_vol = t_rank[t_inv.ravel()]
            n_vol = len(t_first)
            vol_first_tet = t_first[t_order]
        else:
            tet_vol = np.zeros(0, dtype=np.int64)
            n_vol = 0
            vol_first_tet = np.zeros(0, dtype=np.int64)

        nv, ne, nf = npnt, len(e_keys), len(f_keys)
        self.n_vertices, self.n_edges, self.n_faces, self.n_volumes = nv, ne, nf, n_vol
        self.offset = np.array([0, nv, nv + ne, nv + ne + nf, nv + ne + nf + n_vol], dtype=np.int64)
        self.n_cells = int(self.offset[4])
        edge_id = _SortedLookup(e_keys, nv + np.arange(ne, dtype=np.int64))
        face_id = _SortedLookup(f_keys, nv + ne + np.arange(nf, dtype=np.int64))

        # ---- materialisation in block order R/src/mesh/mesh.py:125-237
        e_cand_id, e_cand_nodes = [], []
        f_cand_id, f_cand_nodes = [], []
        material = np.full(self.n_cells, NO_MATERIAL, dtype=np.int64)
        created = np.zeros(self.n_cells, dtype=bool)
        tet_seen = 0
        vol_nodes = np.zeros((n_vol, 4), dtype=np.int64)
        vol_done = np.zeros(n_vol, dtype=bool)
        for name, conn, phys in self._blocks:
            if name == "vertex":
                material[conn[:, 0]] = phys
                created[conn[:, 0]] = True
            elif name == "line":
                ids = edge_id(edge_key(conn))
                e_cand_id.append(ids)
                e_cand_nodes.append(conn)
                material[ids] = phys
            elif name == "triangle":
                ids = face_id(face_key(conn))
                # insert_polygon creates the face, which first creates its (sorted) edges
                pairs = _accel.sort_last(conn[:, TRI_EDGES]).reshape(-1, 2)
                e_cand_id.append(edge_id(edge_key(pairs)))
                e_cand_nodes.append(pairs)
                f_cand_id.append(ids)
                f_cand_nodes.append(conn)
                material[ids] = phys
            elif name == "tetra":
                pairs = conn[:, TET_EDGES].reshape(-1, 2)
                e_cand_id.append(edge_id(edge_key(pairs)))
                e_cand_nodes.append(pairs)
                trip = conn[:, TET_FACES].reshape(-1, 3)
                f_cand_id.append(face_id(face_key(trip)))
                f_cand_nodes.append(trip)
                vids = tet_vol[tet_seen : tet_seen + len(conn)]
                tet_seen += len(conn)
                # first creator keeps node order; later duplicates only overwrite the material id
                uniq_v, first_v = _accel.unique(vids, return_index=True)
                fresh = ~vol_done[uniq_v]
                vol_nodes[uniq_v[fresh]] = conn[first_v[fresh]]
                vol_done[uniq_v] = True
                material[self.offset[3] + vids] = phys
            else:
                raise ValueError("Mesh: unsupported cell type %r" % name)

        def settle(cand_id, cand_nodes, base, count, width):
            nodes = np.full((count, width), -1, dtype=np.int64)
            if len(cand_id):
                ids = np.concatenate(cand_id) - base
                cn = np.concatenate(cand_nodes, axis=0)
                uniq, first = _accel.unique(ids, return_index=True)
                nodes[uniq] = cn[first]
            return nodes

        self.edge_nodes = settle(e_cand_id, e_cand_nodes, nv, ne, 2)
        self.face_nodes = settle(f_cand_id, f_cand_nodes, nv + ne, nf, 3)
        self.vol_nodes = vol_nodes
        self.material = material
        # sub-entity tables (R/src/mesh/mesh.py:150-216)
        self.face_edges = (
            edge_id(edge_key(_accel.sort_last(self.face_nodes[:, TRI_EDGES]))) if nf else np.zeros((0, 3), dtype=np.int64)
        )
        self.vol_edges = (
            edge_id(edge_key(self.vol_nodes[:, TET_EDGES])) if n_vol else np.zeros((0, 6), dtype=np.int64)
        )
        self.vol_faces = (
            face_id(face_key(self.vol_nodes[:, TET_FACES])) if n_vol else np.zeros((0, 4), dtype=np.int64)
        )
        self._vol_first_tet = vol_first_tet
        created[nv:] = True
        if ne:
            created[self.edge_nodes.ravel()] = True
        self._created_vertex = created[:nv]

    def find_cell_ids(self, dim, node_tags):
        """Global cell ids of the edges (``dim = 1``) / faces (``dim = 2``) with the given vertex sets ``[m, dim + 1]``
        (any vertex order); raises ``KeyError`` for a set that is not a cell of the mesh."""
        P = np.int64(self.points.shape[0])
        nodes = self.node_tags_of(dim)
        key = lambda a: (np.sort(a, axis=-1) * (P ** np.arange(dim, -1, -1, dtype=np.int64))).sum(axis=-1)
        if not hasattr(self, "_cell_lookup"):
            self._cell_lookup = {}
        if dim not in self._cell_lookup:
            self._cell_lookup[dim] = _SortedLookup(key(nodes), self.ids_of_dimension(dim))
        return self._cell_lookup[dim](key(np.asarray(node_tags, dtype=np.int64)))

    # -- array accessors used by the spaces layer ----------------------------------------------
    def cell_dimension(self, cell_ids):
        return np.searchsorted(self.offset[1:], np.asarray(cell_ids), side="right")

    def ids_of_dimension(self, dim):
        return np.arange(self.offset[dim], self.offset[dim + 1], dtype=np.int64)

    def node_tags_of(self, dim):
        if dim == 0:
            return np.arange(self.n_vertices, dtype=np.int64)[:, None]
        return (self.edge_nodes, self.face_nodes, self.vol_nodes)[dim - 1]

    def sub_cells_of(self, dim, d):
        """``sub_cells_ids[d]`` of every cell of dimension ``dim`` as an array ``[n_dim, n_sub]``."""
        if d > dim:
            raise ValueError("sub-entity dimension exceeds cell dimension")
        if d == dim:
            return self.ids_of_dimension(dim)[:, None]
        if d == 0:
            return self.node_tags_of(dim)  # vertex id == node tag (R/src/mesh/mesh.py:241-244)
        if dim == 2 and d == 1:
            return self.face_edges
        if dim == 3 and d == 1:
            return self.vol_edges
        if dim == 3 and d == 2:
            return self.vol_faces
        raise ValueError("no such sub-entity table")

    def _make_cell(self, cell_id):
        if cell_id < 0 or cell_id >= self.n_cells:
            raise IndexError(cell_id)
        dim = int(self.cell_dimension(cell_id))
        local = cell_id - int(self.offset[dim])
        if dim == 0 and not self._created_vertex[local]:
            return None
        cell = MeshCell(dim)
        cell.id = cell_id
        m = int(self.material[cell_id])
        cell.material_id = None if m == NO_MATERIAL else m
        cell.set_node_tags(np.array(self.node_tags_of(dim)[local]))
        for d in range(dim + 1):
            cell.set_sub_cells_ids(d, np.array(self.sub_cells_of(dim, d)[local]))
        return cell
