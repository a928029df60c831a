"""ORACLE (test infrastructure only -- never imported by the product path).

Cell-by-cell restatement of the reference mesh layer, kept deliberately slow and literal:

* ``build_conformal_mesh``  follows R/src/mesh/mesh.py:239-339 (entity id dictionaries, scan order tets ->
  triangles -> lines) and R/src/mesh/mesh.py:125-237 (recursive cell materialisation in block order);
* ``entity_tables``          follows R/src/mesh/mesh.py:668-681 + R/src/mesh/mesh_topology.py:16-47
  (``networkx.from_edgelist`` node order == first appearance in the edge list).

Parity status: a line-by-line reading of the reference, PINNED against it: tools/make_reference_pins.py runs the reference's own
``Mesh.build_conformal_mesh`` (over a meshio stand-in) on 45 meshes and tests/test_cpu_reference_pins.py compares every cell's
dimension, node tags, sub-cell ids and material id of this class -- and of the product's vectorised mesh -- with its output.
"""
import numpy as np


class Cell:
    def __init__(self, dim, node_tags, cid, material):
        self.dimension = dim
        self.node_tags = np.array(node_tags, dtype=np.int64)
        self.id = cid
        self.material_id = material
        self.sub_cells_ids = [np.zeros(0, dtype=np.int64) for _ in range(dim + 1)]


class RefMesh:
    def __init__(self, dimension, points, blocks):
        self.dimension = dimension
        self.points = np.asarray(points, dtype=np.float64)
        self.blocks = blocks
        self.cells = []

    def _cells_of(self, name):
        return [np.asarray(row) for t, c, _ in self.blocks if t == name for row in c]

    def build_conformal_mesh(self):
        e0 = {i: i for i in range(len(self.points))}
        e1, e2, e3 = {}, {}, {}
        tets, tris, lines = self._cells_of("tetra"), self._cells_of("triangle"), self._cells_of("line")
        nxt = len(e0)
        for n in tets:  # mesh.py:252-268
            for a, b in ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)):
                k = tuple(sorted((int(n[a]), int(n[b]))))
                if k not in e1:
                    e1[k] = nxt
                    nxt += 1
        for n in tris:  # mesh.py:270-280
            for a, b in ((1, 2), (0, 2), (0, 1)):
                k = tuple(sorted((int(n[a]), int(n[b]))))
                if k not in e1:
                    e1[k] = nxt
                    nxt += 1
        for n in lines:  # mesh.py:282-288
            k = tuple(sorted(int(v) for v in n))
            if k not in e1:
                e1[k] = nxt
                nxt += 1
        for n in tets:  # mesh.py:290-303
            for f in ((1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2)):
                k = tuple(sorted(int(n[i]) for i in f))
                if k not in e2:
                    e2[k] = nxt
                    nxt += 1
        for n in tris:  # mesh.py:305-311
            k = tuple(sorted(int(v) for v in n))
            if k not in e2:
                e2[k] = nxt
                nxt += 1
        for n in tets:  # mesh.py:313-324
            k = tuple(sorted(int(v) for v in n))
            if k not in e3:
                e3[k] = nxt
                nxt += 1
        self.cells = [None] * nxt
        self._e = (e0, e1, e2, e3)
        for name, conn, phys in self.blocks:  # mesh.py:334-337
            for n, p in zip(conn, phys):
                n = np.asarray(n, dtype=np.int64)
                p = int(p)
                if name == "vertex":
                    self._vertex(int(n[0]), p)
                elif name == "line":
                    self._edge(n, p)
                elif name == "triangle":
                    self._polygon(n, p)
                elif name == "tetra":
                    self._polyhedron(n, p)

    def _vertex(self, tag, material=None):  # mesh.py:125-135
        cid = self._e[0][tag]
        if self.cells[cid] is None:
            c = Cell(0, [tag], cid, material)
            c.sub_cells_ids[0] = np.array([tag])
            self.cells[cid] = c
        if material is not None:
            self.cells[cid].material_id = material
        return cid

    def _edge(self, n, material=None):  # mesh.py:137-151
        cid = self._e[1][tuple(sorted(int(v) for v in n))]
        if self.cells[cid] is None:
            vids = [self._vertex(int(t)) for t in n]
            c = Cell(1, n, cid, material)
            c.sub_cells_ids[0] = np.array(vids)
            c.sub_cells_ids[1] = np.array([cid])
            self.cells[cid] = c
        if material is not None:
            self.cells[cid].material_id = material
        return cid

    def _polygon(self, n, material=None):  # mesh.py:153-178
        cid = self._e[2][tuple(sorted(int(v) for v in n))]
        if self.cells[cid] is None:
            vids = [self._e[0][int(t)] for t in n]
            eids = [self._edge(np.sort(n[[a, b]])) for a, b in ((1, 2), (0, 2), (0, 1))]
            c = Cell(2, n, cid, material)
            c.sub_cells_ids[0] = np.array(vids)
            c.sub_cells_ids[1] = np.array(eids)
            c.sub_cells_ids[2] = np.array([cid])
            self.cells[cid] = c
        if material is not None:
            self.cells[cid].material_id = material
        return cid

    def _polyhedron(self, n, material=None):  # mesh.py:180-219
        cid = self._e[3][tuple(sorted(int(v) for v in n))]
        if self.cells[cid] is None:
            vids = [self._e[0][int(t)] for t in n]
            eids = [self._edge(n[[a, b]]) for a, b in ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))]
            fids = [self._polygon(n[list(f)]) for f in ((1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2))]
            c = Cell(3, n, cid, material)
            c.sub_cells_ids[0] = np.array(vids)
            c.sub_cells_ids[1] = np.array(eids)
            c.sub_cells_ids[2] = np.array(fids)
            c.sub_cells_ids[3] = np.array([cid])
            self.cells[cid] = c
        if material is not None:
            self.cells[cid].material_id = material
        return cid

    # -- topology on physical tags -----------------------------------------------------------------
    def entity_tables(self, physical_tags, dim):
        """Returns ``(selected cell ids, entity_ids[d] lists)`` as ``MeshTopology.build_data_on_pysical_tags``
        would (mesh.py:668-681, mesh_topology.py:22-47): a digraph's node order is first appearance."""
        selected = [
            c for c in self.cells if c is not None and c.dimension == dim and c.material_id in physical_tags
        ]
        entity_ids = {}
        for d in range(dim + 1):
            seen, order = set(), []
            for c in selected:
                for sid in c.sub_cells_ids[d]:
                    for node in ((c.dimension, c.id), (self.cells[sid].dimension, self.cells[sid].id)):
                        if node not in seen:
                            seen.add(node)
                            order.append(node)
            entity_ids[d] = [i for dd, i in order if dd == d]
        return [c.id for c in selected], entity_ids
