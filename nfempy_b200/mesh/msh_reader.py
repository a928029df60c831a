"""Minimal MSH 4.1 (ASCII) reader producing what the reference gets from meshio.

The reference calls ``meshio.read(file_name)`` (R/src/mesh/mesh.py:19) and then uses
``points``, ``cells`` (blocks in file order) and ``cell_data["gmsh:physical"]``
(R/src/mesh/mesh.py:247-250, 334-337).  meshio is absent here, so this module restates the
part of its gmsh-4.1 reader that matters for numbering:

* points are stored in ``$Nodes`` file order and node tags are remapped to that position;
* one cell block per ``$Elements`` entity block, in file order;
* the block's physical tag is the first physical tag of its geometric entity.
"""
import numpy as np

_GMSH_TYPES = {15: ("vertex", 1), 1: ("line", 2), 2: ("triangle", 3), 4: ("tetra", 4)}


def _section(lines, name):
    try:
        b = lines.index("$" + name)
        e = lines.index("$End" + name)
    except ValueError:
        return None
    return lines[b + 1 : e]


def read_msh(file_name):
    """Returns ``(points[n,3] f64, blocks)`` with ``blocks = [(type, conn[n,k] i64, physical[n] i64)]``."""
    with open(file_name, "r") as f:
        lines = [ln.strip() for ln in f]
    fmt = _section(lines, "MeshFormat")
    if fmt is None or not fmt[0].startswith("4.1"):
        raise ValueError("read_msh: only MSH 4.1 ASCII is supported: %r" % (fmt,))
    if int(fmt[0].split()[1]) != 0:
        raise ValueError("read_msh: binary MSH files are not supported")

    # --- $Entities: (dim, tag) -> physical tags
    physical = [{}, {}, {}, {}]
    ent = _section(lines, "Entities")
    if ent is not None:
        counts = [int(v) for v in ent[0].split()]
        row = 1
        for dim in range(4):
            for _ in range(counts[dim]):
                tok = ent[row].split()
                row += 1
                tag = int(tok[0])
                off = 4 if dim == 0 else 7
                n_phys = int(tok[off])
                physical[dim][tag] = [abs(int(v)) for v in tok[off + 1 : off + 1 + n_phys]]

    # --- $Nodes
    nod = _section(lines, "Nodes")
    n_blocks, n_nodes = [int(v) for v in nod[0].split()[:2]]
    points = np.empty((n_nodes, 3), dtype=np.float64)
    tags = np.empty(n_nodes, dtype=np.int64)
    row, idx = 1, 0
    for _ in range(n_blocks):
        _, _, parametric, nb = [int(v) for v in nod[row].split()]
        row += 1
        tags[idx : idx + nb] = [int(v) for v in nod[row : row + nb]]
        row += nb
        for k in range(nb):
            points[idx + k] = [float(v) for v in nod[row + k].split()[:3]]
        row += nb
        idx += nb
    inv_tags = np.full(int(tags.max()) + 1, -1, dtype=np.int64)
    inv_tags[tags] = np.arange(n_nodes)

    # --- $Elements
    ele = _section(lines, "Elements")
    n_blocks = int(ele[0].split()[0])
    blocks = []
    row = 1
    for _ in range(n_blocks):
        dim, etag, etype, nb = [int(v) for v in ele[row].split()]
        row += 1
        if etype not in _GMSH_TYPES:
            raise ValueError("read_msh: unsupported gmsh element type %d" % etype)
        name, nn = _GMSH_TYPES[etype]
        data = np.array(
            [[int(v) for v in ele[row + k].split()] for k in range(nb)], dtype=np.int64
        ).reshape(nb, nn + 1)
        row += nb
        phys = physical[dim].get(etag, [])
        if len(phys) == 0:
            # meshio appends no "gmsh:physical" entry for such a block, and the reference then
            # zips blocks and tags out of step (R/src/mesh/mesh.py:334-337); refuse instead.
            raise ValueError(
                "read_msh: element block on entity (%d,%d) carries no physical tag" % (dim, etag)
            )
        blocks.append((name, inv_tags[data[:, 1:]], np.full(nb, phys[0], dtype=np.int64)))
    return points, blocks
