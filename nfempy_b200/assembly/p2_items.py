"""Row-item plan of the P2-tetrahedron fused assembly kernel (``nfa_assemble_p2tet_*``).

The per-cell stiffness of a quadratic Lagrange tetrahedron is written with the physical gradients ``g_m`` of the
barycentric coordinates: ``grad phi_v(i) = (4 lam_i - 1) g_i``, ``grad phi_e(ij) = 4 (lam_j g_i + lam_i g_j)``.  The
integrals of products of the scalar factors are universal constants (1/10, -1/30, 4/15, 2/15), so a block of the
local matrix (R/src/weak_forms/le_primal_weak_form.py:91-126 integrated exactly) is a short sum of outer products of
the ``g_m`` of the vertices involved -- no reference tensor look-up, and the same formula for every cell of the
star of an entity.  The global block row of node ``N`` is therefore cut into ITEMS, each a set of at most three
output blocks that receive contributions from exactly the same cells:

  row = vertex i   A  (v,v)                                   cells of the vertex star
                   B  (v,v'), (v,E)        per edge E=(v,v')  cells around E
                   C  (v,e_xy)             per face (v,x,y)   the two cells of the face
  row = edge (i,j) D  (E,v_i), (E,v_j), (E,E)                 cells around E
                   E  (E,v_x), (E,e_ix), (E,e_jx) per face    the two cells of the face
                   F  (E,E_opp)            per cell           one cell

One thread owns one item: it walks the item's cells in ascending cell order (the order of the reference's
``Mat.setValue(..., addv=True)`` calls, R/src/assembly/ScatterFormData.py:5-11), keeps the sums in registers and
writes each output block exactly once.  This module builds the lists the kernel walks, with torch tensor
operations only (they run on the GPU for production sizes and on the CPU in the ``-m "not gpu"`` tests).

Layout handed to the kernel:
  chunk_hdr  int32 [n_chunks, 4]   b0, n_blocks, task0, n_tasks      (chunk = contiguous block rows)
  task_hdr   int32 [n_tasks, 4]    type | kmax << 8 | (b0 & 1) << 16, ell_base, item_base, n_items (<= 32) | chunk << 6
  item_hdr   uint16 [n_items, 4]   tile offsets of the item's blocks (0xFFFF = unused), row pitch
  ell        uint32 [sum 32 kmax]  entry k of lane l of a task at ell_base + 32 k + l:
                                   cell << 6 | i << 4 | j << 2 | x   (0xFFFFFFFF = none)
  chunk_cells_ptr int32 [n_chunks + 1], chunk_cells int32: the cells each chunk reads (L1 prefetch list)
"""
import numpy as np
import torch

EDGE_PAIRS = [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)]  # local vertices of P2 nodes 4..9 (basix order)
EDGE_NODE = {}
for _n, (_i, _j) in enumerate(EDGE_PAIRS):
    EDGE_NODE[(_i, _j)] = EDGE_NODE[(_j, _i)] = 4 + _n
TYPE_A, TYPE_B, TYPE_C, TYPE_D, TYPE_E, TYPE_F = range(6)
NONE16 = 0xFFFF
NONE32 = 0xFFFFFFFF


def _templates():
    """Per local row node: (row node a, type, i, j, x, partner node, block nodes[3] (-1 = unused), swap rule).
    For edge rows the roles (i, j) are ordered per cell by the GLOBAL ids of the two vertices; ``alt`` holds the
    template with the roles exchanged."""
    out = []
    for i in range(4):
        out.append(dict(a=i, type=TYPE_A, ijx=(i, 0, 0), partner=i, blocks=(i, -1, -1)))
        for k in range(4):
            if k != i:
                e = EDGE_NODE[(i, k)]
                out.append(dict(a=i, type=TYPE_B, ijx=(i, k, 0), partner=e, blocks=(k, e, -1)))
        for z in range(4):
            if z != i:
                x, y = [v for v in range(4) if v not in (i, z)]
                e = EDGE_NODE[(x, y)]
                out.append(dict(a=i, type=TYPE_C, ijx=(i, z, 0), partner=e, blocks=(e, -1, -1)))
    for n, (i, j) in enumerate(EDGE_PAIRS):
        a = 4 + n
        others = [v for v in range(4) if v not in (i, j)]

        def edge_templates(i, j):
            t = [dict(a=a, type=TYPE_D, ijx=(i, j, 0), partner=a, blocks=(i, j, a))]
            for x in others:
                t.append(dict(a=a, type=TYPE_E, ijx=(i, j, x), partner=x,
                              blocks=(x, EDGE_NODE[(i, x)], EDGE_NODE[(j, x)])))
            t.append(dict(a=a, type=TYPE_F, ijx=(i, j, 0), partner=EDGE_NODE[tuple(others)],
                          blocks=(EDGE_NODE[tuple(others)], -1, -1)))
            return t

        for t, alt in zip(edge_templates(i, j), edge_templates(j, i)):
            t["alt"] = alt
            t["swap_on"] = (i, j)
            out.append(t)
    return out


TEMPLATES = _templates()  # 4 * 7 + 6 * 4 = 52 entries per cell


def build(node_dest, bptr, bcols, n_nodes, cap=448, riffle=148, batch=8, lane_order="cell", staged=True):
    """``node_dest`` int64 [n_cells, 10] global node of each local P2 node, ``bptr``/``bcols`` the block CSR of the
    matrix (int64 tensors on the same device).  Returns a dict of tensors (see module docstring) and sizes."""
    dev = node_dest.device
    n_cells = int(node_dest.shape[0])
    if n_cells >= (1 << 26):
        raise ValueError("p2 item plan: more than 2^26 cells")
    n_bnz = int(bcols.shape[0])
    nb_row = bptr[1:] - bptr[:-1]
    max_row = int(nb_row.max().item())
    if (cap + max_row) * 9 >= NONE16:
        raise ValueError("p2 item plan: chunk tile exceeds the 16-bit tile offsets")
    rows_of_block = torch.repeat_interleave(torch.arange(n_nodes, device=dev), nb_row)
    ukeys = rows_of_block * n_nodes + bcols

    def block_id(N, M):
        idx = torch.searchsorted(ukeys, N * n_nodes + M)
        return idx

    cells = torch.arange(n_cells, device=dev, dtype=torch.int64)
    T = len(TEMPLATES)
    rowN = torch.empty((T, n_cells), dtype=torch.int64, device=dev)
    typ = torch.empty((T, n_cells), dtype=torch.int64, device=dev)
    code = torch.empty((T, n_cells), dtype=torch.int64, device=dev)
    item_blk = torch.empty((T, n_cells), dtype=torch.int64, device=dev)
    blk = torch.full((T, n_cells, 3), -1, dtype=torch.int64, device=dev)
    for t, tpl in enumerate(TEMPLATES):
        N = node_dest[:, tpl["a"]]
        variants = [tpl]
        swap = None
        if "alt" in tpl:
            i0, j0 = tpl["swap_on"]
            swap = node_dest[:, i0] > node_dest[:, j0]
            variants.append(tpl["alt"])
        per_variant = []
        for v in variants:
            i, j, x = v["ijx"]
            c = (cells << 6) | (i << 4) | (j << 2) | x
            ib = block_id(N, node_dest[:, v["partner"]])
            bl = [block_id(N, node_dest[:, b]) if b >= 0 else torch.full_like(N, -1) for b in v["blocks"]]
            per_variant.append((c, ib, torch.stack(bl, dim=1)))
        if swap is None:
            c, ib, bl = per_variant[0]
        else:
            c = torch.where(swap, per_variant[1][0], per_variant[0][0])
            ib = torch.where(swap, per_variant[1][1], per_variant[0][1])
            bl = torch.where(swap[:, None], per_variant[1][2], per_variant[0][2])
        rowN[t], typ[t], code[t], item_blk[t], blk[t] = N, tpl["type"], c, ib, bl
    # cell-major entry order so that a stable sort keeps ascending cells inside an item
    rowN, typ, code, item_blk = (v.t().reshape(-1) for v in (rowN, typ, code, item_blk))
    blk = blk.permute(1, 0, 2).reshape(-1, 3)
    # a row belongs to the window its first block falls in; batches of ``batch`` consecutive windows (neighbouring rows:
    # they share cells, so geometry records stay in L1) are then dealt round-robin into ``riffle`` groups so that a
    # CTA's contiguous chunk range is a strided sample of the matrix (vertex and edge rows cost differently)
    win = bptr[:-1] // cap
    uniq_win = torch.unique(win[nb_row > 0])
    dense = torch.searchsorted(uniq_win, win)
    q, r = dense // batch, dense % batch
    n_batches = (int(uniq_win.shape[0]) + batch - 1) // batch
    per_group = (n_batches + riffle - 1) // riffle
    chunk_of_row = ((q % riffle) * per_group + q // riffle) * batch + r
    chunk_raw = chunk_of_row[rowN]
    key = (chunk_raw << 34) | (typ << 31) | item_blk
    if n_bnz >= (1 << 31):
        raise ValueError("p2 item plan: block count exceeds 2^31")
    key, order = torch.sort(key, stable=True)
    code, blk, rowN = code[order], blk[order], rowN[order]
    # items = runs of equal keys
    first = torch.ones_like(key, dtype=torch.bool)
    first[1:] = key[1:] != key[:-1]
    item_start = torch.nonzero(first).reshape(-1)
    n_items = int(item_start.shape[0])
    item_end = torch.cat([item_start[1:], torch.tensor([key.shape[0]], device=dev)])
    item_count = item_end - item_start
    item_key = key[item_start] >> 31  # chunk | type
    # order of the items (= lanes) inside a (chunk, type) group:
    #   "cell"  by the first cell they read: neighbouring lanes then read the same 128-byte geometry records in the same
    #           instruction (one L1 wavefront per distinct line), at the price of some spread in the loop counts;
    #   "count" by descending count: lanes of a task run equally long.
    if lane_order == "cell":
        ikey2 = (item_key << 26) | (code[item_start] >> 6)
        shift = 26
    else:
        ikey2 = (item_key << 16) | (65535 - torch.clamp(item_count, max=65535))
        shift = 16
    ikey2, iperm = torch.sort(ikey2, stable=True)
    grp = ikey2 >> shift
    gfirst = torch.ones_like(grp, dtype=torch.bool)
    gfirst[1:] = grp[1:] != grp[:-1]
    gstart = torch.nonzero(gfirst).reshape(-1)
    gid = torch.cumsum(gfirst.to(torch.int64), 0) - 1
    rank_in_group = torch.arange(n_items, device=dev) - gstart[gid]
    # tasks: 32 items of one (chunk, type) group
    task_local = rank_in_group // 32
    lane = rank_in_group % 32
    tfirst = gfirst | (lane == 0)
    task_of_item = torch.cumsum(tfirst.to(torch.int64), 0) - 1
    n_tasks = int(task_of_item[-1].item()) + 1
    cnt_sorted = item_count[iperm]
    kmax = torch.zeros(n_tasks, dtype=torch.int64, device=dev)
    kmax.scatter_reduce_(0, task_of_item, cnt_sorted, reduce="amax", include_self=True)
    if int(kmax.max().item()) > 255:
        raise ValueError("p2 item plan: a node belongs to more than 255 cells")
    task_items = torch.zeros(n_tasks, dtype=torch.int64, device=dev)
    task_items.scatter_add_(0, task_of_item, torch.ones_like(task_of_item))
    ell_base = torch.cumsum(kmax * 32, 0) - kmax * 32
    ell_size = int((kmax * 32).sum().item())
    task_first_item = torch.nonzero(tfirst).reshape(-1)
    task_group_key = grp[task_first_item]
    task_type = task_group_key & 7
    task_chunk_raw = task_group_key >> 3
    # ELL fill
    inv = torch.empty_like(iperm)
    inv[iperm] = torch.arange(n_items, device=dev)
    item_of_entry = torch.cumsum(first.to(torch.int64), 0) - 1
    r_in_item = torch.arange(key.shape[0], device=dev) - item_start[item_of_entry]
    pos = inv[item_of_entry]
    ell = torch.full((ell_size,), NONE32, dtype=torch.int64, device=dev)
    ell[ell_base[task_of_item[pos]] + r_in_item * 32 + lane[pos]] = code
    # item headers (sorted item order)
    it_first_entry = item_start[iperm]
    it_row = rowN[it_first_entry]
    it_blk = blk[it_first_entry]  # global block ids, -1 unused
    # chunks: compress the non-empty windows
    chunk_ids, chunk_index_of_task = torch.unique(task_chunk_raw, return_inverse=True)
    n_chunks = int(chunk_ids.shape[0])
    row_chunk_index = torch.searchsorted(chunk_ids, chunk_of_row)  # rows in empty windows do not occur
    # first block / block count of each chunk
    nonempty_rows = nb_row > 0
    c_b0 = torch.full((n_chunks,), n_bnz, dtype=torch.int64, device=dev)
    c_b1 = torch.zeros((n_chunks,), dtype=torch.int64, device=dev)
    rr = torch.nonzero(nonempty_rows).reshape(-1)
    rc = torch.clamp(row_chunk_index[rr], max=n_chunks - 1)
    valid = chunk_ids[rc] == chunk_of_row[rr]
    rr, rc = rr[valid], rc[valid]
    c_b0.scatter_reduce_(0, rc, bptr[rr], reduce="amin", include_self=True)
    c_b1.scatter_reduce_(0, rc, bptr[rr + 1], reduce="amax", include_self=True)
    item_chunk = torch.searchsorted(chunk_ids, chunk_of_row[it_row])
    b0_item = c_b0[item_chunk]
    rb = bptr[it_row] - b0_item
    nbr = nb_row[it_row]
    where = torch.where(it_blk >= 0, 9 * rb[:, None] + 3 * (it_blk - bptr[it_row][:, None]), torch.full_like(it_blk, NONE16))
    item_hdr = torch.cat([where, (3 * nbr)[:, None]], dim=1)
    if int(item_hdr.max().item()) > NONE16:
        raise ValueError("p2 item plan: tile offset overflow")
    tfirst_chunk = torch.ones(n_tasks, dtype=torch.bool, device=dev)
    tfirst_chunk[1:] = chunk_index_of_task[1:] != chunk_index_of_task[:-1]
    task0 = torch.nonzero(tfirst_chunk).reshape(-1)
    ntask = torch.cat([task0[1:], torch.tensor([n_tasks], device=dev)]) - task0
    chunk_hdr = torch.stack([c_b0, c_b1 - c_b0, task0, ntask], dim=1)
    # bit 16 of word 0: parity of the chunk's first block (the tile sits on the same 16-byte phase as its image)
    task_hdr = torch.stack(
        [task_type | (kmax << 8) | ((c_b0[chunk_index_of_task] & 1) << 16), ell_base, task_first_item,
         task_items | (chunk_index_of_task << 6)], dim=1)
    # cells each chunk reads (ascending, once each): the kernel prefetches the records of the next chunk into L1
    entry_chunk = torch.searchsorted(chunk_ids, key >> 34)
    cc = torch.unique((entry_chunk << 26) | (code >> 6))
    cell_chunk = cc >> 26
    chunk_cells = cc & ((1 << 26) - 1)
    chunk_cells_ptr = torch.zeros(n_chunks + 1, dtype=torch.int64, device=dev)
    chunk_cells_ptr[1:] = torch.cumsum(torch.bincount(cell_chunk, minlength=n_chunks), 0)
    max_tile = int((c_b1 - c_b0).max().item())
    base = {
        "chunk_hdr": chunk_hdr.to(torch.int32).contiguous(),
        "task_hdr": task_hdr.to(torch.int32).contiguous(),
        "item_hdr": item_hdr.to(torch.int32).to(torch.uint16 if hasattr(torch, "uint16") else torch.int16).contiguous(),
        "chunk_cells_ptr": chunk_cells_ptr.to(torch.int32).contiguous(),
        "chunk_cells": chunk_cells.to(torch.int32).contiguous(), "n_chunk_cells": int(chunk_cells.shape[0]),
        "ell": (ell & NONE32).to(torch.uint32 if hasattr(torch, "uint32") else torch.int64).contiguous(),
        "n_chunks": n_chunks, "n_tasks": n_tasks, "n_items": n_items, "ell_size": ell_size, "max_tile_blocks": max_tile,
        "n_entries": int(key.shape[0]),
    }
    if not staged:  # the lists of the (opt-in) staged kernel cost 0.45 GB at 1 M tets: only built on request
        return base
    # ---- staged variant: everything a chunk reads is copied to shared memory by bulk copies before its tasks run, so
    # the ELL entries name the cell by its SLOT in the chunk's cell list and the task headers hold chunk-relative offsets
    slot = torch.searchsorted(cc, (entry_chunk << 26) | (code >> 6)) - chunk_cells_ptr[entry_chunk]
    ell_local = torch.full((ell_size,), NONE32, dtype=torch.int64, device=dev)
    ell_local[ell_base[task_of_item[pos]] + r_in_item * 32 + lane[pos]] = (slot << 6) | (code & 63)
    ell_end = torch.cat([ell_base[1:], torch.tensor([ell_size], device=dev)])
    item_end_t = torch.cat([task_first_item[1:], torch.tensor([n_items], device=dev)])
    last = task0 + ntask - 1
    c_ell0, c_item0 = ell_base[task0], task_first_item[task0]
    c_nell, c_nitem = ell_end[last] - c_ell0, item_end_t[last] - c_item0
    item_src0 = c_item0 & ~1  # the bulk copy starts on a 16-byte boundary (item headers are 8 bytes)
    item_copy = ((c_item0 + c_nitem + 1) & ~1) - item_src0
    chunk_ext = torch.stack([c_ell0, c_nell, item_src0, item_copy], dim=1)
    task_hdr_rel = task_hdr.clone()
    task_hdr_rel[:, 1] = ell_base - c_ell0[chunk_index_of_task]
    task_hdr_rel[:, 2] = task_first_item - item_src0[chunk_index_of_task]
    n_cc = chunk_cells_ptr[1:] - chunk_cells_ptr[:-1]
    return {
        "ell_local": (ell_local & NONE32).to(torch.uint32 if hasattr(torch, "uint32") else torch.int64).contiguous(),
        "chunk_ext": chunk_ext.to(torch.int32).contiguous(),
        "task_hdr_rel": task_hdr_rel.to(torch.int32).contiguous(),
        "max_chunk_cells": int(n_cc.max().item()), "max_chunk_ell": int(c_nell.max().item()),
        "max_chunk_items": int(item_copy.max().item()), "max_chunk_tasks": int(ntask.max().item()),
        **base,
    }


# ---- reference walk of the plan (numpy): what the kernel does, used by the CPU tests ---------------------------
def geometry_records(coords, cell_nodes, lam, mu):
    """``[n_cells, 4, 4]``: (sqrt(det mu) g_m, lambda / mu) for the four vertices of every cell."""
    X = coords[cell_nodes]  # [n, 4, 3]
    J = np.transpose(X[:, 1:] - X[:, :1], (0, 2, 1))
    det = np.abs(np.linalg.det(J))
    Jinv = np.linalg.inv(J)
    g = np.concatenate([-Jinv.sum(axis=1, keepdims=True), Jinv], axis=1)
    rec = np.empty((len(X), 4, 4))
    rec[:, :, :3] = np.sqrt(det * mu)[:, None, None] * g
    rec[:, :, 3] = (lam / mu)[:, None]
    return rec


def _add(K, rho, u, w, c):
    O = np.outer(u, w)
    K += c * (rho * O + O.T + np.trace(O) * np.eye(3))


def item_blocks(typ, entries, rec):
    """The three 3x3 sums of one item (unused ones stay zero).  ``entries``: packed codes in ascending cell order."""
    K = np.zeros((3, 3, 3))
    for e in entries:
        cell, i, j, x = int(e) >> 6, (int(e) >> 4) & 3, (int(e) >> 2) & 3, int(e) & 3
        g, rho = rec[cell, :, :3], rec[cell, 0, 3]
        gi, gj, gx = g[i], g[j], g[x]
        if typ == TYPE_A:
            _add(K[0], rho, gi, gi, 0.1)
        elif typ == TYPE_B:
            _add(K[0], rho, gi, gj, -1.0 / 30.0)
            _add(K[1], rho, gi, 0.1 * gj - gi / 30.0, 1.0)
        elif typ == TYPE_C:
            _add(K[0], rho, gi, gi + gj, 1.0 / 30.0)  # g_x + g_y = -(g_i + g_z)
        elif typ == TYPE_D:
            s = gi + gj
            _add(K[0], rho, 0.1 * gj - gi / 30.0, gi, 1.0)
            _add(K[1], rho, 0.1 * gi - gj / 30.0, gj, 1.0)
            _add(K[2], rho, s, s, 2.0 / 15.0)
            _add(K[2], rho, gi, gi, 2.0 / 15.0)
            _add(K[2], rho, gj, gj, 2.0 / 15.0)
        elif typ == TYPE_E:
            s = gi + gj
            _add(K[0], rho, s, gx, -1.0 / 30.0)
            _add(K[1], rho, s, gi + gx, 2.0 / 15.0)
            _add(K[1], rho, gj, gx, 2.0 / 15.0)
            _add(K[2], rho, s, gj + gx, 2.0 / 15.0)
            _add(K[2], rho, gi, gx, 2.0 / 15.0)
        else:
            s = gi + gj
            _add(K[0], rho, s, s, -2.0 / 15.0)  # opposite edge: g_x + g_y = -(g_i + g_j)
    return K


def walk(plan, rec, n_bnz, staged=False):
    """CSR block values ``[n_bnz * 9]`` in the scalar CSR order of the fused path (row pitch = 3 * blocks in row).
    ``staged``: walk the lists of the shared-memory staged kernel instead (chunk-local copies of the ELL entries, item
    headers, task headers and geometry records, addressed by chunk-relative offsets and cell slots)."""
    chunk_hdr = plan["chunk_hdr"].cpu().numpy().astype(np.int64)
    task_hdr = plan["task_hdr_rel" if staged else "task_hdr"].cpu().numpy().astype(np.int64)
    item_hdr = plan["item_hdr"].cpu().numpy().astype(np.int64) & 0xFFFF
    ell = plan["ell_local" if staged else "ell"].cpu().numpy().astype(np.int64) & NONE32
    if staged:
        chunk_ext = plan["chunk_ext"].cpu().numpy().astype(np.int64)
        cells_ptr = plan["chunk_cells_ptr"].cpu().numpy().astype(np.int64)
        chunk_cells = plan["chunk_cells"].cpu().numpy().astype(np.int64)
    vals = np.full(n_bnz * 9, np.nan)
    ell_all, item_all, rec_all = ell, item_hdr, rec
    for ci, (b0, nblk, task0, ntask) in enumerate(chunk_hdr):
        tile = np.full(nblk * 9, np.nan)
        if staged:  # what the bulk copies put into the chunk's buffer
            e0, ne, i0, ni = chunk_ext[ci]
            assert e0 % 32 == 0 and ne % 32 == 0 and i0 % 2 == 0 and ni % 2 == 0
            assert ne <= plan["max_chunk_ell"] and ni <= plan["max_chunk_items"] and ntask <= plan["max_chunk_tasks"]
            ell, item_hdr = ell_all[e0 : e0 + ne], item_all[i0 : i0 + ni]
            cl = chunk_cells[cells_ptr[ci] : cells_ptr[ci + 1]]
            assert len(cl) <= plan["max_chunk_cells"]
            rec = rec_all[cl]
        for t in range(task0, task0 + ntask):
            typ, kmax = task_hdr[t, 0] & 0xFF, (task_hdr[t, 0] >> 8) & 0xFF
            assert ((task_hdr[t, 0] >> 16) & 1) == (b0 & 1)
            base, item0, nit = task_hdr[t, 1], task_hdr[t, 2], task_hdr[t, 3] & 63
            assert task_hdr[t, 3] >> 6 == ci
            for lane in range(nit):
                entries = [ell[base + 32 * k + lane] for k in range(kmax)]
                entries = [e for e in entries if e != NONE32]
                K = item_blocks(typ, entries, rec)
                w0, w1, w2, pitch = item_hdr[item0 + lane]
                for k, w in enumerate((w0, w1, w2)):
                    if w == NONE16:
                        continue
                    for r in range(3):
                        assert np.all(np.isnan(tile[w + r * pitch : w + r * pitch + 3])), "block written twice"
                        tile[w + r * pitch : w + r * pitch + 3] = K[k][r]
        vals[9 * b0 : 9 * (b0 + nblk)] = tile
    return vals
