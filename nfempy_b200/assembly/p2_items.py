"""Row-item plan of the P2-tetrahedron fused assembly kernel (``nfa_assemble_p2tet_*``).

The per-cell stiffness of a quadratic Lagrange tetrahedron is written with the physical gradients ``g_m`` of the
barycentric coordinates: ``grad phi_v(i) = (4 lam_i - 1) g_i``, ``grad phi_e(ij) = 4 (lam_j g_i + lam_i g_j)``.  The
integrals of products of the scalar factors are universal constants (1/10, -1/30, 4/15, 2/15), so a block of the
local matrix (R/src/weak_forms/le_primal_weak_form.py:91-126 integrated exactly) is a short sum of outer products of
the ``g_m`` of the vertices involved -- no reference tensor look-up, and the same formula for every cell of the
star of an entity.  The global block row of node ``N`` is therefore cut into ITEMS, each a set of output blocks that
receive contributions from exactly the same cells:

  row = vertex i   A  (v,v)                                   cells of the vertex star
                   B  (v,v'), (v,E)        per edge E=(v,v')  cells around E
                   C  (v,e_xy)             per face (v,x,y)   the two cells of the face
  row = edge (i,j) D  (E,v_i), (E,v_j), (E,E)                 cells around E
                   E  (E,v_x), (E,e_ix), (E,e_jx) per face    the two cells of the face
                      + (E,E_opp) of each of the two cells for which this face is the DESIGNATED one (the face of the
                        cell around E whose third vertex has the smaller global id): that block has a single
                        contribution, -2/15 Kf(s s^T) with s = g_i + g_j -- two vectors the item holds anyway

(46 (item, cell) entries per tet; until round 2 the opposite-edge block was an item of its own, "F": 6 more entries, 12
more record loads and 6 more items per tet.)  One thread owns one item: it walks the item's cells in ascending cell
order (the order of the reference's ``Mat.setValue(..., addv=True)`` calls, R/src/assembly/ScatterFormData.py:5-11),
keeps the sums in registers and writes each output block exactly once.  This module builds the lists the kernel
walks, with torch tensor operations (they run on the GPU for production sizes and on the CPU in the
``-m "not gpu"`` tests).

Layout handed to the kernel:
  chunk_hdr  int32 [n_chunks, 4]   b0, n_blocks, task0, n_tasks      (chunk = contiguous block rows)
  task_hdr   int32 [n_tasks, 4]    type | kmax << 8 | (b0 & 1) << 16, ell_base, item_slot, n_items (<= 32) | chunk << 6
  item_hdr   uint16 [n_slots, 4]   one slot per item: tile offsets of the item's three blocks (0xFFFF = unused), row
                                   pitch; type D and E items take two slots: the second one holds, for type E, the tile
                                   offsets of the opposite-edge blocks of the item's first / second cell (0xFFFF = not
                                   emitted here) and, for type D, the row's node id (low, high half).  Type A items carry
                                   their node id in the two unused block slots.  (Items A and D own the node's three
                                   right-hand-side rows, nfa_assemble_p2tet_elasticity_load.)  A task's slots are
                                   contiguous and start on an even slot (16-byte aligned)
  ell        uint32 [sum 32 kmax]  entry k of lane l of a task at ell_base + 32 k + l:
                                   cell << 6 | i << 4 | j << 2 | x   (0xFFFFFFFF = none)
"""
import numpy as np
import torch

EDGE_PAIRS = [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)]  # local vertices of P2 nodes 4..9 (basix order)
EDGE_NODE = {}
for _n, (_i, _j) in enumerate(EDGE_PAIRS):
    EDGE_NODE[(_i, _j)] = EDGE_NODE[(_j, _i)] = 4 + _n
TYPE_A, TYPE_B, TYPE_C, TYPE_D, TYPE_E = range(5)
NONE16 = 0xFFFF
NONE32 = 0xFFFFFFFF
N_SIG = 15  # shared-memory store positions of an item: 3 blocks x 3 rows + 2 opposite-edge blocks x 3 rows


def _templates():
    """Per local row node: (row node a, type, i, j, x, partner node, block nodes[3] (-1 = unused), swap rule).
    For edge rows the roles (i, j) are ordered per cell by the GLOBAL ids of the two vertices; ``alt`` holds the
    template with the roles exchanged.  Type E templates also name the fourth vertex ``y`` and the opposite edge."""
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
                y = [v for v in others if v != x][0]
                t.append(dict(a=a, type=TYPE_E, ijx=(i, j, x), partner=x, y=y, opp=EDGE_NODE[tuple(others)],
                              blocks=(x, EDGE_NODE[(i, x)], EDGE_NODE[(j, x)])))
            return t

        for t, alt in zip(edge_templates(i, j), edge_templates(j, i)):
            t["alt"] = alt
            t["swap_on"] = (i, j)
            out.append(t)
    return out


TEMPLATES = _templates()  # 4 * 7 + 6 * 3 = 46 entries per cell


def _template_tables(dev):
    """The templates as index tensors on ``dev``: row node ``a``, ``type``, roles ``i, j, x``, ``partner``, block nodes
    ``b0..b2`` (-1 = unused) and the same for the role-swapped twin (``*_alt``; equal to the template where there is none),
    the vertices compared to decide the swap, and for type E the fourth vertex ``y`` and the opposite edge ``opp``."""
    cols = {k: [] for k in ("a", "type", "i", "j", "x", "partner", "b0", "b1", "b2", "i_alt", "j_alt", "x_alt", "partner_alt",
                            "b0_alt", "b1_alt", "b2_alt", "has_alt", "swap_i", "swap_j", "y", "opp")}
    for tpl in TEMPLATES:
        alt = tpl.get("alt", tpl)
        cols["a"].append(tpl["a"])
        cols["type"].append(tpl["type"])
        for suffix, v in (("", tpl), ("_alt", alt)):
            i, j, x = v["ijx"]
            cols["i" + suffix].append(i)
            cols["j" + suffix].append(j)
            cols["x" + suffix].append(x)
            cols["partner" + suffix].append(v["partner"])
            for k in range(3):
                cols["b%d%s" % (k, suffix)].append(v["blocks"][k])
        cols["has_alt"].append("alt" in tpl)
        si, sj = tpl.get("swap_on", (0, 0))
        cols["swap_i"].append(si)
        cols["swap_j"].append(sj)
        cols["y"].append(tpl.get("y", -1))
        cols["opp"].append(tpl.get("opp", -1))
    return {k: torch.as_tensor(v, dtype=torch.bool if k == "has_alt" else torch.int64, device=dev) for k, v in cols.items()}


def store_signature(hdr):
    """``[n, N_SIG]`` bank-pair (8-byte word mod 16) of every 3-double row segment an item stores into the tile, -1 for
    unused blocks.  ``hdr``: ``[n, 6]`` = three block offsets, pitch, two opposite-edge block offsets (NONE16 = unused).
    Two lanes of a half warp whose signatures agree at a position serialise that ``STS.64`` (bank conflict)."""
    hdr = np.asarray(hdr, dtype=np.int64)
    sig = np.full((len(hdr), N_SIG), -1, dtype=np.int64)
    for slot, col in enumerate((0, 1, 2, 4, 5)):
        v = hdr[:, col] != NONE16
        for r in range(3):
            sig[v, 3 * slot + r] = (hdr[v, col] + r * hdr[v, 3]) % 16
    return sig


def greedy_lane_order(group_ptr, hdr, window=48):
    """Reference (host, test-sized) statement of ``nfa_p2_lane_order``: inside every group the items are dealt into half
    warps of 16 so that the half warp's tile stores hit distinct shared-memory banks as far as possible.  Greedy: the next
    lane takes, among the next ``window`` unassigned items (the group is in first-cell order, so locality is kept), the
    one that raises the fewest per-position maxima of the bank counts, ties by the smaller sum of counts, then by
    position.  Returns the permutation (new order -> old index)."""
    group_ptr = np.asarray(group_ptr, dtype=np.int64)
    sig = store_signature(np.asarray(hdr))
    out = np.empty(len(sig), dtype=np.int64)
    for g in range(len(group_ptr) - 1):
        g0, g1 = int(group_ptr[g]), int(group_ptr[g + 1])
        un = list(range(g0, g1))
        pos = g0
        while un:
            cnt = np.zeros((N_SIG, 16), dtype=np.int64)
            mx = np.zeros(N_SIG, dtype=np.int64)
            for _ in range(min(16, len(un))):
                best, best_key = None, None
                for c in un[:window]:
                    inc = tie = 0
                    for p in range(N_SIG):
                        b = sig[c, p]
                        if b >= 0:
                            cc = cnt[p, b]
                            inc += 1 if cc + 1 > mx[p] else 0
                            tie += cc
                    k = (inc, tie)
                    if best_key is None or k < best_key:
                        best, best_key = c, k
                un.remove(best)
                out[pos] = best
                pos += 1
                for p in range(N_SIG):
                    b = sig[best, p]
                    if b >= 0:
                        cnt[p, b] += 1
                        mx[p] = max(mx[p], cnt[p, b])
    return out


def build(node_dest, bptr, bcols, n_nodes, cap=640, riffle=148, batch=8, lane_order="cell", lane_optimizer=None,
          locality=True, type_order=(0, 1, 2, 3, 4), cell_live=None):
    """``node_dest`` int64 [n_cells, 10] global node of each local P2 node, ``bptr``/``bcols`` the block CSR of the
    matrix (int64 tensors on the same device).  ``lane_optimizer(group_ptr, hdr6) -> permutation`` reorders the items
    inside their (chunk, type) groups (``nfa_p2_lane_order`` on the device, ``greedy_lane_order`` in the tests; None keeps
    the first-cell order).  ``type_order[t]`` = position of item type t (A..E) in a chunk's task list.  Returns a dict of
    tensors (see module docstring) and sizes.  ``cell_live`` (bool [n_cells], optional): cells marked False contribute zeros
    by contract (the halo cells of a rank's local plan: they only complete the sparsity pattern) -- their entries are
    left out of the lists, the items (and the zeros they write) stay."""
    dev = node_dest.device
    n_cells = int(node_dest.shape[0])
    if n_cells >= (1 << 26):
        raise ValueError("p2 item plan: more than 2^26 cells")
    n_bnz = int(bcols.shape[0])
    if n_bnz >= (1 << 31):
        raise ValueError("p2 item plan: block count exceeds 2^31")
    nb_row = bptr[1:] - bptr[:-1]
    max_row = int(nb_row.max().item())
    if (cap + max_row) * 9 >= NONE16:
        raise ValueError("p2 item plan: chunk tile exceeds the 16-bit tile offsets")
    rows_of_block = torch.repeat_interleave(torch.arange(n_nodes, device=dev), nb_row)
    ukeys = rows_of_block * n_nodes + bcols

    def block_id(N, M):
        return torch.searchsorted(ukeys, N * n_nodes + M)

    # ---- chunks: a row belongs to the window its first block falls in; batches of ``batch`` consecutive windows
    # (neighbouring rows: they share cells, so geometry records stay in L1) are dealt round-robin into ``riffle`` groups so
    # that a CTA's contiguous chunk range is a strided sample of the matrix (vertex and edge rows cost differently)
    win = bptr[:-1] // cap
    nonempty = nb_row > 0
    uniq_win = torch.unique(win[nonempty])
    dense_win = torch.clamp(torch.searchsorted(uniq_win, win), max=int(uniq_win.shape[0]) - 1)
    if locality:
        # walk order of the windows = ascending first cell they read (not CSR order): vertex rows and edge rows of the same
        # part of the mesh are then assembled at about the same time, so a cell's geometry record is fetched from DRAM
        # once instead of once per phase (the CSR itself keeps the reference numbering; only the walk order changes)
        cells10 = torch.arange(n_cells, device=dev, dtype=torch.int64).repeat_interleave(node_dest.shape[1])
        first_cell = torch.full((n_nodes,), n_cells, dtype=torch.int64, device=dev)
        first_cell.scatter_reduce_(0, node_dest.reshape(-1), cells10, reduce="amin", include_self=True)
        win_key = torch.full((int(uniq_win.shape[0]),), n_cells, dtype=torch.int64, device=dev)
        rr0 = torch.nonzero(nonempty).reshape(-1)
        win_key.scatter_reduce_(0, dense_win[rr0], first_cell[rr0], reduce="amin", include_self=True)
        seq = torch.empty_like(win_key)
        seq[torch.argsort(win_key, stable=True)] = torch.arange(int(win_key.shape[0]), device=dev)
        dense_win = seq[dense_win]
    q, r = dense_win // batch, dense_win % batch
    n_batches = (int(uniq_win.shape[0]) + batch - 1) // batch
    per_group = (n_batches + riffle - 1) // riffle
    chunk_raw_of_row = ((q % riffle) * per_group + q // riffle) * batch + r
    chunk_ids = torch.unique(chunk_raw_of_row[nonempty])
    n_chunks = int(chunk_ids.shape[0])
    chunk_of_row = torch.clamp(torch.searchsorted(chunk_ids, chunk_raw_of_row), max=n_chunks - 1)  # dense, processing order
    rr = torch.nonzero(nonempty).reshape(-1)
    c_b0 = torch.full((n_chunks,), n_bnz, dtype=torch.int64, device=dev)
    c_b1 = torch.zeros((n_chunks,), dtype=torch.int64, device=dev)
    c_b0.scatter_reduce_(0, chunk_of_row[rr], bptr[rr], reduce="amin", include_self=True)
    c_b1.scatter_reduce_(0, chunk_of_row[rr], bptr[rr + 1], reduce="amax", include_self=True)

    # ---- entries (cell, template), cell-major so that a stable sort keeps ascending cells inside an item.  All templates
    # are expanded at once (a handful of tensor operations over [n_cells, 46] arrays instead of a Python loop of ~20 launches
    # per template): local indices per (cell, template) are picked between the template and its role-swapped twin, then
    # gathered from node_dest; block ids come from three searches over the block keys.
    tt = _template_tables(dev)
    T = len(TEMPLATES)
    cells = torch.arange(n_cells, device=dev, dtype=torch.int64)
    swap = tt["has_alt"][None, :] & (node_dest[:, tt["swap_i"]] > node_dest[:, tt["swap_j"]])  # [n_cells, T]

    def pick(name):  # [n_cells, T] local index, template or twin
        return torch.where(swap, tt[name + "_alt"][None, :], tt[name][None, :])

    def nodes_of(local):  # global node of a local index table [n_cells, T]; -1 stays -1
        return torch.where(local >= 0, torch.gather(node_dest, 1, local.clamp(min=0)), torch.full_like(local, -1))

    li, lj, lx = pick("i"), pick("j"), pick("x")
    rowN = node_dest[:, tt["a"]]  # the row node does not depend on the roles
    code = (cells[:, None] << 6) | (li << 4) | (lj << 2) | lx
    typ = tt["type"][None, :].expand(n_cells, T)
    item_blk = block_id(rowN, nodes_of(pick("partner")))
    blk_nodes = torch.stack([nodes_of(pick("b%d" % k)) for k in range(3)], dim=2)  # [n_cells, T, 3]
    blk = torch.where(blk_nodes >= 0, block_id(rowN[:, :, None].expand(-1, -1, 3), blk_nodes.clamp(min=0)),
                      torch.full_like(blk_nodes, -1))
    # type E: the opposite-edge block of (row edge, cell) is emitted by the item of the cell's face around the edge whose
    # third vertex has the smaller global id (x and y, the fourth vertex, do not depend on the roles)
    is_e = tt["y"] >= 0
    gx = node_dest[:, tt["x"].clamp(min=0)]
    gy = node_dest[:, tt["y"].clamp(min=0)]
    designated = is_e[None, :] & (gx < gy)
    fblk = torch.where(designated, block_id(rowN, node_dest[:, tt["opp"].clamp(min=0)]), torch.full_like(rowN, -1))
    rowN, typ, code, item_blk, fblk = (v.reshape(-1) for v in (rowN, typ, code, item_blk, fblk))
    blk = blk.reshape(-1, 3)
    # tasks of a chunk run in ascending (sort rank of the) type: long items first, so that the last tasks of a chunk --
    # the ones that decide when its tile can be handed over -- are short ones
    rank_of_type = torch.as_tensor(type_order, dtype=torch.int64, device=dev)
    key = (chunk_of_row[rowN] << 34) | (rank_of_type[typ] << 31) | item_blk
    if cell_live is not None:
        # dead entries go to the end of their item (one more key bit below the item key) and are cut off the lists below
        dead = (~cell_live.to(torch.bool))[:, None].expand(n_cells, T).reshape(-1).to(torch.int64)
        key2, order = torch.sort((key << 1) | dead, stable=True)
        key, dead = key2 >> 1, (key2 & 1).to(torch.bool)
    else:
        key, order = torch.sort(key, stable=True)
        dead = None
    code, blk, rowN, fblk = code[order], blk[order], rowN[order], fblk[order]
    n_entries = int(key.shape[0])

    # ---- items = runs of equal keys (index: key-sorted item order)
    first = torch.ones_like(key, dtype=torch.bool)
    first[1:] = key[1:] != key[:-1]
    item_start = torch.nonzero(first).reshape(-1)
    n_items = int(item_start.shape[0])
    item_end = torch.cat([item_start[1:], torch.tensor([n_entries], device=dev)])
    item_count = item_end - item_start
    if dead is not None:  # live entries only (they come first inside an item); an item may keep none
        live_cum = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum((~dead).to(torch.int64), 0)])
        item_count = live_cum[item_end] - live_cum[item_start]
    type_of_rank = torch.empty(5, dtype=torch.int64, device=dev)
    type_of_rank[rank_of_type] = torch.arange(5, device=dev)
    item_key = key[item_start] >> 31  # chunk | sort rank of the type
    item_type = type_of_rank[item_key & 7]
    item_of_entry = torch.cumsum(first.to(torch.int64), 0) - 1
    r_in_item = torch.arange(n_entries, device=dev) - item_start[item_of_entry]
    if int(item_count.max().item()) > 255:
        raise ValueError("p2 item plan: a node belongs to more than 255 cells")
    face_items = (item_type == TYPE_C) | (item_type == TYPE_E)
    if bool((item_count[face_items] > 2).any()):
        raise ValueError("p2 item plan: a face belongs to more than two cells (non-manifold mesh)")
    # item headers: tile offsets of the blocks, pitch, opposite-edge blocks of the first / second cell
    it_row = rowN[item_start]
    it_blk = blk[item_start]
    it_chunk = item_key >> 3
    rb = bptr[it_row] - c_b0[it_chunk]
    where = torch.where(it_blk >= 0, 9 * rb[:, None] + 3 * (it_blk - bptr[it_row][:, None]), torch.full_like(it_blk, NONE16))
    fwhere = torch.full((n_items, 2), NONE16, dtype=torch.int64, device=dev)
    ef = torch.nonzero(fblk >= 0).reshape(-1)
    if int(ef.shape[0]):
        ei = item_of_entry[ef]
        fwhere[ei, r_in_item[ef]] = 9 * rb[ei] + 3 * (fblk[ef] - bptr[rowN[ef]])
    hdr6 = torch.cat([where, (3 * nb_row[it_row])[:, None], fwhere], dim=1)  # [n_items, 6]
    if int(hdr6.max().item()) > NONE16:
        raise ValueError("p2 item plan: tile offset overflow")

    # ---- order of the items (= lanes) inside a (chunk, type) group:
    #   "cell"  by the first cell they read: neighbouring lanes then read the same 128-byte geometry records in the same
    #           instruction, at the price of some spread in the loop counts;
    #   "count" by descending count: lanes of a task run equally long;
    # then, optionally, the bank-aware permutation inside the groups (lane_optimizer).
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
    if lane_optimizer is not None:
        group_ptr = torch.cat([gstart, torch.tensor([n_items], device=dev)])
        perm2 = lane_optimizer(group_ptr, hdr6[iperm])
        perm2 = torch.as_tensor(perm2, device=dev).to(torch.int64)
        iperm = iperm[perm2]
    gid = torch.cumsum(gfirst.to(torch.int64), 0) - 1
    rank_in_group = torch.arange(n_items, device=dev) - gstart[gid]
    # tasks: 32 items of one (chunk, type) group
    lane = rank_in_group % 32
    tfirst = gfirst | (lane == 0)
    task_of_item = torch.cumsum(tfirst.to(torch.int64), 0) - 1
    n_tasks = int(task_of_item[-1].item()) + 1
    cnt_sorted = item_count[iperm]
    kmax = torch.zeros(n_tasks, dtype=torch.int64, device=dev)
    kmax.scatter_reduce_(0, task_of_item, cnt_sorted, reduce="amax", include_self=True)
    kmax.clamp_(min=1)  # a task of items without live entries still owns one (empty) ELL row: the kernel reads entry 0
    task_items = torch.zeros(n_tasks, dtype=torch.int64, device=dev)
    task_items.scatter_add_(0, task_of_item, torch.ones_like(task_of_item))
    ell_base = torch.cumsum(kmax * 32, 0) - kmax * 32
    ell_size = int((kmax * 32).sum().item())
    task_first_item = torch.nonzero(tfirst).reshape(-1)
    task_group_key = grp[task_first_item]
    task_type = type_of_rank[task_group_key & 7]
    task_chunk = task_group_key >> 3
    # item-header slots: one per item, two for types D and E; every task starts on an even slot
    per_item = torch.where(task_type >= TYPE_D, 2, 1)
    task_slots = (task_items * per_item + 1) & ~1
    slot_base = torch.cumsum(task_slots, 0) - task_slots
    n_slots = int(task_slots.sum().item())
    item_slot = slot_base[task_of_item] + lane * per_item[task_of_item]
    hdr_sorted = hdr6[iperm]
    item_hdr = torch.full((n_slots, 4), NONE16, dtype=torch.int64, device=dev)
    item_hdr[item_slot] = hdr_sorted[:, :4]
    is_e = item_type[iperm] == TYPE_E
    es = item_slot[is_e] + 1
    item_hdr[es, 0] = hdr_sorted[is_e, 4]
    item_hdr[es, 1] = hdr_sorted[is_e, 5]
    # row node of the items that own the right-hand-side rows: A in its unused block slots, D in its second slot
    row_sorted = it_row[iperm]
    is_a = item_type[iperm] == TYPE_A
    item_hdr[item_slot[is_a], 1] = row_sorted[is_a] & 0xFFFF
    item_hdr[item_slot[is_a], 2] = row_sorted[is_a] >> 16
    is_d = item_type[iperm] == TYPE_D
    ds = item_slot[is_d] + 1
    item_hdr[ds, 0] = row_sorted[is_d] & 0xFFFF
    item_hdr[ds, 1] = row_sorted[is_d] >> 16
    # ELL fill
    inv = torch.empty_like(iperm)
    inv[iperm] = torch.arange(n_items, device=dev)
    pos = inv[item_of_entry]
    ell = torch.full((ell_size,), NONE32, dtype=torch.int64, device=dev)
    if dead is None:
        ell[ell_base[task_of_item[pos]] + r_in_item * 32 + lane[pos]] = code
    else:
        lv = torch.nonzero(~dead).reshape(-1)
        ell[ell_base[task_of_item[pos[lv]]] + r_in_item[lv] * 32 + lane[pos[lv]]] = code[lv]
        n_entries = int(lv.shape[0])
    # chunk headers
    tfirst_chunk = torch.ones(n_tasks, dtype=torch.bool, device=dev)
    tfirst_chunk[1:] = task_chunk[1:] != task_chunk[:-1]
    task0 = torch.nonzero(tfirst_chunk).reshape(-1)
    if int(task0.shape[0]) != n_chunks:
        raise ValueError("p2 item plan: a chunk without tasks")
    ntask = torch.cat([task0[1:], torch.tensor([n_tasks], device=dev)]) - task0
    chunk_hdr = torch.stack([c_b0, c_b1 - c_b0, task0, ntask], dim=1)
    # bit 16 of word 0: parity of the chunk's first block (the tile sits on the same 16-byte phase as its image)
    task_hdr = torch.stack(
        [task_type | (kmax << 8) | ((c_b0[task_chunk] & 1) << 16), ell_base, slot_base, task_items | (task_chunk << 6)], dim=1)
    return {
        "chunk_hdr": chunk_hdr.to(torch.int32).contiguous(),
        "task_hdr": task_hdr.to(torch.int32).contiguous(),
        "item_hdr": item_hdr.to(torch.int32).to(torch.uint16 if hasattr(torch, "uint16") else torch.int16).contiguous(),
        "ell": (ell & NONE32).to(torch.uint32 if hasattr(torch, "uint32") else torch.int64).contiguous(),
        "n_chunks": n_chunks, "n_tasks": n_tasks, "n_items": n_items, "n_slots": n_slots, "ell_size": ell_size,
        "max_tile_blocks": int((c_b1 - c_b0).max().item()), "n_entries": n_entries,
        "all_rows_nonempty": int(bool(nonempty.all().item())),
    }


def balanced_split(plan, n_cta, task_cost=3):
    """Chunk range of each of ``n_cta`` CTAs, ``int32 [n_cta + 1]``: equal WORK instead of equal chunk counts.  A chunk costs
    ``sum over its tasks of (task_cost + kmax)`` -- a fixed part per task (headers, tile hand-over) plus one memory round per
    ELL row; chunks hold equal numbers of CSR blocks, but the rows next to the domain boundary have fewer cells to walk."""
    ch, th = plan["chunk_hdr"].to(torch.int64), plan["task_hdr"].to(torch.int64)
    w = task_cost + ((th[:, 0] >> 8) & 0xFF)
    cs = torch.cat([torch.zeros(1, dtype=torch.int64, device=w.device), torch.cumsum(w, 0)])
    cw = cs[ch[:, 2] + ch[:, 3]] - cs[ch[:, 2]]
    ccs = torch.cat([torch.zeros(1, dtype=torch.int64, device=w.device), torch.cumsum(cw, 0)])
    target = (ccs[-1].double() * torch.arange(n_cta + 1, device=w.device, dtype=torch.float64) / n_cta)
    split = torch.searchsorted(ccs.double(), target, right=False)
    split[0], split[-1] = 0, int(ch.shape[0])
    split = torch.cummax(split, 0).values  # non-decreasing whatever the rounding did
    return split.to(torch.int32).contiguous()


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
    """The 3x3 sums of one item: its three blocks and, for type E, the opposite-edge block of its first / second cell
    (unused ones stay zero).  ``entries``: packed codes in ascending cell order."""
    K = np.zeros((5, 3, 3))
    for n, e in enumerate(entries):
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
        else:
            s = gi + gj
            _add(K[0], rho, s, gx, -1.0 / 30.0)
            _add(K[1], rho, s, gi + gx, 2.0 / 15.0)
            _add(K[1], rho, gj, gx, 2.0 / 15.0)
            _add(K[2], rho, s, gj + gx, 2.0 / 15.0)
            _add(K[2], rho, gi, gx, 2.0 / 15.0)
            _add(K[3 + n], rho, s, s, -2.0 / 15.0)  # opposite edge of THIS cell: g_x + g_y = -(g_i + g_j)
    return K


N_BLOCKS_OF_TYPE = {TYPE_A: 1, TYPE_B: 2, TYPE_C: 1, TYPE_D: 3, TYPE_E: 3}
EDGE_OF_PAIR = {frozenset(pr): 4 + n for n, pr in enumerate(EDGE_PAIRS)}


def walk(plan, rec, n_bnz, re4=None, n_nodes=None):
    """CSR block values ``[n_bnz * 9]`` in the scalar CSR order of the fused path (row pitch = 3 * blocks in row).
    With ``re4`` ``[n_cells, 10, 4]`` (padded local load vectors) also the right-hand side ``[3 n_nodes]`` the items A
    and D write (``nfa_assemble_p2tet_elasticity_load``): returns ``(vals, rhs)``."""
    chunk_hdr = plan["chunk_hdr"].cpu().numpy().astype(np.int64)
    task_hdr = plan["task_hdr"].cpu().numpy().astype(np.int64)
    item_hdr = plan["item_hdr"].cpu().numpy().astype(np.int64) & 0xFFFF
    ell = plan["ell"].cpu().numpy().astype(np.int64) & NONE32
    vals = np.full(n_bnz * 9, np.nan)
    rhs = np.full(3 * n_nodes, np.nan) if re4 is not None else None
    for ci, (b0, nblk, task0, ntask) in enumerate(chunk_hdr):
        tile = np.full(nblk * 9, np.nan)
        for t in range(task0, task0 + ntask):
            typ, kmax = task_hdr[t, 0] & 0xFF, (task_hdr[t, 0] >> 8) & 0xFF
            assert ((task_hdr[t, 0] >> 16) & 1) == (b0 & 1)
            base, slot0, nit = task_hdr[t, 1], task_hdr[t, 2], task_hdr[t, 3] & 63
            assert task_hdr[t, 3] >> 6 == ci and slot0 % 2 == 0
            per = 2 if typ >= TYPE_D else 1
            for lane in range(nit):
                entries = [ell[base + 32 * k + lane] for k in range(kmax)]
                entries = [e for e in entries if e != NONE32]
                assert typ not in (TYPE_C, TYPE_E) or len(entries) <= 2
                K = item_blocks(typ, entries, rec)
                w0, w1, w2, pitch = item_hdr[slot0 + per * lane]
                targets = [(0, w0), (1, w1), (2, w2)][: N_BLOCKS_OF_TYPE[int(typ)]]
                if typ == TYPE_E:
                    f0, f1 = item_hdr[slot0 + per * lane + 1][:2]
                    # (f1 is set without a second entry when that cell is a dead one: its block is written as zeros)
                    targets += [(3, f0), (4, f1)]
                if typ in (TYPE_A, TYPE_D):  # the owner of the node's right-hand-side rows
                    node = (w1 | (w2 << 16)) if typ == TYPE_A else int(
                        item_hdr[slot0 + per * lane + 1][0] | (item_hdr[slot0 + per * lane + 1][1] << 16))
                    if rhs is not None:
                        f = np.zeros(3)
                        for e in entries:
                            cell, i, j = int(e) >> 6, (int(e) >> 4) & 3, (int(e) >> 2) & 3
                            a = i if typ == TYPE_A else EDGE_OF_PAIR[frozenset((i, j))]
                            f += re4[cell, a, :3]
                        assert np.all(np.isnan(rhs[3 * node : 3 * node + 3])), "right-hand-side rows written twice"
                        rhs[3 * node : 3 * node + 3] = f
                for k, w in targets:
                    if w == NONE16:
                        continue
                    for r in range(3):
                        assert np.all(np.isnan(tile[w + r * pitch : w + r * pitch + 3])), "block written twice"
                        tile[w + r * pitch : w + r * pitch + 3] = K[k][r]
        vals[9 * b0 : 9 * (b0 + nblk)] = tile
    return vals if rhs is None else (vals, rhs)
