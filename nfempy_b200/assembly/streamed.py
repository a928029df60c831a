"""Streamed assembly with the ``SequentialAssembler`` "full" pattern (every pair of a cell's dofs, zeros kept) for forms
whose local matrices cannot all be resident -- BASELINE config 4 at its stated order: ``LCEScaledDualWeakForm`` on
BDM_3 x RT_3 x P_2 x P_2 (R/src/weak_forms/lce_scaled_dual_weak_form.py:445-699, driver
R/examples/cosserat_elasticity/mfe_cosserat_elasticity_example_3.py:150-243) has 348 local dofs, 0.97 MB of local matrix
per cell; the two-stage plan of ``GPUAssembler`` keeps all local matrices plus a sorted 32-bit id per local entry.

Here the cells are walked in chunks (local kernel of the form -> scatter into the resident CSR), and the CSR position of a
local entry comes from small rank tables instead of a global sort.  That works in spaces where **every dof belongs to at
most two cells** (H(div) facet dofs + cell-local dofs: all of config 3 / 4 / 5's spaces): a row's columns are the sorted
union of the dofs of its one or two cells, so

    pos(cell, a, b) = indptr[dest[cell, a]] + rank of dest[cell, b] in that union

with one ``uint16`` table per cell (cell-local rows) and per ordered (cell, partner) pair (rows shared across a facet):
3.5 KB per cell at 348 local dofs.  ``indptr`` is int64 -- 73 002 tetrahedra hold 8.7e9 stored entries (the reference's
PETSc build with 32-bit indices cannot address that matrix; the semantics are the same).  Sum order: an entry is stored by
the lower cell of its row and added to by the higher one, chunks in ascending cell order -- the reference's call order
(R/src/assembly/ScatterFormData.py:5-11), no atomics (``csrc/nfa_stream.cu``).

Right-hand side: the local vectors of all cells are small (n_loc doubles per cell) and go through the ordinary row lists
(``nfa_plan_create`` without matrix segments), together with the boundary forms' vectors.
"""
import ctypes
import time

import numpy as np
import torch

from .. import _lib
from ..device import space_device


def build_tables(dest, n_dof, pair_batch=32768):
    """Rank tables and CSR pattern sources from the destination table ``dest`` (int64 tensor ``[n_cells, n_loc]``, any
    device).  Returns a dict of tensors on the same device:

    ``indptr`` int64 [n_dof + 1]; ``tab_of`` int32 [n_cells, n_loc] (table of every local row: the cell's own table for a
    cell-local row, ``n_cells + pair`` for a shared one); ``tabs`` int16 [n_cells + n_pairs, n_loc] (read as uint16: rank |
    add-flag << 15); ``src`` int32 + ``src_off`` int64 [n_dof] (where the sorted columns of every row start in ``src``)."""
    dev = dest.device
    n_cells, n_loc = int(dest.shape[0]), int(dest.shape[1])
    if 2 * n_loc >= 32768:
        raise ValueError("streamed assembly: more than 16383 local dofs")
    flat = dest.reshape(-1)
    if int(flat.min().item()) < 0 or int(flat.max().item()) >= n_dof:
        raise ValueError("streamed assembly: destination index outside [0, n_dof)")
    order = torch.argsort(flat, stable=True)  # ascending dof, ascending (cell, a) inside a dof
    sd = flat[order]
    cell_of = torch.div(order, n_loc, rounding_mode="floor")
    first = torch.ones_like(sd, dtype=torch.bool)
    first[1:] = sd[1:] != sd[:-1]
    start = torch.nonzero(first).reshape(-1)
    count = torch.cat([start[1:], torch.tensor([sd.shape[0]], device=dev)]) - start
    if int(count.max().item()) > 2:
        raise NotImplementedError("streamed assembly needs spaces whose dofs belong to at most two cells "
                                  "(H(div) facet dofs + cell-local dofs); use GPUAssembler")
    dofs = sd[start]
    first_cell = torch.full((n_dof,), -1, dtype=torch.int64, device=dev)
    second_cell = torch.full((n_dof,), -1, dtype=torch.int64, device=dev)
    first_cell[dofs] = cell_of[start]
    two = count == 2
    second_cell[dofs[two]] = cell_of[start[two] + 1]
    if bool((first_cell[dofs[two]] == second_cell[dofs[two]]).any()):
        raise NotImplementedError("streamed assembly: a cell lists a dof twice")
    me = torch.arange(n_cells, device=dev, dtype=torch.int64)[:, None]
    fc, sc = first_cell[dest], second_cell[dest]
    partner = torch.where(sc < 0, torch.full_like(sc, -1), torch.where(fc == me, sc, fc))
    shared = partner >= 0
    pkey = me * n_cells + partner
    upk = torch.unique(pkey[shared])  # ordered (cell, partner) pairs, ascending
    n_pairs = int(upk.shape[0])
    pair_c = torch.div(upk, n_cells, rounding_mode="floor")
    pair_p = upk - pair_c * n_cells
    tab_of = torch.where(shared, n_cells + torch.searchsorted(upk, pkey.clamp(min=0)), me.expand(-1, n_loc))
    # cell-local rows: rank inside the cell's sorted dofs
    sorted_dest, sidx = torch.sort(dest, dim=1)
    if bool((sorted_dest[:, 1:] == sorted_dest[:, :-1]).any()):
        raise NotImplementedError("streamed assembly: a cell lists a dof twice")
    tabs = torch.empty((n_cells + n_pairs, n_loc), dtype=torch.int16, device=dev)
    ranks = torch.empty_like(dest)
    ranks.scatter_(1, sidx, torch.arange(n_loc, device=dev, dtype=torch.int64).expand(n_cells, -1))
    tabs[:n_cells] = ranks.to(torch.int16)
    # shared rows: rank inside the sorted union of both cells' dofs; the union itself is kept once per unordered pair
    lower = pair_c < pair_p
    n_lower = int(lower.sum().item())
    lower_index = torch.cumsum(lower.to(torch.int64), 0) - 1  # position of an ordered pair (c < p) among the lower ones
    ucols = torch.zeros((n_lower, 2 * n_loc), dtype=torch.int32, device=dev)
    usize = torch.zeros(n_pairs, dtype=torch.int64, device=dev)
    for b0 in range(0, n_pairs, pair_batch):
        b1 = min(n_pairs, b0 + pair_batch)
        pc, pp = pair_c[b0:b1], pair_p[b0:b1]
        cat = torch.cat([dest[pc], dest[pp]], dim=1)
        sv, si = torch.sort(cat, dim=1)
        new = torch.ones_like(sv, dtype=torch.bool)
        new[:, 1:] = sv[:, 1:] != sv[:, :-1]
        urank = torch.cumsum(new.to(torch.int64), 1) - 1
        usize[b0:b1] = urank[:, -1] + 1
        dup = torch.zeros_like(new)
        dup[:, 1:] |= ~new[:, 1:]
        dup[:, :-1] |= ~new[:, 1:]
        rank_orig = torch.empty_like(urank)
        rank_orig.scatter_(1, si, urank)
        dup_orig = torch.empty_like(dup)
        dup_orig.scatter_(1, si, dup)
        add = (pc > pp)[:, None] & dup_orig[:, :n_loc]  # the higher cell adds to what the lower cell stored
        word = rank_orig[:, :n_loc] | (add.to(torch.int64) << 15)
        tabs[n_cells + b0 : n_cells + b1] = torch.where(word >= 32768, word - 65536, word).to(torch.int16)
        low = torch.nonzero(lower[b0:b1]).reshape(-1)
        if int(low.shape[0]):
            rows = lower_index[b0:b1][low]
            tmp = torch.zeros((int(low.shape[0]), 2 * n_loc), dtype=torch.int32, device=dev)
            tmp.scatter_(1, urank[low], sv[low].to(torch.int32))  # equal dofs land on the same slot
            ucols[rows] = tmp
    # rows: number of columns and where their sorted columns start in ``src``
    ncols = torch.zeros(n_dof, dtype=torch.int64, device=dev)
    src_off = torch.zeros(n_dof, dtype=torch.int64, device=dev)
    single = dofs[~two]
    ncols[single] = n_loc
    src_off[single] = first_cell[single] * n_loc
    pd = dofs[two]
    if int(pd.shape[0]):
        pid = torch.searchsorted(upk, first_cell[pd] * n_cells + second_cell[pd])  # the ordered pair (lower, higher)
        ncols[pd] = usize[pid]
        src_off[pd] = n_cells * n_loc + lower_index[pid] * (2 * n_loc)
    indptr = torch.zeros(n_dof + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(ncols, 0)
    src = torch.cat([sorted_dest.to(torch.int32).reshape(-1), ucols.reshape(-1)])
    return {"indptr": indptr, "tab_of": tab_of.to(torch.int32).contiguous(), "tabs": tabs.contiguous(), "src": src.contiguous(),
            "src_off": src_off, "n_pairs": n_pairs, "nnz": int(indptr[-1].item())}


def partition_cells(dest, world, rank):
    """Contiguous cell ranges with replicated halo cells instead of an exchange: rank ``r`` owns the cells
    ``[bounds[r], bounds[r+1])`` and assembles them TOGETHER with the cells that share a dof with them (at most one layer
    across a partition face: a dof belongs to at most two cells), so that every row its own cells touch is complete on the
    rank -- no interface sums, no collective.  A row belongs to the rank of the lowest cell that touches it (first-touch,
    like ``assembly/sharding.py``).  Returns ``(cells to assemble (ascending), owned rows (ascending), (c0, c1))``."""
    dest = np.asarray(dest)
    n_cells = len(dest)
    bounds = [(n_cells * r) // world for r in range(world + 1)]
    c0, c1 = bounds[rank], bounds[rank + 1]
    n_dof = int(dest.max()) + 1
    mine = np.zeros(n_dof, dtype=bool)
    mine[dest[c0:c1].reshape(-1)] = True
    touching = np.flatnonzero(mine[dest].any(axis=1))  # own cells + their neighbours across shared dofs
    first_cell = np.full(n_dof, n_cells, dtype=np.int64)
    np.minimum.at(first_cell, dest[touching].reshape(-1), np.repeat(touching, dest.shape[1]))
    owned = np.flatnonzero(mine & (first_cell >= c0) & (first_cell < c1))
    return touching.astype(np.int64), owned, (c0, c1)


class StreamedCSR:
    """Device CSR with 64-bit row pointers.  ``rows(r0, r1)`` copies a row block to the host (scipy-ready triple)."""

    def __init__(self, n_dof, indptr, indices, values):
        self.n_dof, self.indptr, self.indices, self.values = n_dof, indptr, indices, values

    @property
    def nnz(self):
        return int(self.values.shape[0])

    def rows(self, r0, r1):
        a, b = int(self.indptr[r0].item()), int(self.indptr[r1].item())
        ip = (self.indptr[r0 : r1 + 1] - a).cpu().numpy()
        return ip, self.indices[a:b].cpu().numpy(), self.values[a:b].cpu().numpy()


class StreamedAssembler:
    """``form_to_input_list`` like ``SequentialAssembler`` (R/src/assembly/SequentialAssembler.py:13-86) with ONE volume
    ``form`` over all cells plus ``bc_form`` entries that carry no matrix (natural boundary conditions).  ``assemble()``
    returns ``(StreamedCSR, rhs device tensor)``; ``stats`` holds the times of the stages."""

    def __init__(self, space, device=None, chunk_cells=2048, cells=None):
        """``cells`` (ascending global cell ids, optional): assemble these cells only -- one rank's share of a partitioned
        mesh, see ``partition_cells``.  Rows and columns keep the global numbering."""
        self.space, self.device, self.chunk_cells = space, device, int(chunk_cells)
        self.cells = None if cells is None else np.ascontiguousarray(cells, dtype=np.int64)
        if self.cells is not None and len(self.cells) > 1 and not np.all(np.diff(self.cells) > 0):
            raise ValueError("StreamedAssembler: cells must be ascending and distinct")
        self.form_to_input_list = {}
        self.stats = {}
        self._tables = None

    def _entries(self):
        volume, bcs = None, []
        for name, entry in self.form_to_input_list.items():
            kind, items, form, alpha = entry
            if kind == "form":
                if volume is not None:
                    raise NotImplementedError("StreamedAssembler: one volume form")
                n = form.n_items_total()
                if items is not None and (len(items) != n or not np.array_equal(np.asarray(items), np.arange(n))):
                    raise NotImplementedError("StreamedAssembler: the volume form must run over all cells, in order")
                volume = (form, alpha)
            elif kind == "bc_form":
                bcs.append((items, form, alpha))
            else:
                raise NotImplementedError("StreamedAssembler: form type %r" % kind)
        if volume is None:
            raise ValueError("StreamedAssembler: no volume form listed")
        return volume, bcs

    def assemble(self):
        (form, alpha), bcs = self._entries()
        dev = space_device(self.space, self.device)
        with torch.cuda.device(dev.device):
            return self._assemble(dev, form, alpha, bcs)

    def _assemble(self, dev, form, alpha, bcs):
        lib, stream = _lib.load(), _lib.stream_ptr()
        n_dof = int(self.space.n_dof)
        if n_dof > 2**31 - 1:
            raise NotImplementedError("StreamedAssembler: more than 2^31 dofs")
        t0 = time.perf_counter()
        dest_host = np.asarray(form.dest_table(), dtype=np.int64)
        cell_ids = np.arange(len(dest_host), dtype=np.int64) if self.cells is None else self.cells
        dest64 = torch.as_tensor(np.ascontiguousarray(dest_host[cell_ids])).to(dev.device)
        del dest_host
        n_cells, n_loc = int(dest64.shape[0]), int(dest64.shape[1])
        if self._tables is None:
            self._tables = build_tables(dest64, n_dof)
        T = self._tables
        dest = dest64.to(torch.int32).contiguous()
        del dest64
        nnz = T["nnz"]
        indices = torch.empty(nnz, dtype=torch.int32, device=dev.device)
        _lib.check(lib.nfa_rows_copy_i32(n_dof, _lib.ptr(T["indptr"]), _lib.ptr(T["src_off"]), _lib.ptr(T["src"]),
                                         _lib.ptr(indices), stream), "nfa_rows_copy_i32")
        vals = torch.empty(nnz, dtype=torch.float64, device=dev.device)
        torch.cuda.synchronize(dev.device)
        t1 = time.perf_counter()
        alpha_dev = None
        if alpha is not None and np.any(np.asarray(alpha) != 0.0):
            alpha_dev = torch.as_tensor(np.ascontiguousarray(alpha, dtype=np.float64).reshape(-1)).to(dev.device)
        re_all = torch.empty((n_cells, n_loc), dtype=torch.float64, device=dev.device)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        t_prepare = ms_local = ms_scatter = 0.0
        for c0 in range(0, n_cells, self.chunk_cells):
            c1 = min(n_cells, c0 + self.chunk_cells)
            tp = time.perf_counter()
            data = form.prepare(dev, cell_ids[c0:c1])  # host callables over the chunk's points, upload
            torch.cuda.synchronize(dev.device)
            t_prepare += time.perf_counter() - tp
            dchunk = dest[c0:c1]
            ev[0].record()
            Ke, re = form.local_blocks(dev, data, alpha_dev, dchunk if alpha_dev is not None else None)
            ev[1].record()
            re_all[c0:c1] = re
            ev[2].record()
            for p in (0, 1):
                _lib.check(lib.nfa_stream_scatter(c1 - c0, n_loc, _lib.ptr(dchunk), _lib.ptr(T["tab_of"][c0:c1]),
                                                  _lib.ptr(T["tabs"]), _lib.ptr(T["indptr"]), _lib.ptr(Ke), _lib.ptr(vals), p,
                                                  stream), "nfa_stream_scatter")
            ev[3].record()
            torch.cuda.synchronize(dev.device)
            ms_local += ev[0].elapsed_time(ev[1])
            ms_scatter += ev[2].elapsed_time(ev[3])
            del Ke, re, data
        t2 = time.perf_counter()
        # right-hand side: the volume form's vectors + the boundary forms', through the ordinary row lists
        dests, res, keep = [dest], [re_all], []
        for items, bform, balpha in bcs:
            bdev = space_device(self.space, self.device)
            bitems = None if items is None else np.asarray(list(items), dtype=np.int64)
            bdata = bform.prepare(bdev, bitems)
            bdest = torch.as_tensor(np.ascontiguousarray(bform.dest_table(), dtype=np.int32)).to(dev.device)
            if bitems is not None:
                bdest = bdest.index_select(0, torch.as_tensor(bitems).to(dev.device)).contiguous()
            bal = None
            if balpha is not None and np.any(np.asarray(balpha) != 0.0):
                bal = alpha_dev if balpha is alpha and alpha_dev is not None else \
                    torch.as_tensor(np.ascontiguousarray(balpha, dtype=np.float64).reshape(-1)).to(dev.device)
            Kb, rb = bform.local_blocks(bdev, bdata, bal, bdest if bal is not None else None)
            if bool((Kb != 0.0).any()):
                raise NotImplementedError("StreamedAssembler: boundary form %s carries a matrix" % type(bform).__name__)
            dests.append(bdest)
            res.append(rb)
            keep += [bdata, Kb]
        nseg = len(dests)
        dest_ptrs = (ctypes.c_void_p * nseg)(*[d.data_ptr() for d in dests])
        n_items = (ctypes.c_int64 * nseg)(*[int(d.shape[0]) for d in dests])
        n_locs = (ctypes.c_int32 * nseg)(*[int(d.shape[1]) for d in dests])
        flags = (ctypes.c_int32 * nseg)(*([0] * nseg))  # right-hand-side rows only
        handle = ctypes.c_void_p()
        _lib.check(lib.nfa_plan_create(n_dof, nseg, dest_ptrs, n_items, n_locs, flags, None, ctypes.byref(handle), stream),
                   "nfa_plan_create")
        rhs = torch.empty(n_dof, dtype=torch.float64, device=dev.device)
        re_ptrs = (ctypes.c_void_p * nseg)(*[r.data_ptr() for r in res])
        try:
            _lib.check(lib.nfa_plan_reduce(handle, None, re_ptrs, 0, None, 0, _lib.ptr(rhs), stream), "nfa_plan_reduce")
            torch.cuda.synchronize(dev.device)
        finally:
            lib.nfa_plan_destroy(handle)
        t3 = time.perf_counter()
        self.stats = {
            "n_cells": n_cells, "n_loc": n_loc, "n_dof": n_dof, "nnz": nnz, "n_pairs": T["n_pairs"],
            "chunk_cells": self.chunk_cells, "n_chunks": (n_cells + self.chunk_cells - 1) // self.chunk_cells,
            "tables_and_pattern_s": t1 - t0, "host_coefficients_s": t_prepare, "local_kernel_ms": ms_local,
            "scatter_ms": ms_scatter, "device_ms": ms_local + ms_scatter, "chunk_loop_wall_s": t2 - t1, "rhs_s": t3 - t2,
            "table_bytes": int(T["tabs"].numel() * 2 + T["tab_of"].numel() * 4),
            "csr_bytes": int(nnz * 12 + (n_dof + 1) * 8),
        }
        return StreamedCSR(n_dof, T["indptr"], indices, vals), rhs
