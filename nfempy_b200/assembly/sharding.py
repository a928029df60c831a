"""Multi-GPU assembly: cells partitioned into contiguous ranges, one rank per GPU.

The reference is single-process; this is the N-GPU extension the north star asks for.  Design:

* rank r owns the cells ``[bounds[r], bounds[r+1])``; a dof (matrix row) is owned by the rank of the FIRST cell
  touching it -- consistent with the reference's first-appearance numbering (R/src/mesh/mesh_topology.py:28-35),
  so a rank's owned rows are contiguous intervals inside each entity block of the numbering
  (R/src/spaces/dof_map.py:74-84);
* the rank's local plan covers its own cells plus every other cell touching one of its rows ("halo" cells, a
  single layer), with the halo cells' coefficients and load set to zero: the local CSR then has the complete global
  pattern for every row the rank touches, while the values only hold the rank's own cells' contributions;
* after the local (fused) assembly, rows owned by a lower rank carry partial sums: their CSR segments and
  right-hand-side entries are packed (``nfa_pack_rows``), sent to the owner (``ncclSend/ncclRecv`` through
  ``torch.distributed`` P2P ops) and added there in ascending sender rank (``nfa_unpack_add_rows``).  Both sides
  hold the same pattern for those rows, so the segments line up entry by entry;
* result: every rank holds the finished rows it owns, in global numbering -- a row-block CSR.

Exchange paths: ``enable_p2p(dist)`` switches the exchange to the library's peer-memory kernels (``nfa_push_rows`` stores
the rows straight into the owner's receive area over NVLink, ``nfa_pull_add_rows`` waits for the epoch flag and adds;
csrc/nfa_p2p.cu) -- pure stream-ordered device work, so ``capture_step()`` records local assembly + exchange in ONE CUDA
graph.  Without it (or when CUDA IPC is unavailable) the rows go through pack kernels + ``ncclSend/ncclRecv``.

``ShardPlan`` is pure host logic (numpy) on the replicated destination table; the exchange takes pack/unpack
callables so that the same code runs over NCCL on device buffers and over gloo on CPU tensors (tests).
"""
import os

import numpy as np

from ..mesh import _accel


class ShardPlan:
    def __init__(self, dest, n_dof, world, rank, bounds=None):
        dest = np.asarray(dest)
        n_cells, n_loc = dest.shape
        self.world, self.rank, self.n_dof = world, rank, int(n_dof)
        if bounds is None:
            bounds = np.linspace(0, n_cells, world + 1).astype(np.int64)
        self.bounds = np.asarray(bounds, dtype=np.int64)
        c0, c1 = int(self.bounds[rank]), int(self.bounds[rank + 1])
        cell_rank = np.searchsorted(self.bounds, np.arange(n_cells), side="right") - 1
        first_cell = np.full(self.n_dof, n_cells, dtype=np.int64)
        touched, first_pos = _accel.unique(dest.ravel(), return_index=True)  # cells ascend along the flattened table
        first_cell[touched] = first_pos // n_loc
        owner = np.full(self.n_dof, -1, dtype=np.int64)
        touched_any = first_cell < n_cells
        owner[touched_any] = cell_rank[first_cell[touched_any]]
        self.owner = owner
        self.own_cells = np.arange(c0, c1, dtype=np.int64)
        rows_touched = _accel.unique(dest[c0:c1].ravel())
        self.rows_touched = rows_touched
        mask = np.zeros(self.n_dof, dtype=bool)
        mask[rows_touched] = True
        self.local_cells = np.flatnonzero(mask[dest].any(axis=1)).astype(np.int64)  # own + halo, ascending
        self.is_own = (self.local_cells >= c0) & (self.local_cells < c1)
        self.owned_rows = rows_touched[owner[rows_touched] == rank]
        # rows I touch but a lower rank owns -> I send my partial sums there
        self.send_rows = {}
        for q in np.unique(owner[rows_touched]):
            q = int(q)
            if q != rank:
                if q > rank:
                    raise AssertionError("first-touch ownership: a row touched by rank %d is owned by %d" % (rank, q))
                self.send_rows[q] = rows_touched[owner[rows_touched] == q].astype(np.int32)
        # rows I own that higher ranks touch -> I receive their partial sums (same lists, computed locally)
        self.recv_rows = {}
        for q in range(rank + 1, world):
            rq = _accel.unique(dest[int(self.bounds[q]) : int(self.bounds[q + 1])].ravel())
            rq = rq[owner[rq] == rank]
            if len(rq):
                self.recv_rows[q] = rq.astype(np.int32)

    def row_block_ranges(self):
        """Owned rows as a list of maximal contiguous ``(start, stop)`` intervals."""
        r = self.owned_rows
        if len(r) == 0:
            return []
        cut = np.flatnonzero(np.diff(r) != 1) + 1
        starts = np.concatenate([[0], cut])
        stops = np.concatenate([cut, [len(r)]])
        return [(int(r[a]), int(r[b - 1]) + 1) for a, b in zip(starts, stops)]


def buffer_offsets(indptr, rows):
    """Prefix sums of the CSR row lengths of ``rows`` (numpy int64 ``[len(rows)+1]``)."""
    lengths = (indptr[rows.astype(np.int64) + 1] - indptr[rows.astype(np.int64)]).astype(np.int64)
    return np.concatenate([[0], np.cumsum(lengths)])


def exchange_interface_rows(dist, plan, make_buffer, pack, unpack_add, device_sync=None):
    """Sums the interface rows onto their owners.

    ``make_buffer(n)`` allocates a float64 buffer tensor; ``pack(rows, buf)`` fills it with [matrix segments | rhs
    entries] of ``rows``; ``unpack_add(rows, buf)`` adds a received buffer.  Buffer sizes are known on both sides
    from the shared pattern (``sizes[q]``), so no size handshake is needed."""
    ops, recv_bufs, keep = [], {}, []
    for q, rows in sorted(plan.send_rows.items()):
        buf = make_buffer(("send", q), rows)
        pack(rows, buf)
        keep.append(buf)
        ops.append(dist.P2POp(dist.isend, buf, q))
    for q, rows in sorted(plan.recv_rows.items()):
        buf = make_buffer(("recv", q), rows)
        recv_bufs[q] = (rows, buf)
        ops.append(dist.P2POp(dist.irecv, buf, q))
    if ops:
        if device_sync is not None:
            device_sync()
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for q in sorted(recv_bufs):  # fixed ascending-rank order => deterministic sums
        rows, buf = recv_bufs[q]
        unpack_add(rows, buf)


class ShardedAssembler:
    """One rank's share of a cell-partitioned assembly on its GPU.

    ``space`` is the (replicated) global ``ProductSpace``; ``form`` a fused-capable volume form on it.  ``assemble``
    returns ``(indptr, indices, values, rhs)`` device tensors in GLOBAL numbering whose owned rows
    (``shard.owned_rows`` / ``shard.row_block_ranges()``) are final."""

    def __init__(self, space, form, world, rank, device=None, bounds=None):
        import torch

        from ..device import space_device
        from .gpu_assembler import GPUAssembler

        self.space, self.form, self.world, self.rank = space, form, world, rank
        dest = space.destination_table()
        self.shard = ShardPlan(dest, space.n_dof, world, rank, bounds)
        self.dev = space_device(space, device)
        self.asm = GPUAssembler(space, None, None, device=device)
        self.asm.form_to_input_list = {"volume": ["form", self.shard.local_cells, form, None]}
        self.state = None
        self._bufs = {}
        self._graph = None
        self._step_graph = None
        self._p2p = None
        self.rhs = None
        self._torch = torch

    def prepare(self):
        torch = self._torch
        if os.environ.get("NFA_P2TET_PRUNE", "1") != "0":
            # the halo cells contribute zeros by construction: the row-item lists leave their entries out (the items stay and
            # write the zeros that complete the pattern), so the halo layers cost list space only, no record loads
            self.asm.p2_cell_live = self.shard.is_own
        dev, entries, datas, dests, plan, key = self.asm.prepare()
        if plan.fused_index is None:
            raise NotImplementedError("ShardedAssembler needs the fused path (affine H1 form, cell-wise coefficients)")
        data = datas[0]
        own = torch.as_tensor(self.shard.is_own.astype(np.float64)).to(dev.device)
        from .. import _lib

        lib = _lib.load()
        row_items = bool(data.get("p2tet_ok")) and bool(lib.nfa_bplan_has_p2_items(plan.bplan))
        if row_items and all(kind == "constant" for (_, _, _, kind) in data.coefs.values()):
            # row-item kernel with mesh-constant coefficients: the halo cells are switched off by the plan's cell mask and
            # lambda, mu stay single numbers, so every rank runs the constant-rho instantiation like a single GPU does
            self._mask = torch.as_tensor(self.shard.is_own.astype(np.uint8)).to(dev.device)
            _lib.check(lib.nfa_bplan_set_cell_mask(plan.bplan, _lib.ptr(self._mask), _lib.stream_ptr()), "nfa_bplan_set_cell_mask")
        else:
            coefs = {}
            for name, (arr, sc, sq, kind) in data.coefs.items():
                if kind == "constant":
                    coefs[name] = ((arr * own).contiguous(), 1, 0, "cell")  # halo cells contribute exact zeros
                else:
                    coefs[name] = ((arr * own).contiguous(), sc, sq, kind)
            data.coefs = coefs
        if data.f is not None:
            data.f = (data.f * own[:, None, None]).contiguous()
        self.state = (dev, entries, datas, dests, plan, key)
        self.plan = plan
        self._indptr_host = plan.indptr.cpu().numpy()
        self._rows_dev, self._off_dev, self._sizes = {}, {}, {}
        for kind, table in (("send", self.shard.send_rows), ("recv", self.shard.recv_rows)):
            for q, rows in table.items():
                off = buffer_offsets(self._indptr_host, rows)
                self._rows_dev[(kind, q)] = torch.as_tensor(rows).to(dev.device)
                self._off_dev[(kind, q)] = torch.as_tensor(off).to(dev.device)
                self._sizes[(kind, q)] = int(off[-1])
        return self

    # -- the three stages of one assembly ------------------------------------------------------------
    def local_assemble(self):
        """Fused assembly of the rank's own cells (halo cells contribute exact zeros)."""
        torch = self._torch
        dev, entries, datas, dests, plan, key = self.state
        if getattr(self, "rhs", None) is None:
            self.rhs = torch.zeros(plan.n_dof, dtype=torch.float64, device=dev.device)
        if self._graph is not None:
            self._graph.replay()  # zeroes rhs, overwrites every CSR value (GPUAssembler.capture)
        else:
            self.asm.jacobian.is_zero = True
            self.asm.run(self.state, self.rhs, fresh="rows")  # zeroes rhs where the kernels do not overwrite it
        self.vals = self.asm.jacobian.values_dev
        return self.vals, self.rhs

    def capture_local(self):
        """Captures the local assembly (the kernels between two interface exchanges) in a CUDA graph: the per-launch host
        work of the eager path (Python + ctypes, about as long as the device needs) then no longer gates the device.
        The exchange itself (pack kernels, NCCL send/recv, unpack kernels) stays eager."""
        torch = self._torch
        dev, entries, datas, dests, plan, key = self.state
        if getattr(self, "rhs", None) is None:
            self.rhs = torch.zeros(plan.n_dof, dtype=torch.float64, device=dev.device)
        self._graph = self.asm.capture(self.state, self.rhs, fresh="rows")
        self.vals = self.asm.jacobian.values_dev
        return self

    def _buffer(self, tag):
        if tag not in self._bufs:
            n = self._sizes[tag] + int(self._rows_dev[tag].shape[0])
            self._bufs[tag] = self._torch.empty(n, dtype=self._torch.float64, device=self.dev.device)
        return self._bufs[tag]

    def pack(self, q):
        """Send buffer for owner ``q``: [CSR segments of the rows q owns | their rhs entries]."""
        from .. import _lib

        tag = ("send", q)
        buf, rows = self._buffer(tag), self._rows_dev[tag]
        _lib.check(
            _lib.load().nfa_pack_rows(_lib.ptr(self.plan.indptr), _lib.ptr(self.vals), _lib.ptr(rows),
                                      _lib.ptr(self._off_dev[tag]), int(rows.shape[0]), _lib.ptr(buf), _lib.stream_ptr()),
            "nfa_pack_rows")
        buf[self._sizes[tag] :] = self.rhs.index_select(0, rows.long())
        return buf

    def unpack_add(self, q, buf):
        """Adds the partial sums received from rank ``q`` (call in ascending q)."""
        from .. import _lib

        tag = ("recv", q)
        rows = self._rows_dev[tag]
        _lib.check(
            _lib.load().nfa_unpack_add_rows(_lib.ptr(self.plan.indptr), _lib.ptr(self.vals), _lib.ptr(rows),
                                            _lib.ptr(self._off_dev[tag]), int(rows.shape[0]), _lib.ptr(buf),
                                            _lib.stream_ptr()),
            "nfa_unpack_add_rows")
        self.rhs.index_add_(0, rows.long(), buf[self._sizes[tag] :])  # rows are distinct: deterministic

    # -- peer-memory exchange (csrc/nfa_p2p.cu) ---------------------------------------------------------
    def _arena_layout(self):
        """Byte offsets inside this rank's receive arena: [flags u64[world] | acks u64[world] | payload of each sender]."""
        off = (2 * self.world * 8 + 255) & ~255
        table = {}
        for q in sorted(self.shard.recv_rows):
            table[q] = off
            n = self._sizes[("recv", q)] + int(self._rows_dev[("recv", q)].shape[0])
            off += (8 * n + 255) & ~255
        return table, max(off, 256)

    def _p2p_alloc(self):
        import ctypes

        from .. import _lib

        table, size = self._arena_layout()
        ptr, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        _lib.check(_lib.load().nfa_p2p_alloc(size, ctypes.byref(ptr), handle), "nfa_p2p_alloc")
        return {"base": int(ptr.value), "handle": handle.raw, "table": table, "size": size}

    def _p2p_wire(self, mine, peer_base, peer_table):
        """``peer_base[q]`` = address of rank q's arena in this process, ``peer_table[q]`` its payload offsets."""
        torch = self._torch
        w, r = self.world, self.rank
        self._p2p = {
            "mine": mine,
            # one (epoch, blocks-done) state pair per peer and direction, local memory
            "state": torch.zeros((2 * w, 2), dtype=torch.int64, device=self.dev.device),
            "push": {q: (peer_base[q] + peer_table[q][r], peer_base[q] + 8 * r, mine["base"] + 8 * (w + q))
                     for q in self.shard.send_rows},
            "pull": {q: (mine["base"] + mine["table"][q], mine["base"] + 8 * q, peer_base[q] + 8 * (w + r))
                     for q in self.shard.recv_rows},
        }

    def enable_p2p(self, dist):
        """Collective over the process group: every rank allocates its receive arena, the CUDA IPC handles are exchanged
        and each rank maps the arenas of the ranks it exchanges rows with.  Returns True when the peer-memory exchange is
        active on ALL ranks (else the NCCL path stays)."""
        import ctypes

        from .. import _lib

        lib = _lib.load()
        ok, mine, err = True, None, ""
        try:
            mine = self._p2p_alloc()
        except _lib.NfaError as exc:
            ok, err = False, str(exc)
        info = [None] * self.world
        dist.all_gather_object(info, (ok, mine["handle"] if ok else None, mine["table"] if ok else None))
        peer_base, peer_table = {}, {}
        if all(i[0] for i in info):
            for q in sorted(set(self.shard.send_rows) | set(self.shard.recv_rows)):
                ptr = ctypes.c_void_p()
                try:
                    _lib.check(lib.nfa_p2p_open(ctypes.create_string_buffer(info[q][1], 64), ctypes.byref(ptr)), "nfa_p2p_open")
                    peer_base[q], peer_table[q] = int(ptr.value), info[q][2]
                except _lib.NfaError as exc:
                    ok, err = False, str(exc)
                    break
        else:
            ok = False
        flags = [None] * self.world
        dist.all_gather_object(flags, ok)
        if not all(flags):
            self._p2p, self.p2p_error = None, err or "a peer could not map the exchange arenas"
            return False
        self._p2p_wire(mine, peer_base, peer_table)
        return True

    def exchange_p2p(self):
        """Pushes this rank's partial rows to their owners and adds the rows pushed to it (ascending sender rank)."""
        import ctypes

        from .. import _lib

        lib, p, stream = _lib.load(), self._p2p, _lib.stream_ptr()
        vp = ctypes.c_void_p
        for q in sorted(self.shard.send_rows):
            tag = ("send", q)
            rows, (buf, flag, ack) = self._rows_dev[tag], p["push"][q]
            _lib.check(lib.nfa_push_rows(_lib.ptr(self.plan.indptr), _lib.ptr(self.vals), _lib.ptr(self.rhs), _lib.ptr(rows),
                                         _lib.ptr(self._off_dev[tag]), int(rows.shape[0]), vp(buf), vp(flag), vp(ack),
                                         vp(p["state"][q].data_ptr()), stream), "nfa_push_rows")
        for q in sorted(self.shard.recv_rows):
            tag = ("recv", q)
            rows, (buf, flag, ack) = self._rows_dev[tag], p["pull"][q]
            _lib.check(lib.nfa_pull_add_rows(_lib.ptr(self.plan.indptr), _lib.ptr(self.vals), _lib.ptr(self.rhs),
                                             _lib.ptr(rows), _lib.ptr(self._off_dev[tag]), int(rows.shape[0]), vp(buf), vp(flag),
                                             vp(ack), vp(p["state"][self.world + q].data_ptr()), stream), "nfa_pull_add_rows")

    def capture_step(self):
        """One CUDA graph for the whole step: local fused assembly + peer-memory exchange (needs ``enable_p2p``)."""
        torch = self._torch
        if self._p2p is None:
            raise RuntimeError("ShardedAssembler.capture_step: the peer-memory exchange is not enabled")
        dev, entries, datas, dests, plan, key = self.state
        if getattr(self, "rhs", None) is None:
            self.rhs = torch.zeros(plan.n_dof, dtype=torch.float64, device=dev.device)
        self._graph = None

        def body():
            self.asm.jacobian.is_zero = True
            self.asm._capturing = True
            try:
                self.asm.run(self.state, self.rhs, fresh="rows")
            finally:
                self.asm._capturing = False
            self.vals = self.asm.jacobian.values_dev
            self.exchange_p2p()

        with torch.cuda.device(dev.device):
            body()  # binds the pattern; every rank runs the same number of real exchanges, so the epochs stay in step
            cur = torch.cuda.current_stream(dev.device)
            side = torch.cuda.Stream(device=dev.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                body()
            cur.wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                body()
        self._step_graph = graph
        self._graph_keep = self.asm._keep
        return self

    def assemble(self, dist=None):
        """Local fused assembly + (for world > 1) the NCCL interface exchange.  Returns device tensors
        ``(indptr, indices, values, rhs)``; the rank's owned rows are final."""
        if self._step_graph is not None:
            self._step_graph.replay()
            return self.plan.indptr, self.plan.indices, self.vals, self.rhs
        self.local_assemble()
        if self.world > 1 and self._p2p is not None:
            self.exchange_p2p()
        elif self.world > 1:
            if dist is None:
                import torch.distributed as dist
            exchange_interface_rows(
                dist, self.shard, lambda tag, rows: self._buffer(tag),
                lambda rows, buf: self.pack(self._peer(self.shard.send_rows, rows)),
                lambda rows, buf: self.unpack_add(self._peer(self.shard.recv_rows, rows), buf),
            )
        return self.plan.indptr, self.plan.indices, self.vals, self.rhs

    @staticmethod
    def _peer(table, rows):
        return next(q for q, v in table.items() if v is rows)


class HaloReplicatedAssembler:
    """One rank's share of a cell-partitioned assembly WITHOUT an exchange, for spaces in which a dof belongs to at most two
    cells -- H(div) facet dofs + cell-local dofs: BASELINE configs 3, 4 and 5.

    Rank r owns the cells ``[bounds[r], bounds[r+1])`` and assembles them TOGETHER with the cells that share a dof with
    them (one layer across a partition face, ``streamed.partition_cells``): every row whose lowest cell is the rank's own is
    then complete on the rank, so the ranks' row blocks are final without any data-path collective (the
    path shards by independent partitions; the halo cells are assembled twice instead of exchanged).  Boundary forms keep
    the items whose rows the rank owns (a boundary facet's dofs belong to one cell).  ``form_to_input_list`` is the
    GLOBAL list a ``SequentialAssembler`` would get (R/src/assembly/SequentialAssembler.py:13-86); volume forms must run
    over all cells.  ``assemble()`` returns ``(indptr, indices, values, rhs)`` in global numbering plus ``owned_rows``;
    only the owned rows are meaningful (on the row-owner path the others are not even computed).  The lowest-order forms go through the row-owner kernels (``hdiv_rows.py``)."""

    _VOLUME = ("form", "time_dependent_form")
    _BOUNDARY = ("bc_form", "time_dependent_bc_form")

    def __init__(self, space, world, rank, device=None):
        from .gpu_assembler import GPUAssembler
        from .streamed import partition_cells

        self.space, self.world, self.rank = space, int(world), int(rank)
        self.cells, self.owned_rows, self.cell_range = partition_cells(space.destination_table(), self.world, self.rank)
        self.asm = GPUAssembler(space, None, np.zeros(int(space.n_dof)), device=device)
        self._owned_mask = np.zeros(int(space.n_dof), dtype=bool)
        self._owned_mask[self.owned_rows] = True
        self.asm.rows_wanted = self._owned_mask  # the halo cells' other rows belong to other ranks: not computed here

    def set_forms(self, form_to_input_list):
        mine = {}
        for name, entry in form_to_input_list.items():
            kind, sequence, form = entry[0], entry[1], entry[2]
            n_total = form.n_items_total()
            if kind in self._VOLUME:
                if len(sequence) != n_total:
                    raise NotImplementedError("HaloReplicatedAssembler: volume forms must be listed over all cells")
                items = self.cells
            elif kind in self._BOUNDARY:
                seq = np.asarray(list(sequence), dtype=np.int64)
                rows = np.asarray(form.dest_table())[seq]
                keep = self._owned_mask[rows]
                if np.any(keep.any(axis=1) != keep.all(axis=1)):
                    raise NotImplementedError("HaloReplicatedAssembler: a boundary item straddles two ranks' rows")
                items = seq[keep[:, 0]]
            else:
                raise NotImplementedError("HaloReplicatedAssembler: form type %r" % (kind,))
            mine[name] = [kind, items, form] + list(entry[3:])
        self.asm.form_to_input_list = mine
        return self

    def assemble(self):
        indptr, indices, values, rhs = self.asm.assemble()
        return indptr, indices, values, rhs, self.owned_rows
