"""``GPUAssembler``: drop-in for the reference's ``SequentialAssembler`` (R/src/assembly/SequentialAssembler.py:13-86).

Same constructor ``(space, jacobian, residual)``, same ``form_to_input_list`` dictionary
``{name: [form_type, sequence, weak_form, alpha]}`` and same ``scatter_forms(measure_time_q=False)``; the
per-cell Python loop + ``Mat.setValue`` calls (R/src/assembly/ScatterFormData.py:5-56) are replaced by

1. one sparsity plan per form list (built once, reused across Newton / time steps),
2. one local-kernel launch per form (or the fused kernel for affine H1 forms with cell-wise coefficients),
3. one deterministic reduce into the CSR values and the right-hand side.

``residual`` stays a numpy vector mutated in place (``+=``), ``jacobian`` is a ``CSRMatrix`` (PETSc ``Mat``
shim).  ``assemble()`` is the one-shot convenience returning ``(indptr, indices, data, rhs)``.
"""
import ctypes
import os
import time

import numpy as np
import torch

from .. import _host, _lib
from ..device import space_device
from ..weak_forms.weak_form import BatchedWeakForm
from . import hdiv_rows
from .csr_matrix import CSRMatrix

_SUPPORTED = ("form", "bc_form", "time_dependent_form", "time_dependent_bc_form", "interface_form")
_KNOWN = (
    "form", "bc_form", "interface_form", "bc_interface_form", "time_dependent_form", "time_dependent_bc_form",
)


def _as_items(sequence, n_total):
    """None when the sequence is exactly ``range(n_total)``, else an int64 index array."""
    if isinstance(sequence, range) and sequence == range(n_total):
        return None
    items = np.asarray(list(sequence) if not isinstance(sequence, np.ndarray) else sequence, dtype=np.int64)
    if items.ndim != 1:
        raise ValueError("GPUAssembler: sequence must be a flat list of element indices")
    if len(items) == n_total and np.array_equal(items, np.arange(n_total)):
        return None
    if len(items) and (items.min() < 0 or items.max() >= n_total):
        raise IndexError("GPUAssembler: element index out of range")
    return items


_CHUNK = 1 << 19  # doubles per host-side work item (4 MB)


def _chunks(n):
    """Host passes over long vectors (state upload staging, ``residual += rhs``) are memory-bound numpy calls that release
    the GIL: cut into chunks and run on a few threads (``_host.pool``) they cost a fraction of one single-threaded sweep."""
    pool = _host.pool() if n >= 2 * _CHUNK else None
    if pool is None:
        return None, [(0, n)]
    return pool, _host.parts(n, _CHUNK)


def _add_into(dst, src):
    """``dst += src`` for 1-D float64 numpy arrays, chunked over the host threads."""
    pool, parts = _chunks(len(dst))
    if pool is None:
        dst += src
        return
    list(pool.map(lambda ab: np.add(dst[ab[0]:ab[1]], src[ab[0]:ab[1]], out=dst[ab[0]:ab[1]]), parts))


class SpaceGroup:
    """The spaces of a mixed-dimensional problem stacked into one system the way the reference drivers do by hand
    (R/examples/md_elliptic_scalar/md_elliptic_scalar.py:205-212): space k occupies the dofs
    ``[dof_shift_k, dof_shift_k + n_dof_k)``; the caller sets the shifts (``ProductSpace.dof_shift``)."""

    def __init__(self, spaces):
        self.spaces = list(spaces)
        if not self.spaces:
            raise ValueError("SpaceGroup: no spaces")
        self.discrete_spaces = {}  # no single field list: forms address their own space
        self.dof_shift = 0

    @property
    def n_dof(self):
        return int(max(s.dof_shift + s.n_dof for s in self.spaces))


def _primary_space(form):
    if hasattr(form, "primary_space"):
        return form.primary_space()
    return form.space


class _RowsUnavailable(Exception):
    """The row-owner path does not apply to this form list (a dof with more than two items, rows outside the direct
    forms' ranges): the list takes the generic two-stage plan."""


class _Plan:
    """Owns the native plan handles of one form list."""

    def __init__(self):
        self.rows = None  # row-owner path: {segment index: tables}, see _build_row_plan
        self.rows_cover_all = False
        self.plan = None  # nfa_plan* with gather lists
        self.pattern = None  # nfa_plan* adopting the block plan's pattern (fused mode)
        self.bplan = None
        self.fused_index = None
        self.rhs_gather = False  # the fused segment's right-hand side is summed by nfa_bplan_gather_load
        self.n_dof = 0
        self.nnz = 0
        self.indptr = None
        self.indices = None
        self.build_seconds = 0.0
        self.keep = []

    def __del__(self):
        try:
            lib = _lib.load()
            if self.plan:
                lib.nfa_plan_destroy(self.plan)
            if self.pattern:
                lib.nfa_plan_destroy(self.pattern)
            if self.bplan:
                lib.nfa_bplan_destroy(self.bplan)
        except Exception:
            pass


class GPUAssembler:
    def __init__(self, space, jacobian=None, residual=None, device=None, allow_fused=True, pattern_policy="full",
                 explicit_diagonal=False, cache_coefficients=True):
        """``pattern_policy``: "full" = every ``n_loc^2`` pair of every element, exact zeros kept (the library
        scatter, R/src/assembly/ScatterFormData.py:5-11); "filtered" = only entries with
        ``~np.isclose(value, 1e-16)``; "filtered_upper" = additionally ``row <= col`` (PETSc "sbaij" storage) -- the
        inline closures of the example drivers (R/examples/cauchy_elasticity/cg_elasticity_example_1.py:92-145,
        R/examples/cosserat_elasticity/mfe_cosserat_elasticity_example_3.py:150-243).  ``explicit_diagonal`` adds a
        stored ``(i, i, 0.0)`` for every local dof of the volume forms (``:110-111`` / ``:175-176``).

        ``cache_coefficients``: the reference evaluates ``form.functions[name](x, y, z)`` inside every ``evaluate_form``
        call; here the callables are evaluated over all quadrature points once per (form, sequence, functions, t) and the
        device arrays are kept across ``scatter_forms()`` calls -- a Newton / time loop written the reference's way
        (R/examples/geothermal_flow/three_fields_advection_diffusion.py:397-431) then only uploads the state vectors.
        The cache is keyed on the IDENTITY of the form, of its ``functions`` dictionary and of every callable in it
        (numbers by value) and on ``t``: rebind ``form.functions`` (or an entry) to change a coefficient, or call
        ``invalidate()``; closures that read mutable globals need ``cache_coefficients=False``."""
        if pattern_policy not in ("full", "filtered", "filtered_upper"):
            raise ValueError("GPUAssembler: unknown pattern policy %r" % (pattern_policy,))
        self.pattern_policy = pattern_policy
        self.explicit_diagonal = bool(explicit_diagonal)
        if isinstance(space, (list, tuple)):  # mixed-dimensional problem: one system over several product spaces
            space = SpaceGroup(space)
        self.space = space
        self.jacobian = jacobian if jacobian is not None else CSRMatrix(getattr(space, "n_dof", None))
        self.residual = residual
        self.device = device
        self.allow_fused = allow_fused
        self._form_to_input_list = None
        self._plans = {}
        self.last_timings = {}
        self.cache_coefficients = bool(cache_coefficients)
        self._entry_cache = []  # [(signature, FormData, dest)] most recent last
        self._rhs_dev = None
        self._rhs_host = None
        self._state_bufs = []
        self._state_stage = []  # pinned staging buffers of the state uploads, one per slot
        self._state_event = []
        self._capturing = False
        self._rows_refused = None
        self._filtered_plan = None  # plan of the last value-dependent pattern (reused while the masks do not change)
        # row-owner path only: bool mask [n_dof] of the rows the caller wants (a rank's owned rows); the other rows the listed
        # cells feed are neither computed nor written
        self.rows_wanted = None

    def invalidate(self):
        """Drops the cached coefficient arrays (the sparsity plans stay)."""
        self._entry_cache = []

    @property
    def form_to_input_list(self):
        return self._form_to_input_list

    @form_to_input_list.setter
    def form_to_input_list(self, value):
        self._form_to_input_list = value

    # ------------------------------------------------------------------------------------------
    def _entries(self):
        if not self._form_to_input_list:
            raise ValueError("GPUAssembler: form_to_input_list is not set")
        out = []
        for name, input_list in self._form_to_input_list.items():
            form_type = input_list[0]
            if form_type not in _KNOWN:
                raise KeyError(form_type)
            if form_type not in _SUPPORTED:
                raise NotImplementedError("GPUAssembler: form type %r has no CUDA kernel yet" % form_type)
            _, sequence, form, alpha = input_list[:4]
            t = None
            if "time_dependent" in form_type:
                # [type, sequence, form, (alpha_n, alpha), t]; the reference hands alpha_n[dest] to the form as its
                # first (differentiated) vector and alpha[dest] as the second (ScatterFormData.py:28-33, 43-49)
                if len(input_list) < 5:
                    raise ValueError("GPUAssembler: time-dependent forms are listed as [type, sequence, form, (alpha_n, alpha), t]")
                t = input_list[4]
            if not isinstance(form, BatchedWeakForm):
                raise NotImplementedError(
                    "GPUAssembler: %s has no CUDA kernel (no CPU fallback is provided)" % type(form).__name__
                )
            if form.form_kind != form_type:
                raise ValueError("GPUAssembler: %s is a %r, listed as %r" % (type(form).__name__, form.form_kind, form_type))
            if form_type == "interface_form":
                # (c1_idx, c0_p_idx, c0_n_idx) triples -- or the library's (idx, (idx_p, idx_n)) nesting
                # (R/src/assembly/ScatterFormData.py:59-68) -- handed to the form, which owns the destination rule
                if not hasattr(form, "interface_dest"):
                    raise NotImplementedError("GPUAssembler: %s does not describe its interface destinations" % type(form).__name__)
                rows = sequence
                if not isinstance(sequence, np.ndarray):
                    try:  # flat triples: one conversion (a Python loop over 10^4 triples costs more than the kernels)
                        rows = np.asarray(sequence, dtype=np.int64)
                        if rows.ndim != 2 or rows.shape[1] != 3:
                            raise ValueError
                    except (ValueError, TypeError):  # the library's (idx, (idx_p, idx_n)) nesting
                        rows = [[int(e[0]), int(e[1][0]), int(e[1][1])] if np.ndim(e[1]) else [int(v) for v in e]
                                for e in sequence]
                items = np.ascontiguousarray(rows, dtype=np.int64).reshape(-1, 3)
            else:
                items = _as_items(sequence, form.n_items_total())
            out.append((name, form_type, items, form, (alpha, t) if t is not None else alpha))
        return out

    @staticmethod
    def _dest_dev(dev, form, items):
        if form.form_kind == "interface_form":
            return torch.as_tensor(np.ascontiguousarray(form.interface_dest(items), dtype=np.int32)).to(dev.device)
        full = dev.dest() if form.form_kind in ("form", "time_dependent_form") else dev.bc_dest()
        if items is None:
            return full
        idx = torch.as_tensor(items).to(full.device)
        return full.index_select(0, idx).contiguous()

    def _plan_key(self, entries, fused_index):
        key = [self.space.n_dof, self.space.dof_shift, fused_index]
        for name, kind, items, form, _ in entries:
            spaces = form.space if isinstance(form.space, (list, tuple)) else [form.space]
            key.append((kind, tuple((id(sp), sp.dof_shift) for sp in spaces), type(form).__name__,
                        None if items is None else items.tobytes()))
        return tuple(key)

    def _build_plan(self, dev, entries, dests, fused_index, direct=None, datas=None):
        t0 = time.perf_counter()
        lib = _lib.load()
        p = _Plan()
        p.n_dof = int(self.space.n_dof)
        p.fused_index = fused_index
        stream = _lib.stream_ptr()
        nseg = len(entries)
        dest_ptrs = (ctypes.c_void_p * nseg)(*[d.data_ptr() for d in dests])
        n_items = (ctypes.c_int64 * nseg)(*[int(d.shape[0]) for d in dests])
        n_loc = (ctypes.c_int32 * nseg)(*[int(d.shape[1]) for d in dests])
        handle = ctypes.c_void_p()
        if direct:
            self._build_row_plan(lib, dev, p, entries, dests, datas, direct, dest_ptrs, n_items, n_loc, stream,
                                 self.rows_wanted)
            torch.cuda.synchronize(dev.device)
            p.build_seconds = time.perf_counter() - t0
            return p
        if fused_index is None:
            _lib.check(
                lib.nfa_plan_create(p.n_dof, nseg, dest_ptrs, n_items, n_loc, None, None, ctypes.byref(handle), stream),
                "nfa_plan_create",
            )
            p.plan = handle
        else:
            form = entries[fused_index][3]
            s = form._field()
            nc = s.n_comp
            fd = dests[fused_index]
            node_dest = (fd[:, ::nc] // nc).contiguous()
            if p.n_dof % nc != 0:
                raise _lib.NfaError("fused path: n_dof is not a multiple of n_comp")
            # the block path reads dof = node * n_comp + component (R/src/spaces/dof_map.py:68-83: an entity's components
            # are consecutive); anything else goes through the two-stage path
            for c in range(nc):
                if not bool(torch.equal(fd[:, c::nc], fd[:, ::nc] + c)) or int(node_dest.min()) < 0 or \
                        int(node_dest.max()) >= p.n_dof // nc:
                    raise _lib.NfaError("fused path: destination table is not node-major / component-minor")
            bh = ctypes.c_void_p()
            _lib.check(
                lib.nfa_bplan_create(
                    p.n_dof // nc, int(node_dest.shape[1]), nc, _lib.ptr(node_dest), int(node_dest.shape[0]),
                    ctypes.byref(bh), stream,
                ),
                "nfa_bplan_create",
            )
            p.bplan = bh
            if int(node_dest.shape[1]) == 10 and nc == 3 and s.dimension == 3 and os.environ.get("NFA_P2TET", "1") != "0":
                self._attach_p2_items(lib, dev, p, node_dest, stream, getattr(self, "p2_cell_live", None))
            ph = ctypes.c_void_p()
            _lib.check(lib.nfa_plan_from_bplan(bh, ctypes.byref(ph), stream), "nfa_plan_from_bplan")
            p.pattern = ph
            # the fused segment carries no local matrices; when it is the first of the list (its rows then come first in the
            # reference's call order) its right-hand side is summed by row owner (nfa_bplan_gather_load) and its rows stay out
            # of the plan's generic row lists (flag bit 1)
            p.rhs_gather = fused_index == 0 and os.environ.get("NFA_RHS_GATHER", "1") != "0"
            flags = (ctypes.c_int32 * nseg)(*[(2 if p.rhs_gather else 0) if i == fused_index else 1 for i in range(nseg)])
            _lib.check(
                lib.nfa_plan_create(p.n_dof, nseg, dest_ptrs, n_items, n_loc, flags, ph, ctypes.byref(handle), stream),
                "nfa_plan_create",
            )
            p.plan = handle
        p.nnz = int(lib.nfa_plan_nnz(p.plan))
        p.indptr = torch.empty(p.n_dof + 1, dtype=torch.int32, device=dev.device)
        p.indices = torch.empty(p.nnz, dtype=torch.int32, device=dev.device)
        _lib.check(lib.nfa_plan_csr(p.plan, _lib.ptr(p.indptr), _lib.ptr(p.indices), stream), "nfa_plan_csr")
        torch.cuda.synchronize(dev.device)
        p.build_seconds = time.perf_counter() - t0
        return p

    @staticmethod
    def _build_row_plan(lib, dev, p, entries, dests, datas, direct, dest_ptrs, n_items, n_loc, stream, rows_wanted=None):
        """Plan of a form list whose volume forms ``direct`` are assembled by row owner: (1) the CSR of the WHOLE list
        (every segment's pairs, so boundary / coupling forms may add columns), (2) per direct segment the rows' sources and
        the CSR slot of every local column (``nfa_rows_locate``), (3) gather lists of the other segments merged into the
        CSR.  The direct segments' row ranges must not overlap: every value of the rows they feed is then written by exactly
        one row thread before the other segments add to it; rows they do not feed are never visited."""
        nseg = len(entries)
        # row sources first (cheap, and the test that decides whether the path applies at all)
        rows, ranges, n_active = {}, [], 0
        for i in direct:
            found = hdiv_rows.row_sources(dests[i])
            if found is None:
                raise _RowsUnavailable()
            rlist, src = found
            if rows_wanted is not None:  # a rank's share: rows it does not own are left alone (their cells' halo layer)
                keep = torch.as_tensor(rows_wanted).to(rlist.device)[rlist.to(torch.int64)]
                rlist, src = rlist[keep].contiguous(), src[keep].contiguous()
                if int(rlist.shape[0]) == 0:
                    raise _RowsUnavailable()
            rows[i] = {"rows": rlist, "src": src, "blocks": hdiv_rows.row_blocks(rlist)}
            ranges.append((int(rlist[0].item()), int(rlist[-1].item()) + 1))
            n_active += int(rlist.shape[0])
        ranges.sort()
        if any(a[1] > b[0] for a, b in zip(ranges, ranges[1:])):  # a row fed by two direct forms
            raise _RowsUnavailable()
        # rows the direct forms do not feed (a sequence over part of the cells: one rank's share of a partitioned mesh) are
        # fed by the other segments only, which ADD to the zeroed values
        p.rows_cover_all = n_active == p.n_dof
        h0 = ctypes.c_void_p()
        flags = (ctypes.c_int32 * nseg)(*([3] * nseg))  # matrix pairs of every segment, no right-hand-side lists
        _lib.check(lib.nfa_plan_create(p.n_dof, nseg, dest_ptrs, n_items, n_loc, flags, None, ctypes.byref(h0), stream),
                   "nfa_plan_create")
        try:
            p.nnz = int(lib.nfa_plan_nnz(h0))
            p.indptr = torch.empty(p.n_dof + 1, dtype=torch.int32, device=dev.device)
            p.indices = torch.empty(p.nnz, dtype=torch.int32, device=dev.device)
            _lib.check(lib.nfa_plan_csr(h0, _lib.ptr(p.indptr), _lib.ptr(p.indices), stream), "nfa_plan_csr")
            torch.cuda.synchronize(dev.device)
        finally:
            lib.nfa_plan_destroy(h0)
        row_len = p.indptr[1:] - p.indptr[:-1]
        for i in direct:
            t = rows[i]
            form, data = entries[i][3], datas[i]
            spec = form.row_spec(data)
            nl = int(dests[i].shape[1])
            t["max_len"] = int(row_len[t["rows"].to(torch.int64)].max().item())
            runs = hdiv_rows.row_runs(spec["emap"], spec["ridx"])
            if t["max_len"] > 256 or t["max_len"] * 8 * 256 > 200 * 1024 or runs is None:
                raise _RowsUnavailable()
            t["roff"] = torch.as_tensor(runs[0]).to(dev.device)
            t["codes"] = torch.as_tensor(runs[1]).to(dev.device)
            n_rows = int(t["rows"].shape[0])
            t["slots"] = torch.empty((n_rows, 2, hdiv_rows.RUN), dtype=torch.uint8, device=dev.device)
            _lib.check(lib.nfa_rows_locate(n_rows, _lib.ptr(t["rows"]), _lib.ptr(p.indptr), _lib.ptr(p.indices),
                                           _lib.ptr(t["src"]), nl, _lib.ptr(t["codes"]), _lib.ptr(dests[i]), _lib.ptr(t["slots"]),
                                           stream), "nfa_rows_locate")
            t["stride"], t["n_loc"] = int(spec["stride"]), nl
        p.rows = rows
        ph = ctypes.c_void_p()
        _lib.check(lib.nfa_plan_from_csr(p.n_dof, p.nnz, _lib.ptr(p.indptr), _lib.ptr(p.indices), ctypes.byref(ph)),
                   "nfa_plan_from_csr")
        p.pattern = ph
        flags = (ctypes.c_int32 * nseg)(*[2 if i in rows else 1 for i in range(nseg)])  # direct: no lists at all
        handle = ctypes.c_void_p()
        _lib.check(lib.nfa_plan_create(p.n_dof, nseg, dest_ptrs, n_items, n_loc, flags, ph, ctypes.byref(handle), stream),
                   "nfa_plan_create")
        p.plan = handle

    @staticmethod
    def _attach_p2_items(lib, dev, p, node_dest, stream, cell_live=None):
        """Row-item lists of the vector-P2 tetrahedron kernel (assembly/p2_items.py), built on the device with torch
        tensor operations from the plan's block pattern and copied into the plan."""
        from . import p2_items

        n_nodes = p.n_dof // 3
        n_blocks = int(lib.nfa_bplan_n_blocks(p.bplan))
        bptr = torch.empty(n_nodes + 1, dtype=torch.int32, device=dev.device)
        bcols = torch.empty(n_blocks, dtype=torch.int32, device=dev.device)
        _lib.check(lib.nfa_bplan_block_pattern(p.bplan, _lib.ptr(bptr), _lib.ptr(bcols), stream), "nfa_bplan_block_pattern")
        cap = int(os.environ.get("NFA_P2TET_CAP", "640"))
        lane_optimizer = None
        if os.environ.get("NFA_P2TET_LANEOPT", "0") != "0":  # measured: costs the lanes' cell-order locality more than it saves
            window = int(os.environ.get("NFA_P2TET_LANEWIN", "48"))

            def lane_optimizer(group_ptr, hdr6):
                # bank-aware lane order inside the (chunk, type) groups: one device thread per group (nfa_p2_lane_order)
                gp = group_ptr.to(torch.int64).contiguous()
                h = hdr6.to(torch.int32).contiguous()
                perm = torch.empty(int(h.shape[0]), dtype=torch.int64, device=h.device)
                _lib.check(lib.nfa_p2_lane_order(int(gp.shape[0]) - 1, _lib.ptr(gp), _lib.ptr(h), window, _lib.ptr(perm), stream),
                           "nfa_p2_lane_order")
                return perm

        items = p2_items.build(node_dest.to(torch.int64), bptr.to(torch.int64), bcols.to(torch.int64), n_nodes, cap=cap,
                               riffle=int(os.environ.get("NFA_P2TET_RIFFLE", "148")),
                               batch=int(os.environ.get("NFA_P2TET_BATCH", "8")),
                               lane_order=os.environ.get("NFA_P2TET_LANES", "cell"), lane_optimizer=lane_optimizer,
                               locality=os.environ.get("NFA_P2TET_LOCAL", "1") != "0",
                               type_order=tuple(int(c) for c in os.environ.get("NFA_P2TET_TYPEORDER", "01234")),
                               cell_live=None if cell_live is None else torch.as_tensor(cell_live).to(dev.device))
        _lib.check(
            lib.nfa_bplan_set_p2_items(
                p.bplan, items["n_chunks"], _lib.ptr(items["chunk_hdr"]), items["n_tasks"], _lib.ptr(items["task_hdr"]),
                items["n_slots"], _lib.ptr(items["item_hdr"]), items["ell_size"], _lib.ptr(items["ell"]),
                items["max_tile_blocks"], stream,
            ),
            "nfa_bplan_set_p2_items",
        )
        p.p2_stats = {k: v for k, v in items.items() if isinstance(v, int)}
        if os.environ.get("NFA_P2TET_BALANCE", "1") != "0":  # one CTA per SM: work-balanced chunk ranges
            n_cta = int(torch.cuda.get_device_properties(dev.device).multi_processor_count)
            if items["n_chunks"] >= n_cta:
                split = p2_items.balanced_split(items, n_cta)
                _lib.check(lib.nfa_bplan_set_p2_split(p.bplan, n_cta, _lib.ptr(split), stream), "nfa_bplan_set_p2_split")

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _signature(kind, items, form, t):
        fns = form.functions
        vals = tuple(fns.values()) if isinstance(fns, dict) else ()
        return (kind, form, None if items is None else items.tobytes(), fns, vals, t)

    @staticmethod
    def _same_signature(a, b):
        if a[0] != b[0] or a[1] is not b[1] or a[2] != b[2] or a[3] is not b[3] or len(a[4]) != len(b[4]):
            return False
        for u, v in zip(a[4], b[4]):
            if u is v:
                continue
            if callable(u) or callable(v) or type(u) is not type(v):
                return False
            try:
                if not bool(np.all(u == v)):
                    return False
            except Exception:
                return False
        return a[5] == b[5]

    def _entry_data(self, dev, kind, items, form, t):
        """``(FormData, dest)`` of one entry: from the cache when form / sequence / functions / t are the ones seen
        before, else evaluated (``form.prepare``: host callables over all quadrature points, upload)."""
        dev = space_device(_primary_space(form), self.device)  # the form's own space (mixed-dimensional lists mix spaces)
        sig = self._signature(kind, items, form, t)
        previous = None
        if self.cache_coefficients:
            for k, (old, data, dest) in enumerate(self._entry_cache):
                if data.device_key == (dev.device, id(dev)) and self._same_signature(old, sig):
                    self._entry_cache.append(self._entry_cache.pop(k))
                    return data, dest
            for old, data, dest in reversed(self._entry_cache):  # same entry at another time: geometry / tables reusable
                if data.device_key == (dev.device, id(dev)) and self._same_signature(old[:5] + (t,), sig):
                    previous = (data, dest)
                    break
        if "time_dependent" in kind:
            data = form.prepare(dev, items, t=t, reuse=previous[0]) if previous and getattr(form, "supports_reuse", False) \
                else form.prepare(dev, items, t=t)
        else:
            data = form.prepare(dev, items)
        dest = previous[1] if previous else self._dest_dev(dev, form, items)
        data.device_key = (dev.device, id(dev))
        data.dev = dev
        data.alpha_dev = None
        data.alpha_zero = True
        if self.cache_coefficients:
            self._entry_cache.append((sig, data, dest))
            if len(self._entry_cache) > 16:
                self._entry_cache.pop(0)
        return data, dest

    def prepare(self):
        """Evaluates the coefficient callables (or fetches their cached device arrays), uploads the state vectors and
        builds (or fetches) the sparsity plan.  Returns the state ``run`` consumes."""
        first = self.space.spaces[0] if isinstance(self.space, SpaceGroup) else self.space
        dev = space_device(first, self.device)
        with torch.cuda.device(dev.device):  # libnfa launches on the current device
            return self._prepare(dev)

    def _prepare(self, dev):
        tp = [time.perf_counter()]
        entries = self._entries()
        tp.append(time.perf_counter())
        if isinstance(self.space, SpaceGroup):
            for (_, _, _, form, _) in entries:
                for sp in (form.space if isinstance(form.space, (list, tuple)) else [form.space]):
                    if not any(sp is known for known in self.space.spaces):
                        raise ValueError("GPUAssembler: %s works on a space outside the assembler's group" % type(form).__name__)
        datas, dests = [], []
        for (_, kind, items, form, al) in entries:
            data, dest = self._entry_data(dev, kind, items, form, al[1] if "time_dependent" in kind else None)
            datas.append(data)
            dests.append(dest)
        tp.append(time.perf_counter())
        # state vectors: one persistent device buffer per entry, refreshed in place on every call (a captured graph
        # reads the same buffers); ``alpha_zero`` lets eager runs skip the K.alpha terms of an all-zero vector
        # (the same numpy vector listed by several entries -- every form of a Newton iteration names the same iterate --
        # goes up once: slot k = the k-th distinct vector of the list, one persistent device buffer per slot)
        slots = {}

        def state(vec, want_zero_flag=True):
            """``want_zero_flag=False``: the caller launches its state terms whatever the vector holds (time-dependent
            forms), so the host scan for an all-zero vector (a full pass over n_dof doubles per call) is skipped."""
            if vec is None:
                return None, True
            if id(vec) not in slots:
                k = len(slots)
                while len(self._state_bufs) <= k:
                    self._state_bufs.append(None)
                buf, zero = self._upload(dev, vec, self._state_bufs[k], want_zero_flag, slot=k)
                self._state_bufs[k] = buf
                slots[id(vec)] = [buf, zero, want_zero_flag]
            elif want_zero_flag and not slots[id(vec)][2]:  # first listed by a time-dependent form: test now
                slots[id(vec)][1] = not bool(torch.any(slots[id(vec)][0]).item())
                slots[id(vec)][2] = True
            return slots[id(vec)][0], slots[id(vec)][1]

        for (_, kind, _, _, alpha), data in zip(entries, datas):
            if "time_dependent" in kind:
                (first, second), _t = alpha
                a0, z0 = state(first, False)
                a1, z1 = state(second, False) if kind == "time_dependent_form" else (None, True)
                data.alpha_dev, data.alpha_zero = (a0, a1), (z0 and z1)
            else:
                data.alpha_dev, data.alpha_zero = state(alpha)
        tp.append(time.perf_counter())
        self._prepare_split = {"entries_s": tp[1] - tp[0], "form_data_s": tp[2] - tp[1], "state_upload_s": tp[3] - tp[2]}
        fused_index = None
        if self.pattern_policy != "full":
            return dev, entries, datas, dests, None, None  # the pattern depends on the values: built in run()
        if self.allow_fused:
            for i, ((_, kind, _, form, _), data) in enumerate(zip(entries, datas)):
                if kind == "form" and form.fused_spec(data) is not None and len(self.space.discrete_spaces) == 1:
                    fused_index = i
                    break
        # lowest-order H(div) x L2 volume forms: records + row-owner expansion instead of local matrices (hdiv_rows.py)
        direct = None
        if fused_index is None and self.allow_fused and os.environ.get("NFA_HDIV_ROWS", "1") != "0" \
                and self._rows_refused != self._plan_key(entries, None):
            direct = tuple(i for i, ((_, kind, _, form, _), data) in enumerate(zip(entries, datas))
                           if kind in ("form", "time_dependent_form") and hasattr(form, "row_spec")
                           and form.row_spec(data) is not None) or None
        key = self._plan_key(entries, fused_index if direct is None else ("rows", id(self.rows_wanted)) + direct)
        if key not in self._plans:
            try:
                self._plans[key] = self._build_plan(dev, entries, dests, fused_index, direct, datas)
            except _RowsUnavailable:
                self._rows_refused = self._plan_key(entries, None)
                key = self._rows_refused
                if key not in self._plans:
                    self._plans[key] = self._build_plan(dev, entries, dests, None)
            except _lib.NfaError:
                if fused_index is None:
                    raise
                # e.g. a second volume form on other cells couples dofs outside the fused form's pattern: the reference
                # accepts such lists, so they take the generic two-stage plan instead of failing
                fused_index = None
                key = self._plan_key(entries, None)
                if key not in self._plans:
                    self._plans[key] = self._build_plan(dev, entries, dests, None)
        plan = self._plans[key]
        return dev, entries, datas, dests, plan, key

    def _upload(self, dev, vec, buf=None, want_zero_flag=True, slot=0):
        """``(device buffer, is_zero)`` of a state vector; ``None`` stays ``None`` (the form then has no state at all).
        The buffer is allocated once and refreshed in place: the host vector is copied chunk by chunk into a pinned
        staging buffer on a few threads, every chunk's DMA starting as soon as it is staged (a pageable ``cudaMemcpy`` of
        the 91 MB vectors of config 5 ran at 9 GB/s and was two thirds of a ``scatter_forms()`` call).  The all-zero test
        runs on the device copy; ``want_zero_flag=False`` skips it (``is_zero`` is then reported ``False``)."""
        if vec is None:
            return None, True
        host = np.ascontiguousarray(vec, dtype=np.float64).reshape(-1)
        n = len(host)
        if buf is None or buf.shape[0] != n:
            buf = torch.empty(n, dtype=torch.float64, device=dev.device)
        pool, parts = _chunks(n)
        if pool is None:
            buf.copy_(torch.from_numpy(host), non_blocking=False)
        else:
            while len(self._state_stage) <= slot:
                self._state_stage.append(None)
            stage = self._state_stage[slot]
            if stage is None or stage.shape[0] != n:
                stage = self._state_stage[slot] = torch.empty(n, dtype=torch.float64).pin_memory()
            while len(self._state_event) <= slot:
                self._state_event.append(None)
            if self._state_event[slot] is not None:
                self._state_event[slot].synchronize()  # the previous upload's DMA has left the staging buffer
            stage_np = stage.numpy()
            futures = [pool.submit(np.copyto, stage_np[a:b], host[a:b]) for a, b in parts]
            for (a, b), fut in zip(parts, futures):
                fut.result()
                buf[a:b].copy_(stage[a:b], non_blocking=True)  # stream-ordered before the kernels that read it
            self._state_event[slot] = torch.cuda.Event()
            self._state_event[slot].record()
        return buf, (not bool(torch.any(buf).item())) if want_zero_flag else False

    def run(self, state, rhs_dev=None, fresh=False):
        """Launches the kernels of one assembly.  Adds into ``self.jacobian`` and into ``rhs_dev`` (the device residual,
        returned).  ``fresh=True``: ``rhs_dev`` holds nothing the caller wants to keep -- the assembly result replaces it
        (the buffer is zeroed here unless the kernels overwrite every row anyway).  ``fresh="rows"``: the same for the rows of
        the nodes that have cells in this form list; every other row is the caller's (a rank's share of a sharded assembly)."""
        with torch.cuda.device(state[0].device):
            return self._run(state, rhs_dev, fresh)

    def _run(self, state, rhs_dev=None, fresh=False):
        dev, entries, datas, dests, plan, key = state
        lib = _lib.load()
        stream = _lib.stream_ptr()
        jac = self.jacobian
        if plan is None:
            if fresh and rhs_dev is not None:
                rhs_dev.zero_()
            return self._run_filtered(state, rhs_dev)
        if jac.plan_key != key:
            if jac.plan_key is not None and not jac.is_zero:
                raise _lib.NfaError("GPUAssembler: the jacobian holds values assembled on a different sparsity plan")
            jac.bind_pattern(key, plan.n_dof, plan.indptr, plan.indices)
        # Right-hand side of the fused form: summed by ROW OWNER from the cells' local load vectors
        # (nfa_bplan_gather_load, plan built with the segment's rows left out of the generic lists); the sums OVERWRITE the
        # rows of a fresh buffer -- no fill.  NFA_P2TET_RHS=1: the row-item matrix kernel gathers them itself (measured slower)
        gather = bool(getattr(plan, "rhs_gather", False))
        fold = gather and self._folds_load(lib, plan, entries, datas)
        fdata = datas[plan.fused_index] if gather else None
        has_load = gather and fdata.f is not None
        # tile-merged load (local load vectors never reach HBM) where the streaming load kernel exists
        merged = bool(has_load and not fold and fdata.n_q <= 64
                      and (fdata.dim, fdata.n_phi, fdata.n_comp) in ((3, 10, 3), (3, 4, 3), (2, 6, 2))
                      and os.environ.get("NFA_LOAD_MERGE", "1") != "0")
        # fresh == "rows": only the rows of nodes that have cells in this plan are replaced, the caller keeps the others (a
        # rank's plan over the global numbering must not sweep the other ranks' rows on every step)
        rows_only = isinstance(fresh, str) and fresh == "rows"
        fresh = bool(fresh) or rhs_dev is None
        if merged:
            all_active = int(lib.nfa_bplan_n_active_nodes(plan.bplan)) * fdata.n_comp == plan.n_dof
            overwrite = fresh and (all_active or rows_only)
        elif fold:
            overwrite = fresh and bool(plan.p2_stats.get("all_rows_nonempty"))
        else:
            overwrite = fresh and has_load  # nfa_bplan_gather_load writes every row (zeros for nodes without cells)
        if plan.rows:  # the row threads write every right-hand-side row of a fresh buffer themselves
            overwrite = fresh and (plan.rows_cover_all or rows_only)
        self.last_rhs_fill = fresh and not overwrite  # a torch fill kernel ran (launch accounting)
        if rhs_dev is None:
            rhs_dev = (torch.empty if (overwrite and not rows_only) else torch.zeros)(plan.n_dof, dtype=torch.float64, device=dev.device)
        elif fresh and not overwrite:
            rhs_dev.zero_()
        nseg = len(entries)
        Ke_list, re_list, keep = [None] * nseg, [None] * nseg, []
        if plan.rows:
            keep += self._run_rows(lib, dev, plan, entries, datas, dests, jac, rhs_dev, 0 if overwrite else 1, stream)
        for i, ((_, kind, _, form, alpha), data, dest) in enumerate(zip(entries, datas, dests)):
            if plan.rows and i in plan.rows:
                continue
            if "time_dependent" in kind:
                first_dev, second_dev = data.alpha_dev
                nl = form.n_loc()
                Ke = torch.empty((data.n_items, nl, nl), dtype=torch.float64, device=dev.device)
                re = torch.empty((data.n_items, nl), dtype=torch.float64, device=dev.device)
                form.launch_local(data.dev, data, first_dev, dest, Ke, re, second_dev)
                Ke_list[i], re_list[i] = Ke, re
                keep += [Ke, re, first_dev, second_dev]
                continue
            # an all-zero state vector skips the K.alpha terms in eager runs; a captured graph always carries them (the
            # replayed kernels read the persistent buffer, whatever it holds by then)
            alpha_dev = data.alpha_dev if (self._capturing or not data.alpha_zero) else None
            if i == plan.fused_index:
                spec = form.fused_spec(data)
                target = jac.values_dev if jac.is_zero else torch.empty_like(jac.values_dev)
                kernel = spec["kernel"]
                p2tet = bool(spec.get("p2tet_kernel") and lib.nfa_bplan_has_p2_items(plan.bplan))
                if p2tet:
                    kernel = spec["p2tet_kernel"]
                    args = [plan.bplan, _lib.ptr(dev.coords), _lib.ptr(data.cell_nodes)]
                else:
                    args = [plan.bplan, data.dim, _lib.ptr(form.mref(dev)), _lib.ptr(dev.coords), _lib.ptr(data.cell_nodes)]
                for c in spec["coefs"]:
                    args += [_lib.ptr(c[0]), c[1]]
                args += [_lib.ptr(target)]
                # right-hand side of the form's load: tile-merged (local load vectors never reach HBM) where the streaming
                # load kernel exists, else local load vectors [n, n_loc] + the row-owner gather (or the plan's generic lists)
                re = None
                if data.f is not None and not merged:
                    if fold:  # padded node records [n, 10, 4]: one 32-byte sector per cell and node
                        re = torch.empty((data.n_items, 10, 4), dtype=torch.float64, device=dev.device)
                        form.launch_load(data.dev, data, re, comp_pitch=4)
                    else:
                        re = torch.empty((data.n_items, form.n_loc()), dtype=torch.float64, device=dev.device)
                        form.launch_load(data.dev, data, re)
                    keep.append(re)
                if fold:
                    kernel += "_load"
                    args += [_lib.ptr(re), _lib.ptr(rhs_dev), 1 if overwrite else 0]
                args += [stream]
                _lib.check(getattr(lib, kernel)(*args), kernel)
                if merged:
                    _lib.check(lib.nfa_bplan_load_rhs(plan.bplan, data.dim, data.n_q, _lib.ptr(data.tab[0]), _lib.ptr(data.w),
                                                      _lib.ptr(dev.coords), _lib.ptr(data.cell_nodes), _lib.ptr(data.f),
                                                      float(form.load_sign), 1 if overwrite else 0, _lib.ptr(rhs_dev), stream),
                               "nfa_bplan_load_rhs")
                elif re is not None and gather and not fold:
                    _lib.check(lib.nfa_bplan_gather_load(plan.bplan, _lib.ptr(re), data.n_comp, 1 if overwrite else 0,
                                                         _lib.ptr(rhs_dev), stream), "nfa_bplan_gather_load")
                elif re is not None and not gather:
                    re_list[i] = re
                if alpha_dev is not None:
                    _lib.check(
                        lib.nfa_csr_spmv(plan.pattern, _lib.ptr(target), _lib.ptr(alpha_dev), 1, _lib.ptr(rhs_dev), stream),
                        "nfa_csr_spmv",
                    )
                if target is not jac.values_dev:
                    jac.values_dev.add_(target)
                jac.is_zero = False
            else:
                Ke, re = form.local_blocks(data.dev, data, alpha_dev, dest if alpha_dev is not None else None)
                Ke_list[i], re_list[i] = Ke, re
                keep += [Ke, re, alpha_dev]
        ke_ptrs = (ctypes.c_void_p * nseg)(*[None if k is None else k.data_ptr() for k in Ke_list])
        re_ptrs = (ctypes.c_void_p * nseg)(*[None if r is None else r.data_ptr() for r in re_list])
        accumulate_vals = 0 if (jac.is_zero and plan.fused_index is None and not plan.rows) else 1
        if any(k is not None for k in Ke_list) or any(r is not None for r in re_list):
            _lib.check(
                lib.nfa_plan_reduce(plan.plan, ke_ptrs, re_ptrs, accumulate_vals, _lib.ptr(jac.values_dev), 1, _lib.ptr(rhs_dev), stream),
                "nfa_plan_reduce",
            )
        jac.is_zero = False
        self._keep = keep  # buffers must outlive the stream-ordered kernels
        return rhs_dev

    def _run_rows(self, lib, dev, plan, entries, datas, dests, jac, rhs_dev, rhs_add, stream):
        """Row-owner segments (hdiv_rows.py): records kernel of the form, then one thread per matrix row.  The rows kernel
        OVERWRITES the CSR values of its rows; values assembled earlier are kept by expanding into a scratch array first."""
        target = jac.values_dev if jac.is_zero else \
            (torch.empty_like if plan.rows_cover_all else torch.zeros_like)(jac.values_dev)
        keep = []
        for i, t in plan.rows.items():
            (_, kind, _, form, _), data, dest = entries[i], datas[i], dests[i]
            if "time_dependent" in kind:
                first_dev, second_dev = data.alpha_dev
            else:
                first_dev = data.alpha_dev if (self._capturing or not data.alpha_zero) else None
                second_dev = None
            rec = torch.empty((data.n_items, t["stride"]), dtype=torch.float64, device=dev.device)
            form.launch_records(data.dev, data, first_dev, second_dev, dest, rec, t["stride"])
            _lib.check(lib.nfa_rows_expand(int(t["blocks"].shape[0]), _lib.ptr(t["blocks"]), t["max_len"], _lib.ptr(plan.indptr), _lib.ptr(t["src"]),
                                           _lib.ptr(t["slots"]), t["n_loc"], _lib.ptr(t["roff"]), _lib.ptr(t["codes"]),
                                           _lib.ptr(rec), t["stride"], _lib.ptr(target), _lib.ptr(rhs_dev), rhs_add, stream),
                       "nfa_rows_expand")
            keep += [rec, first_dev, second_dev]
        if target is not jac.values_dev:
            jac.values_dev.add_(target)
        jac.is_zero = False
        return keep

    @staticmethod
    def _folds_load(lib, plan, entries, datas):
        """True when the fused form runs on the row-item kernel and has a load: the kernel then also writes the form's
        right-hand side (``nfa_assemble_p2tet_elasticity_load``; ``NFA_P2TET_RHS=0`` keeps the separate gather)."""
        i = plan.fused_index
        # measured: the gather inside the matrix kernel costs more (0.77 -> 0.98 ms) than the separate pass it replaces, so it
        # is opt-in (profiles/r02_p2tet_variants.md, batch 9)
        if i is None or os.environ.get("NFA_P2TET_RHS", "0") == "0" or not getattr(plan, "p2_stats", None):
            return False
        form, data = entries[i][3], datas[i]
        spec = form.fused_spec(data)
        return bool(spec and spec.get("p2tet_kernel") and data.f is not None and lib.nfa_bplan_has_p2_items(plan.bplan)
                    and hasattr(lib, spec["p2tet_kernel"] + "_load"))

    def capture(self, state, rhs_dev, fresh=True):
        """CUDA graph of one fresh assembly (``rhs_dev`` zeroed, every CSR value overwritten) on the buffers of ``state``:
        ``graph.replay()`` re-runs the kernels without the per-launch host work -- for time / Newton loops that
        reassemble on one plan.  Every entry listed with a state vector (``alpha`` not ``None``) keeps its state terms
        in the graph even when the vector is all zero at capture time: the kernels read the persistent device buffers
        ``data.alpha_dev`` of ``state[2]``, which ``prepare()`` refreshes in place (or ``data.alpha_dev.copy_(...)``).
        List ``alpha=None`` for a form whose state is identically zero to leave those terms out.  Full pattern policy only."""
        if state[4] is None:
            raise _lib.NfaError("GPUAssembler.capture: value-dependent pattern policies rebuild the plan every assembly")
        with torch.cuda.device(state[0].device):
            return self._capture(state, rhs_dev, fresh)

    def _capture(self, state, rhs_dev, fresh=True):
        dev = state[0]

        def body():
            self.jacobian.is_zero = True
            self._capturing = True
            try:
                self.run(state, rhs_dev, fresh=fresh)
            finally:
                self._capturing = False

        body()  # binds the pattern, allocates the values
        cur = torch.cuda.current_stream(dev.device)
        side = torch.cuda.Stream(device=dev.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            body()
            body()
        cur.wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            body()
        self._graph_keep = self._keep
        return graph

    def _run_filtered(self, state, rhs_dev):
        """Value-dependent pattern: local blocks first, then masks, then a plan over the kept entries."""
        dev, entries, datas, dests, _, _ = state
        lib = _lib.load()
        stream = _lib.stream_ptr()
        jac = self.jacobian
        if jac.plan_key is not None and not jac.is_zero:
            raise _lib.NfaError("GPUAssembler: a filtered pattern cannot be accumulated onto existing values; scale(0.0) first")
        n_dof = int(self.space.n_dof)
        if rhs_dev is None:
            rhs_dev = torch.zeros(n_dof, dtype=torch.float64, device=dev.device)
        nseg = len(entries)
        Ke_list, re_list, masks, keep = [], [], [], []
        tol = 1.0e-8 + 1.0e-5 * 1.0e-16  # np.isclose(value, 1e-16): |value - 1e-16| <= atol + rtol * |1e-16|
        for (_, kind, _, form, alpha), data, dest in zip(entries, datas, dests):
            if "time_dependent" in kind:
                raise NotImplementedError("GPUAssembler: filtered patterns are implemented for form / bc_form lists")
            alpha_dev = None if data.alpha_zero else data.alpha_dev
            Ke, re = form.local_blocks(data.dev, data, alpha_dev, dest if alpha_dev is not None else None)
            # one pass over the local matrices (nfa_filter_local_entries): value test, row <= col, forced diagonal; entries
            # that fail the value test are zeroed in place (a forced diagonal entry then adds 0.0)
            mask = torch.empty(Ke.shape, dtype=torch.uint8, device=dev.device)
            _lib.check(lib.nfa_filter_local_entries(int(Ke.shape[0]), int(Ke.shape[1]), _lib.ptr(dest), tol,
                                                    1 if self.pattern_policy == "filtered_upper" else 0,
                                                    1 if (self.explicit_diagonal and kind == "form") else 0, _lib.ptr(Ke),
                                                    _lib.ptr(mask), stream), "nfa_filter_local_entries")
            Ke_list.append(Ke)
            re_list.append(re)
            masks.append(mask)
            keep += [alpha_dev]
        t0 = time.perf_counter()
        # the pattern depends on the values, but in a Newton / time loop it rarely changes: when the masks (and the
        # destination tables) are the ones of the previous assembly, its plan -- a sort of every kept local entry -- is reused
        prev = self._filtered_plan
        same = prev is not None and prev.n_dof == n_dof and len(prev.masks) == nseg and \
            all(pd is d and pm.shape == m.shape and bool(torch.equal(pm, m)) for pd, pm, d, m in zip(prev.dests, prev.masks, dests, masks))
        if same:
            p = prev
            p.masks = masks  # the previous arrays may be released
            p.reused += 1
        else:
            p = _Plan()
            p.n_dof = n_dof
            dest_ptrs = (ctypes.c_void_p * nseg)(*[d.data_ptr() for d in dests])
            n_items = (ctypes.c_int64 * nseg)(*[int(d.shape[0]) for d in dests])
            n_loc = (ctypes.c_int32 * nseg)(*[int(d.shape[1]) for d in dests])
            mask_ptrs = (ctypes.c_void_p * nseg)(*[m.data_ptr() for m in masks])
            handle = ctypes.c_void_p()
            _lib.check(
                lib.nfa_plan_create_filtered(n_dof, nseg, dest_ptrs, n_items, n_loc, None, mask_ptrs, ctypes.byref(handle), stream),
                "nfa_plan_create_filtered",
            )
            p.plan = handle
            p.nnz = int(lib.nfa_plan_nnz(p.plan))
            p.indptr = torch.empty(n_dof + 1, dtype=torch.int32, device=dev.device)
            p.indices = torch.empty(p.nnz, dtype=torch.int32, device=dev.device)
            _lib.check(lib.nfa_plan_csr(p.plan, _lib.ptr(p.indptr), _lib.ptr(p.indices), stream), "nfa_plan_csr")
            p.build_seconds = time.perf_counter() - t0
            p.dests, p.masks, p.reused = list(dests), masks, 0
            self._filtered_plan = p
        if jac.plan_key != ("filtered", id(p)):
            jac.bind_pattern(("filtered", id(p)), n_dof, p.indptr, p.indices)
        ke_ptrs = (ctypes.c_void_p * nseg)(*[k.data_ptr() for k in Ke_list])
        re_ptrs = (ctypes.c_void_p * nseg)(*[r.data_ptr() for r in re_list])
        _lib.check(
            lib.nfa_plan_reduce(p.plan, ke_ptrs, re_ptrs, 0, _lib.ptr(jac.values_dev), 1, _lib.ptr(rhs_dev), stream),
            "nfa_plan_reduce",
        )
        jac.is_zero = False
        self._keep = keep + Ke_list + re_list + masks
        return rhs_dev

    def scatter_forms(self, measure_time_q=False):
        """The reference's call (R/src/assembly/SequentialAssembler.py:69-86): adds every listed form into ``jacobian`` and
        ``residual``.  From the second call on the coefficient arrays, destination tables and the plan come from the
        caches; only the state vectors go up and the residual comes down."""
        t0 = time.perf_counter()
        state = self.prepare()
        t1 = time.perf_counter()
        dev = state[0]
        n_dof = int(self.space.n_dof)
        if self._rhs_dev is None or self._rhs_dev.shape[0] != n_dof or self._rhs_dev.device != dev.device:
            self._rhs_dev = torch.empty(n_dof, dtype=torch.float64, device=dev.device)
            self._rhs_host = torch.empty(n_dof, dtype=torch.float64).pin_memory()
        rhs_dev = self.run(state, self._rhs_dev, fresh=True)
        self._rhs_host.copy_(rhs_dev, non_blocking=False)
        rhs = self._rhs_host.numpy()
        t2 = time.perf_counter()
        if self.residual is None:
            self.residual = np.zeros(len(rhs))
        if isinstance(self.residual, np.ndarray) and self.residual.ndim == 1 and self.residual.dtype == np.float64 \
                and self.residual.flags.c_contiguous:
            _add_into(self.residual, rhs)
        else:
            self.residual += rhs
        t3 = time.perf_counter()
        plan = state[4] if state[4] is not None else self._filtered_plan
        self.last_timings = {"prepare_s": t1 - t0, "kernels_s": t2 - t1, "plan_build_s": plan.build_seconds,
                             "residual_add_s": t3 - t2, **getattr(self, "_prepare_split", {})}
        if measure_time_q:
            print("GPUAssembler:: prepare (callables, upload, plan):", t1 - t0, "seconds")
            print("GPUAssembler:: kernels + residual download:", t2 - t1, "seconds")

    def assemble(self):
        """One-shot: scatter all forms into a zeroed system and return ``(indptr, indices, data, rhs)``."""
        self.jacobian.scale(0.0)
        if self.residual is not None:
            self.residual[:] = 0.0
        self.scatter_forms()
        indptr, indices, data = self.jacobian.getValuesCSR()
        return indptr, indices, data, self.residual


# the reference's name, so that `from assembly.SequentialAssembler import SequentialAssembler` ports by import path
SequentialAssembler = GPUAssembler
