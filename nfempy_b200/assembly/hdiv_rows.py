"""Row-owner assembly of lowest-order H(div) x L2 forms (``csrc/nfa_hdivrows.cu``): the host side.

``LaplaceDualWeakForm`` on RT_1 x P_0 (BASELINE config 3 at k = 1) and ``SundusDualWeakForm`` on RT_1 x RT_1 x P_0 x P_0
(config 5) take this path inside ``GPUAssembler`` when the pattern policy is "full": the form's ``row_spec`` describes a
compact per-cell record and which record entry feeds local entry (a, b); one thread per cell writes the records, one thread
per matrix row expands them into the CSR (lower cell first: the order of the reference's ``Mat.setValue(..., addv=True)``
calls, R/src/assembly/ScatterFormData.py:5-11).  Works in spaces where a dof belongs to at most two items of the form
(facet dofs of H(div) fields, cell-local dofs).  Boundary / coupling forms of the same list go through the generic plan,
merged into the same CSR.
"""
import numpy as np
import torch


ROWS_PER_BLOCK = 256  # rows of one CTA pass of rows_expand_kernel (csrc/nfa_hdivrows.cu)


def row_sources(dest):
    """The rows the items feed and their sources: ``(rows, src)`` with ``rows`` int32 ``[n_rows]`` (ascending global
    dofs) and ``src`` int32 ``[n_rows, 2]`` = the one or two sources of each row as ``item * 32 + local row`` (``-1`` =
    none), lower item first.  ``dest``: integer tensor ``[n_items, n_loc]`` of global dofs (any device).  Returns ``None``
    when a dof belongs to more than two items or is listed twice by one item."""
    n_items, n_loc = int(dest.shape[0]), int(dest.shape[1])
    if n_loc > 32 or n_items >= 2**26 or n_items == 0:
        return None
    flat = dest.reshape(-1).to(torch.int64)
    if int(flat.min().item()) < 0 or int(flat.max().item()) >= 2**31 - 1:
        return None
    order = torch.argsort(flat, stable=True)  # ascending dof; ascending (item, local row) inside a dof
    sd = flat[order]
    first = torch.ones_like(sd, dtype=torch.bool)
    first[1:] = sd[1:] != sd[:-1]
    start = torch.nonzero(first).reshape(-1)
    count = torch.cat([start[1:], torch.tensor([sd.shape[0]], device=sd.device)]) - start
    if int(count.max().item()) > 2:
        return None
    item = torch.div(order, n_loc, rounding_mode="floor")
    code = (item * 32 + (order - item * n_loc)).to(torch.int32)
    two = count == 2
    if bool((item[start[two]] == item[start[two] + 1]).any()):
        return None
    src = torch.full((int(start.shape[0]), 2), -1, dtype=torch.int32, device=dest.device)
    src[:, 0] = code[start]
    src[two, 1] = code[start[two] + 1]
    return sd[start].to(torch.int32).contiguous(), src.contiguous()


def row_blocks(rows):
    """Cuts the ascending row list into blocks of at most ``ROWS_PER_BLOCK`` CONSECUTIVE rows: int32 ``[n_blocks, 4]`` =
    (first row, number of rows, index of the first row in ``rows``, 0)."""
    n = int(rows.shape[0])
    r = rows.to(torch.int64)
    idx = torch.arange(n, device=rows.device, dtype=torch.int64)
    brk = torch.ones(n, dtype=torch.bool, device=rows.device)
    brk[1:] = r[1:] != r[:-1] + 1
    run_start = torch.cummax(torch.where(brk, idx, torch.zeros_like(idx)), 0).values  # index of the run's first row
    head = ((idx - run_start) % ROWS_PER_BLOCK) == 0
    first = torch.nonzero(head).reshape(-1)
    count = torch.cat([first[1:], torch.tensor([n], device=rows.device)]) - first
    blocks = torch.zeros((int(first.shape[0]), 4), dtype=torch.int32, device=rows.device)
    blocks[:, 0] = r[first].to(torch.int32)
    blocks[:, 1] = count.to(torch.int32)
    blocks[:, 2] = first.to(torch.int32)
    return blocks.contiguous()


def sundus_maps(n_u):
    """Record layout of ``nfa_hdiv_rt1_records`` (form 1) as the tables ``nfa_rows_expand`` reads: local order
    [mp | mc | p | c]; ``emap[a, b]`` = +-(record index + 1) or 0 for a structural zero, ``ridx[a]`` = residual entry."""
    roww, n_loc = (n_u + 3) // 2 * 2, 2 * n_u + 2  # rows start 16-byte aligned
    soff = 2 * n_u * roww
    emap = np.zeros((n_loc, n_loc), dtype=np.int16)
    ridx = np.zeros(n_loc, dtype=np.int16)
    for F in range(2):
        s_col = 2 * n_u + F  # the scalar field paired with flux field F
        for a in range(n_u):
            base = (F * n_u + a) * roww
            for b in range(n_u):
                emap[F * n_u + a, F * n_u + b] = base + b + 1  # flux mass
            emap[F * n_u + a, s_col] = -(base + n_u + 1)  # -(w, div psi_a)
            emap[s_col, F * n_u + a] = soff + 4 + a + 1  # +(div psi_a, w): the copy inside the scalar rows' block
            ridx[F * n_u + a] = base + n_u + 1
        emap[s_col, s_col] = soff + F + 1  # time derivative (+ reaction tangent on c)
        ridx[s_col] = soff + 2 + F
    stride = (soff + 4 + n_u + 3) // 4 * 4
    return emap, ridx, stride


def laplace_dual_maps(n_u):
    """The same for the mixed Poisson form (form 0), local order [u | p]; the (p, p) entry is a structural zero."""
    roww, n_loc = (n_u + 3) // 2 * 2, n_u + 1
    soff = n_u * roww
    emap = np.zeros((n_loc, n_loc), dtype=np.int16)
    ridx = np.zeros(n_loc, dtype=np.int16)
    for a in range(n_u):
        base = a * roww
        for b in range(n_u):
            emap[a, b] = base + b + 1
        emap[a, n_u] = -(base + n_u + 1)
        emap[n_u, a] = soff + 1 + a + 1  # the copy inside the scalar row's block
        ridx[a] = base + n_u + 1
    ridx[n_u] = soff
    stride = (soff + 1 + n_u + 3) // 4 * 4
    return emap, ridx, stride


def general_maps(n_u):
    """The same for the generalised mixed elliptic forms (form 2, ``HdivL2GeneralForm``), local order [u | p]: like the
    mixed Poisson layout plus the (p, p) mass; scalar block ``K_pp, r_p, C[.]``."""
    roww, n_loc = (n_u + 3) // 2 * 2, n_u + 1
    soff = n_u * roww
    emap = np.zeros((n_loc, n_loc), dtype=np.int16)
    ridx = np.zeros(n_loc, dtype=np.int16)
    for a in range(n_u):
        base = a * roww
        for b in range(n_u):
            emap[a, b] = base + b + 1
        emap[a, n_u] = -(base + n_u + 1)
        emap[n_u, a] = soff + 2 + a + 1
        ridx[a] = base + n_u + 1
    emap[n_u, n_u] = soff + 1
    ridx[n_u] = soff + 1
    stride = (soff + 2 + n_u + 3) // 4 * 4
    return emap, ridx, stride


RUN = 8  # record doubles a local row may read (csrc/nfa_hdivrows.cu)
CODE_RHS = 127


def row_runs(emap, ridx):
    """The entry maps as the RUNS ``nfa_rows_locate`` / ``nfa_rows_expand`` read: local row ``a`` loads the record doubles
    ``[roff[a], roff[a] + 8)`` with 16-byte loads; ``codes[a, k]`` says what entry ``k`` is for that row: 0 = nothing,
    ``+-(b + 1)`` = local column ``b`` with that sign, 127 = the row's residual.  ``None`` when a row's entries do not fit
    one run or a record entry feeds two columns of a row."""
    n_loc = emap.shape[0]
    roff = np.zeros(n_loc, dtype=np.int32)
    codes = np.zeros((n_loc, RUN), dtype=np.int8)
    if n_loc > 126:
        return None
    for a in range(n_loc):
        used = {int(ridx[a]): CODE_RHS}
        for b in range(n_loc):
            code = int(emap[a, b])
            if code == 0:
                continue
            k = abs(code) - 1
            if k in used:
                return None
            used[k] = (b + 1) if code > 0 else -(b + 1)
        lo = min(used) // 2 * 2
        if max(used) - lo >= RUN:
            return None
        roff[a] = lo
        for k, code in used.items():
            codes[a, k - lo] = code
    return roff, codes
