"""Host side of the row-owner H(div) path (``nfempy_b200/assembly/hdiv_rows.py``) on the CPU: the rows' sources, the
record layouts / entry maps of ``nfa_hdiv_rt1_records`` + ``nfa_rows_expand`` and the slot search, driven with records
read off the ORACLE's local matrices -- expanding them with a numpy restatement of the two kernels must reproduce the
oracle's assembled CSR (pattern, values, right-hand side) entry by entry."""
import numpy as np
import pytest
import torch

from nfempy_b200.assembly import hdiv_rows
from oracle import ref_assembly, ref_forms
from tests.parity_helpers import SundusCase, mixed_poisson_case


def expand_rows_host(indptr, row_lo, row_src, slots, roff, codes, rec):
    """numpy restatement of ``rows_expand_kernel`` (csrc/nfa_hdivrows.cu)."""
    n_rows = len(row_src)
    vals = np.zeros(int(indptr[row_lo + n_rows] - indptr[row_lo]))
    rhs = np.zeros(n_rows)
    base = int(indptr[row_lo])
    for r in range(n_rows):
        off = int(indptr[row_lo + r]) - base
        for side in range(2):
            s = int(row_src[r, side])
            if s < 0:
                continue
            c, a = s >> 5, s & 31
            for k in range(hdiv_rows.RUN):
                code = int(codes[a, k])
                if code == 0:
                    continue
                v = rec[c, roff[a] + k]
                if code == hdiv_rows.CODE_RHS:
                    rhs[r] += v
                elif side == 0:
                    vals[off + int(slots[r, side, k])] = v if code > 0 else -v
                else:
                    vals[off + int(slots[r, side, k])] += v if code > 0 else -v
    return vals, rhs


def locate_slots_host(indptr, indices, row_lo, row_src, codes, dest):
    """numpy restatement of ``rows_locate_kernel``."""
    n_rows = len(row_src)
    slots = np.zeros((n_rows, 2, hdiv_rows.RUN), dtype=np.uint8)
    for r in range(n_rows):
        b0, b1 = int(indptr[row_lo + r]), int(indptr[row_lo + r + 1])
        seg = np.asarray(indices[b0:b1])
        for side in range(2):
            s = int(row_src[r, side])
            if s < 0:
                continue
            c, a = s >> 5, s & 31
            for k in range(hdiv_rows.RUN):
                code = int(codes[a, k])
                if code == 0 or code == hdiv_rows.CODE_RHS:
                    continue
                col = dest[c, abs(code) - 1]
                pos = int(np.searchsorted(seg, col))
                if pos >= len(seg) or seg[pos] != col:
                    raise ValueError("a local column is missing from the row's CSR segment")
                slots[r, side, k] = pos
    return slots


def _records_from_local(emap, ridx, stride, blocks):
    """Fills the records from local ``(r_el, j_el)`` pairs THROUGH the entry maps; checks on the way that the maps describe
    the local matrix completely: structural zeros are zeros, entries sharing a record slot agree (up to the sign)."""
    rec = np.full((len(blocks), stride), np.nan)
    for c, (r_el, j_el) in enumerate(blocks):
        scale = np.abs(j_el).max()
        for a in range(emap.shape[0]):
            for b in range(emap.shape[1]):
                code = int(emap[a, b])
                if code == 0:
                    assert abs(j_el[a, b]) <= 1e-13 * scale, (a, b, j_el[a, b])
                    continue
                v = j_el[a, b] if code > 0 else -j_el[a, b]
                k = abs(code) - 1
                if not np.isnan(rec[c, k]):
                    assert abs(rec[c, k] - v) <= 1e-12 * scale, (a, b)
                rec[c, k] = v
            assert np.isnan(rec[c, ridx[a]])
            rec[c, ridx[a]] = r_el[a]
    return np.nan_to_num(rec)


def _check(dest, n_dof, emap, ridx, stride, blocks, want):
    indptr, indices, data, rhs = want
    found = hdiv_rows.row_sources(torch.as_tensor(dest))
    assert found is not None
    rows, src = found[0].numpy(), found[1].numpy()
    assert np.array_equal(rows, np.arange(n_dof))  # the whole mesh: every row is fed
    cuts = hdiv_rows.row_blocks(found[0]).numpy()
    assert np.array_equal(cuts[:, 0], np.arange(0, n_dof, hdiv_rows.ROWS_PER_BLOCK)) and cuts[:, 1].sum() == n_dof
    assert np.array_equal(cuts[:, 0], cuts[:, 2]) and cuts[:, 1].max() <= hdiv_rows.ROWS_PER_BLOCK
    # every row names its (item, local row) sources, lower item first
    for r in range(n_dof):
        hits = [(c, a) for c in range(len(dest)) for a in range(dest.shape[1]) if dest[c, a] == r]
        got = [(int(s) >> 5, int(s) & 31) for s in src[r] if s >= 0]
        assert got == hits
    runs = hdiv_rows.row_runs(emap, ridx)
    assert runs is not None
    roff, codes = runs
    assert not (roff % 2).any() and (roff + hdiv_rows.RUN <= stride + hdiv_rows.RUN - 2).all()
    for a in range(len(roff)):  # the 16-byte loads of a run stay inside the record
        used = int(np.flatnonzero(codes[a]).max()) + 1
        assert roff[a] + (used + 1) // 2 * 2 <= stride
    slots = locate_slots_host(indptr, indices, 0, src, codes, dest)
    rec = _records_from_local(emap, ridx, stride, blocks)
    vals, res = expand_rows_host(indptr, 0, src, slots, roff, codes, rec)
    assert np.allclose(vals, data, rtol=1e-13, atol=1e-13 * np.abs(data).max())
    assert np.allclose(res, rhs, rtol=1e-13, atol=1e-13 * np.abs(rhs).max())


@pytest.mark.parametrize("dim", [2, 3])
def test_sundus_records_expand_to_the_oracle_matrix(dim):
    case = SundusCase(dim=dim, k=1, n=2)
    space, ref = case.space(), case.ref_space()
    a_n, a = case.alphas(ref.n_dof)
    mp = ref.discrete_spaces["mp"]
    n_cells = len(mp.element_ids)
    forms = [("time_dependent_form", ref_forms.sundus_dual, case.functions, n_cells)]
    want = ref_assembly.assemble_time_dependent(ref, forms, (a_n, a), case.t)
    dest = np.asarray(space.destination_table(), dtype=np.int64)
    blocks = [ref_forms.sundus_dual(ref, case.functions, i, a_n[dest[i]], a[dest[i]], case.t) for i in range(n_cells)]
    emap, ridx, stride = hdiv_rows.sundus_maps(dim + 1)
    assert stride % 4 == 0 and stride >= np.abs(emap).max() and stride > ridx.max()
    _check(dest, ref.n_dof, emap, ridx, stride, blocks, want)


@pytest.mark.parametrize("dim", [2, 3])
def test_mixed_poisson_records_expand_to_the_oracle_matrix(dim):
    case = mixed_poisson_case(dim=dim, k=1, n=2, with_bc=False, varying=True)
    space, ref = case.space(), case.ref_space()
    n_cells = len(ref.discrete_spaces["u"].element_ids)
    forms = [("form", ref_forms.laplace_dual, case.ref_forms_fn(ref)[0][2], n_cells)]
    want = ref_assembly.assemble(ref, forms)
    dest = np.asarray(space.destination_table(), dtype=np.int64)
    functions = case.ref_forms_fn(ref)[0][2]
    blocks = [ref_forms.laplace_dual(ref, functions, i, np.zeros(dest.shape[1])) for i in range(n_cells)]
    emap, ridx, stride = hdiv_rows.laplace_dual_maps(dim + 1)
    _check(dest, ref.n_dof, emap, ridx, stride, blocks, want)


def test_row_sources_refuse_spaces_with_widely_shared_dofs():
    dest = torch.tensor([[0, 1, 2], [2, 3, 4], [4, 2, 5]])  # dof 2 belongs to three items
    assert hdiv_rows.row_sources(dest) is None
    assert hdiv_rows.row_sources(torch.tensor([[0, 1, 1]])) is None  # an item lists a dof twice
    rows, src = hdiv_rows.row_sources(torch.tensor([[3, 1], [1, 2]]))  # rows 1, 2, 3 are fed; 0 and 4.. are not
    assert rows.tolist() == [1, 2, 3] and src.tolist() == [[1, 32], [33, -1], [0, -1]]


def test_row_blocks_cut_runs_of_consecutive_rows():
    rows = torch.tensor(list(range(5, 8)) + list(range(20, 20 + 600)) + [1000], dtype=torch.int32)
    blocks = hdiv_rows.row_blocks(rows).tolist()
    assert blocks == [[5, 3, 0, 0], [20, 256, 3, 0], [276, 256, 259, 0], [532, 88, 515, 0], [1000, 1, 603, 0]]


@pytest.mark.parametrize("kind,dim", [("degenerate", 3), ("unscaled", 2)])
def test_general_elliptic_records_expand_to_the_oracle_matrix(kind, dim):
    """Record layout / entry maps of form 2 (DegenerateEllipticWeakForm, UnscaledEllipticWeakForm on RT_1 x P_0) against
    the oracle's local matrices and assembled system, non-zero state vector."""
    from tests.parity_helpers import degenerate_case

    case = degenerate_case(kind=kind, dim=dim, k=1, n=2, with_bc=False)
    space, ref = case.space(), case.ref_space()
    alpha = case.alpha_fn(ref.n_dof)
    (_, form_fn, functions, n_cells), = case.ref_forms_fn(ref)
    want = ref_assembly.assemble(ref, [("form", form_fn, functions, n_cells)], alpha)
    dest = np.asarray(space.destination_table(), dtype=np.int64)
    blocks = [form_fn(ref, functions, i, alpha[dest[i]]) for i in range(n_cells)]
    emap, ridx, stride = hdiv_rows.general_maps(dim + 1)
    _check(dest, ref.n_dof, emap, ridx, stride, blocks, want)


def test_row_runs_describe_every_map_and_refuse_what_does_not_fit():
    """``row_runs``: every local row's record entries fit one 16-byte aligned run of eight doubles for the three record
    layouts (both dimensions); a row whose entries span more, or a record entry that feeds two columns of one row, is
    refused (the form list then takes the two-stage path)."""
    for maps in (hdiv_rows.sundus_maps, hdiv_rows.laplace_dual_maps, hdiv_rows.general_maps):
        for n_u in (3, 4):
            emap, ridx, stride = maps(n_u)
            roff, codes = hdiv_rows.row_runs(emap, ridx)
            assert roff.dtype == np.int32 and codes.dtype == np.int8 and codes.shape == (emap.shape[0], hdiv_rows.RUN)
            for a in range(emap.shape[0]):
                assert roff[a] % 2 == 0 and roff[a] + hdiv_rows.RUN <= stride + 6
                seen_cols = sorted(abs(int(c)) - 1 for c in codes[a] if c not in (0, hdiv_rows.CODE_RHS))
                assert seen_cols == sorted(np.flatnonzero(emap[a]).tolist())
                assert (codes[a] == hdiv_rows.CODE_RHS).sum() == 1
                k = int(np.flatnonzero(codes[a] == hdiv_rows.CODE_RHS)[0])
                assert roff[a] + k == ridx[a]
                for kk, c in enumerate(codes[a]):
                    if c not in (0, hdiv_rows.CODE_RHS):
                        b = abs(int(c)) - 1
                        assert abs(int(emap[a, b])) - 1 == roff[a] + kk and np.sign(emap[a, b]) == np.sign(c)
    wide = np.zeros((2, 2), dtype=np.int16)
    wide[0, 0], wide[0, 1] = 1, 12  # entries 0 and 11 of the record: more than one run
    assert hdiv_rows.row_runs(wide, np.array([0, 1], dtype=np.int16)) is None
    twice = np.array([[1, 1], [0, 2]], dtype=np.int16)  # record entry 0 feeds two columns of row 0
    assert hdiv_rows.row_runs(twice, np.array([3, 3], dtype=np.int16)) is None
