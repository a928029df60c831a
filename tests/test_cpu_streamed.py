"""Rank tables of the streamed scatter (assembly/streamed.py), walked on the CPU: positions are unique, cover the sorted
union pattern, and the two-pass store / add rule reproduces the dense sum of the local matrices."""
import numpy as np
import pytest
import torch

from nfempy_b200.assembly.streamed import build_tables
from nfempy_b200.mesh import box_tet_mesh, square_tri_mesh
from nfempy_b200.spaces import ProductSpace


def _space(dim, k, n):
    mesh = box_tet_mesh(n, jitter=0.2, shuffle_local_nodes=True) if dim == 3 else square_tri_mesh(n, jitter=0.2)
    tc = 3 if dim == 3 else 1
    d = {"s": (dim, dim, "BDM", k, mesh), "m": (dim, tc, "RT", k, mesh), "u": (dim, dim, "Lagrange", k - 1, mesh),
         "t": (dim, tc, "Lagrange", k - 1, mesh)}
    space = ProductSpace(d)
    space.make_subspaces_discontinuous({"u": True, "t": True})
    tags = [2, 3, 4, 5, 6, 7] if dim == 3 else [2, 3, 4, 5]
    space.build_structures({f: [1] for f in d}, {"s": tags, "m": tags})
    return space


@pytest.mark.parametrize("dim,k,n,batch", [(3, 1, 2, 7), (2, 2, 3, 1000), (3, 2, 1, 3)])
def test_tables_reproduce_the_dense_sum(dim, k, n, batch):
    space = _space(dim, k, n)
    dest = np.asarray(space.destination_table(), dtype=np.int64)
    n_cells, n_loc = dest.shape
    n_dof = int(space.n_dof)
    T = build_tables(torch.as_tensor(dest), n_dof, pair_batch=batch)
    indptr = T["indptr"].numpy()
    tab_of = T["tab_of"].numpy()
    tabs = T["tabs"].numpy().view(np.uint16)
    src, src_off = T["src"].numpy(), T["src_off"].numpy()
    # pattern: row r holds the sorted union of the dofs of its cells
    cols_of = [set() for _ in range(n_dof)]
    for c in range(n_cells):
        for a in dest[c]:
            cols_of[a].update(dest[c].tolist())
    indices = np.empty(indptr[-1], dtype=np.int64)
    for r in range(n_dof):
        want = np.array(sorted(cols_of[r]))
        assert indptr[r + 1] - indptr[r] == len(want)
        got = src[src_off[r] : src_off[r] + len(want)]
        assert np.array_equal(got, want)
        indices[indptr[r] : indptr[r + 1]] = got
    # two passes over random local matrices against the dense sum
    rng = np.random.default_rng(3)
    Ke = rng.standard_normal((n_cells, n_loc, n_loc))
    dense = np.zeros((n_dof, n_dof))
    for c in range(n_cells):
        dense[np.ix_(dest[c], dest[c])] += Ke[c]
    vals = np.full(indptr[-1], np.nan)
    for p in (0, 1):
        for c in range(n_cells):
            for a in range(n_loc):
                t = tabs[tab_of[c, a]]
                pos = indptr[dest[c, a]] + (t & 0x7FFF)
                sel = (t >> 15) == p
                if p == 0:
                    assert np.all(np.isnan(vals[pos[sel]])), "an entry is stored twice"
                    vals[pos[sel]] = Ke[c, a, sel]
                else:
                    assert not np.isnan(vals[pos[sel]]).any(), "an entry is added to before it was stored"
                    vals[pos[sel]] += Ke[c, a, sel]
            assert np.array_equal(indices[indptr[dest[c, 0]] + (tabs[tab_of[c, 0]] & 0x7FFF)], dest[c])
    assert not np.isnan(vals).any(), "an entry of the pattern was never written"
    rows = np.repeat(np.arange(n_dof), np.diff(indptr))
    assert np.abs(vals - dense[rows, indices]).max() <= 1e-13
    mask = np.ones_like(dense, dtype=bool)
    mask[rows, indices] = False
    assert not dense[mask].any()


def test_a_dof_in_three_cells_is_refused():
    mesh = square_tri_mesh(2, jitter=0.1)
    space = ProductSpace({"p": (2, 1, "Lagrange", 1, mesh)})
    space.build_structures({"p": [1]}, {})
    dest = torch.as_tensor(np.asarray(space.destination_table(), dtype=np.int64))
    with pytest.raises(NotImplementedError):
        build_tables(dest, int(space.n_dof))


def test_partition_with_replicated_halo_cells():
    """Every row a rank's own cells touch has all of its cells on that rank; the owned rows partition the dofs."""
    from nfempy_b200.assembly.streamed import partition_cells

    space = _space(3, 1, 3)
    dest = np.asarray(space.destination_table(), dtype=np.int64)
    owned_all = []
    for r in range(3):
        cells, owned, (c0, c1) = partition_cells(dest, 3, r)
        assert np.all(np.diff(cells) > 0) and set(range(c0, c1)) <= set(cells.tolist())
        rows = np.unique(dest[c0:c1])
        touching = np.flatnonzero(np.isin(dest, rows).any(axis=1))
        assert set(touching.tolist()) <= set(cells.tolist())
        assert np.isin(owned, rows).all()
        owned_all.append(owned)
    u = np.concatenate(owned_all)
    assert len(u) == len(np.unique(u)) == space.n_dof
