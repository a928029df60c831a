"""ORACLE-SIDE BASELINE (test / bench infrastructure only -- never imported by the product path).

The "strong CPU" baseline of BASELINE.md section 3 for config 2: vector-P2 elasticity on affine tetrahedra assembled
with batched numpy (one einsum per stage over all cells, the exact reference stiffness tensor instead of a quadrature
loop) and ``scipy.sparse`` COO -> CSR, i.e. what a careful CPU implementation of the reference's path
(R/src/weak_forms/le_primal_weak_form.py:91-126 + R/src/assembly/ScatterFormData.py:5-11) would do without changing
the result: same pattern (all ``n_loc^2`` pairs of every cell, duplicates summed), same values up to round-off.
"""
import numpy as np
import scipy.sparse as sp

from . import ref_basis


def reference_stiffness_tensor(dim, k):
    """``M[a, b, i, j] = int_ref d_i phi_a d_j phi_b`` of the scalar Lagrange element (exact: the rule integrates degree
    2k + 1 >= 2k - 2)."""
    pts, w = ref_basis.gauss_jacobi(dim, 2 * k + 1)
    tab = ref_basis.tabulate_lagrange(dim, k, pts)[:, :, :, 0]
    return np.einsum("q,iqa,jqb->abij", w, tab[1:], tab[1:])


def assemble_elasticity(points, cell_nodes, dest, n_dof, lam, mu, k=2):
    """CSR ``(indptr, indices, data)`` of the elasticity matrix for cell-wise constant ``lam``, ``mu``.
    ``cell_nodes`` [n, 4] vertex ids, ``dest`` [n, n_phi * 3] destination indices (basis-major, component-minor)."""
    dim = 3
    M = reference_stiffness_tensor(dim, k)
    X = points[cell_nodes]
    J = np.transpose(X[:, 1:] - X[:, :1], (0, 2, 1))  # J[n, r, c] = dx_r / dxi_c
    det = np.abs(np.linalg.det(J))
    Jinv = np.linalg.inv(J)  # [n, c, r] = dxi_c / dx_r
    # T[n, a, b, r, s] = int d_r phi_a d_s phi_b / det = sum_ij Jinv[i, r] M[a, b, i, j] Jinv[j, s]
    T = np.einsum("nir,abij,njs->nabrs", Jinv, M, Jinv, optimize=True)
    tr = np.einsum("nabrr->nab", T)
    lam = np.broadcast_to(np.asarray(lam, dtype=float), det.shape)
    mu = np.broadcast_to(np.asarray(mu, dtype=float), det.shape)
    # K[(a, r), (b, s)] = det (lam T[a, b, r, s] + mu T[a, b, s, r] + mu tr(T[a, b]) delta_rs)
    K = (det * lam)[:, None, None, None, None] * T + (det * mu)[:, None, None, None, None] * np.swapaxes(T, 3, 4)
    K = K + ((det * mu)[:, None, None] * tr)[:, :, :, None, None] * np.eye(dim)
    n, n_phi = K.shape[0], K.shape[1]
    Ke = np.transpose(K, (0, 1, 3, 2, 4)).reshape(n, n_phi * dim, n_phi * dim)
    rows = np.repeat(dest, dest.shape[1], axis=1).reshape(-1)
    cols = np.tile(dest, (1, dest.shape[1])).reshape(-1)
    A = sp.coo_matrix((Ke.reshape(-1), (rows, cols)), shape=(n_dof, n_dof)).tocsr()
    A.sort_indices()
    return A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data
