"""ORACLE (test infrastructure only -- never imported by the product path).

H(div) pieces of the basix stand-in, written independently of ``nfempy_b200.basis.hdiv`` (numerical quadrature, moments
snapped to their exact rationals and a 60-digit solve here; exact rational arithmetic from the start there):

* ``tabulate``                -- RT_k (k = 1..3) / BDM_1..3 values and reference derivatives, ``[1+dim, nq, nphi, dim]``;
* ``entity_transformations``  -- signed permutations of the facet dof families, as explicit tables;
* ``permute_and_transform``   -- R/src/basis/permute_and_transform.py:8-149 line by line (edge test in 2-D against the
  edge cell's own ``node_tags``, face table of six permutation cases, forced reflection of tagged faces with
  local index 1 or 3, faces before edges).

PARITY UNPINNED against real basix (see oracle/ref_basis.py); ``permute_and_transform`` below is pinned against the
reference's own (tests/test_cpu_reference_pins.py: the reference's source run over these tables).
"""
from itertools import permutations

import numpy as np

from . import ref_basis
from .ref_basis import SUB

_REF_V = {
    1: np.array([[0.0], [1.0]]),
    2: np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]),
    3: np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]),
}


def _monomials(dim, degree):
    out = []
    for total in range(degree + 1):
        for e in np.ndindex(*([total + 1] * dim)):
            if sum(e) == total:
                out.append(e)
    return out


def _mono_tab(exps, pts):
    """values [nq, nm] and gradients [dim, nq, nm] of monomials."""
    pts = np.asarray(pts, dtype=float)
    nq, dim = pts.shape
    val = np.ones((nq, len(exps)))
    grad = np.zeros((dim, nq, len(exps)))
    for m, e in enumerate(exps):
        val[:, m] = np.prod(pts ** np.array(e), axis=1)
        for d in range(dim):
            if e[d] > 0:
                ee = np.array(e)
                ee[d] -= 1
                grad[d, :, m] = e[d] * np.prod(pts**ee, axis=1)
    return val, grad


def _cubic_lattice(fdim):
    """Equispaced cubic nodes on the reference interval / triangle: vertices, two per edge (1/3, 2/3 from the edge's first
    vertex, edges (1,2), (0,2), (0,1) on the triangle), centroid."""
    v = _REF_V[fdim]
    nodes = [x for x in v]
    for a, b in SUB[fdim][1]:
        nodes += [(2.0 * v[a] + v[b]) / 3.0, (v[a] + 2.0 * v[b]) / 3.0]
    if fdim == 2:
        nodes.append(v.mean(axis=0))
    return np.array(nodes)


_CUBIC = {}


def _cubic_family_coefficients(fdim):
    """Monomial coefficients [n_mono, n_nodes] of the cubic nodal family on the lattice, exact (entries are k / 2)."""
    if fdim not in _CUBIC:
        from fractions import Fraction

        exps = _monomials(fdim, 3)
        nodes = [[Fraction(int(round(3 * float(c))), 3) for c in x] for x in _cubic_lattice(fdim)]
        n = len(exps)
        A = [[Fraction(1)] * 0 for _ in range(n)]
        for i, x in enumerate(nodes):
            A[i] = [Fraction(1) for _ in exps]
            for m, e in enumerate(exps):
                v = Fraction(1)
                for c, p in zip(x, e):
                    v *= c ** p
                A[i][m] = v
        # Gauss-Jordan on [A | I]
        aug = [row + [Fraction(int(i == j)) for j in range(n)] for i, row in enumerate(A)]
        for col in range(n):
            piv = next(r for r in range(col, n) if aug[r][col] != 0)
            aug[col], aug[piv] = aug[piv], aug[col]
            pv = aug[col][col]
            aug[col] = [v / pv for v in aug[col]]
            for r in range(n):
                if r != col and aug[r][col] != 0:
                    f = aug[r][col]
                    aug[r] = [a - f * b for a, b in zip(aug[r], aug[col])]
        _CUBIC[fdim] = np.array([[float(aug[i][n + j]) for j in range(n)] for i in range(n)])
    return _CUBIC[fdim]


def _facet_family(fdim, degree, s):
    """Nodal Lagrange family of ``degree`` on the reference facet at parameter points ``s`` -> [nq, n]."""
    if fdim == 0:
        return np.ones((len(s), 1))
    if degree == 3:  # BDM_3: equispaced cubic family; the Vandermonde matrix of the lattice is inverted in rationals
        exps = _monomials(fdim, 3)
        return _mono_tab(exps, s)[0] @ _cubic_family_coefficients(fdim)
    return ref_basis.tabulate_lagrange(fdim, degree, s)[0, :, :, 0]


def _nedelec2(dim, pts):
    """[n_fn, nq, dim]: a basis of the degree-2 first-kind Nedelec space at ``pts`` -- [P1]^d, then (-y, x) x, (-y, x) y
    (2-D) or x cross (x_i e_j), (i, j) != (2, 2) (3-D).  BDM_3's interior moments are taken against these."""
    pts = np.asarray(pts, dtype=float)
    nq = len(pts)
    out = []
    p1 = _mono_tab(_monomials(dim, 1), pts)[0].T  # the P1 monomials in the exponent order both implementations use
    for c in range(dim):
        for m in p1:
            w = np.zeros((nq, dim))
            w[:, c] = m
            out.append(w)
    if dim == 2:
        rot = np.stack([-pts[:, 1], pts[:, 0]], axis=1)
        out += [rot * pts[:, [0]], rot * pts[:, [1]]]
    else:
        for i in range(3):
            for j in range(3):
                if (i, j) == (2, 2):
                    continue
                e = np.zeros(3)
                e[j] = 1.0
                out.append(np.cross(pts, pts[:, [i]] * e[None, :]))
    return np.array(out)


_CACHE = {}

# Every entry of the dual matrix is a moment of a polynomial with integer (cubic facet family: half-integer) coefficients
# over the reference simplex or one of its facets: a rational whose denominator divides (degree + dim)! times those few
# factors.  The Gauss-Jacobi rules reproduce the moments to ~1e-16, so each entry can be snapped to the exact rational
# (asserted unambiguous) and the system solved in 60-digit arithmetic: the basis coefficients are then correct to the last
# bit of a double, instead of carrying the 1e-11 the condition number of a 60 x 60 float solve leaves (BDM_3).
_DENOM = 362880 * 16 * 27  # 9! * 2^4 * 3^3


def _solve_dual(D):
    import mpmath

    scaled = D * _DENOM
    snapped = np.rint(scaled)
    if np.abs(scaled - snapped).max() > 1e-4:
        raise AssertionError("oracle H(div) element: a dual-matrix moment is not a multiple of 1 / %d" % _DENOM)
    old = mpmath.mp.dps
    mpmath.mp.dps = 60
    try:
        M = mpmath.matrix([[mpmath.mpf(int(v)) / _DENOM for v in row] for row in snapped])
        Ci = mpmath.inverse(M)
        return np.array([[float(Ci[i, j]) for j in range(Ci.cols)] for i in range(Ci.rows)])
    finally:
        mpmath.mp.dps = old


def _element(family, dim, k):
    key = (family, dim, k)
    if key in _CACHE:
        return _CACHE[key]
    base_deg = k - 1 if family == "RT" else k
    exps = _monomials(dim, k)
    index = {e: i for i, e in enumerate(exps)}
    # space functions as coefficient arrays [n_space, dim, n_mono]
    space = []
    for c in range(dim):
        for e in _monomials(dim, base_deg):
            a = np.zeros((dim, len(exps)))
            a[c, index[e]] = 1.0
            space.append(a)
    if family == "RT":
        for e in _monomials(dim, k - 1):
            if sum(e) != k - 1:
                continue
            a = np.zeros((dim, len(exps)))
            for c in range(dim):
                ee = list(e)
                ee[c] += 1
                a[c, index[tuple(ee)]] = 1.0
            space.append(a)
    space = np.array(space)
    rows, entity = [], []
    fdim = dim - 1
    qdeg = 2 * k + 2
    fp, fw = ref_basis.gauss_jacobi(fdim, qdeg) if fdim > 0 else (np.zeros((1, 0)), np.ones(1))
    verts = _REF_V[dim]
    mdeg = k - 1 if family == "RT" else k
    for f, lv in enumerate(SUB[dim][fdim]):
        v0 = verts[lv[0]]
        E = np.array([verts[lv[j]] - v0 for j in range(1, dim)])
        x = v0 + fp @ E
        normal = np.array([-E[0][1], E[0][0]]) if dim == 2 else np.cross(E[0], E[1])
        val, _ = _mono_tab(exps, x)
        q = _facet_family(fdim, mdeg, fp)
        for i in range(q.shape[1]):
            # l(space_s) = sum_q w q_i (space_s(x_q) . n)
            sn = np.einsum("scm,qm,c->sq", space, val, normal)
            rows.append(sn @ (fw * q[:, i]))
            entity.append((fdim, f))
    if family == "RT" and k >= 2:
        cp, cw = ref_basis.gauss_jacobi(dim, qdeg)
        val, _ = _mono_tab(exps, cp)
        q = ref_basis.tabulate_lagrange(dim, k - 2, cp)[0, :, :, 0]
        for i in range(q.shape[1]):
            for d in range(dim):
                rows.append(np.einsum("sm,qm->sq", space[:, d, :], val) @ (cw * q[:, i]))
                entity.append((dim, 0))
    if family == "BDM" and k == 2:
        # interior moments against the Whitney edge functions lam_a grad lam_b - lam_b grad lam_a
        cp, cw = ref_basis.gauss_jacobi(dim, qdeg)
        val, _ = _mono_tab(exps, cp)
        lam = np.concatenate([1.0 - cp.sum(axis=1, keepdims=True), cp], axis=1)
        glam = np.concatenate([-np.ones((1, dim)), np.eye(dim)], axis=0)
        for a, b in SUB[dim][1]:
            w = lam[:, [a]] * glam[b][None, :] - lam[:, [b]] * glam[a][None, :]  # [nq, dim]
            rows.append(np.einsum("scm,qm,qc,q->s", space, val, w, cw))
            entity.append((dim, 0))
    elif family == "BDM" and k == 3:
        cp, cw = ref_basis.gauss_jacobi(dim, qdeg)
        val, _ = _mono_tab(exps, cp)
        for w in _nedelec2(dim, cp):
            rows.append(np.einsum("scm,qm,qc,q->s", space, val, w, cw))
            entity.append((dim, 0))
    elif family == "BDM" and k > 3:
        raise NotImplementedError("oracle BDM degree >= 4")
    D = np.array(rows)  # D[i, s]
    C = _solve_dual(D)  # phi_j = sum_s C[s, j] space_s
    coeffs = np.einsum("sj,scm->jcm", C, space)
    entity_dofs = [[[] for _ in SUB[dim][d]] for d in range(dim + 1)]
    for i, (d, e) in enumerate(entity):
        entity_dofs[d][e].append(i)
    _CACHE[key] = (exps, coeffs, entity_dofs)
    return _CACHE[key]


def tabulate(family, dim, k, pts):
    exps, coeffs, _ = _element(family, dim, k)
    pts = np.asarray(pts, dtype=float).reshape(-1, dim)
    val, grad = _mono_tab(exps, pts)
    tab = np.zeros((1 + dim, len(pts), coeffs.shape[0], dim))
    tab[0] = np.einsum("qm,jcm->qjc", val, coeffs)
    for d in range(dim):
        tab[1 + d] = np.einsum("qm,jcm->qjc", grad[d], coeffs)
    return tab


def entity_dofs(family, dim, k):
    return _element(family, dim, k)[2]


SUB_EDGES = {2: ((1, 2), (0, 2), (0, 1)), 3: ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))}


def _perm_matrix(image):
    """``P[image[j], j] = 1``: node j goes to node image[j]."""
    P = np.zeros((len(image), len(image)))
    for j, i in enumerate(image):
        P[i, j] = 1.0
    return P


def entity_transformations(family, dim, k):
    """Signed node permutations of the facet dof family (degree k-1 for RT, k for BDM) under
    interval reflection, triangle rotation (0,1,2)->(1,2,0) and triangle reflection (0,1,2)->(0,2,1)."""
    m = k - 1 if family == "RT" else k
    out = {}
    if dim == 2:
        image = {0: [0], 1: [1, 0], 2: [1, 0, 2], 3: [1, 0, 3, 2]}[m]
        out["interval"] = [-_perm_matrix(image)]
    else:
        # degree 3: nodes v0 v1 v2 | edge (1,2): 3 4 | edge (0,2): 5 6 | edge (0,1): 7 8 | centroid 9
        rot = {0: [0], 1: [1, 2, 0], 2: [1, 2, 0, 4, 5, 3], 3: [1, 2, 0, 6, 5, 8, 7, 3, 4, 9]}[m]
        ref = {0: [0], 1: [0, 2, 1], 2: [0, 2, 1, 3, 5, 4], 3: [0, 2, 1, 4, 3, 7, 8, 5, 6, 9]}[m]
        out["interval"] = [np.zeros((0, 0))]
        out["triangle"] = [_perm_matrix(rot), -_perm_matrix(ref)]
    return out


def permute_and_transform(tab, element):
    """``element``: oracle RefElement (``.cell``, ``.mesh``, ``.family``, ``.k``, ``.disc``)."""
    cell, mesh = element.cell, element.mesh
    dim = cell.dimension
    if element.disc or dim < 2:
        return tab
    if element.family == "Lagrange":
        # P_3 is the first Lagrange degree whose dof transformations are not the identity: the two interior nodes of a
        # mis-oriented edge are exchanged (transformations["interval"][0] = [[0, 1], [1, 0]]); the single face node needs
        # nothing.  Edge rule: R/src/basis/permute_and_transform.py:8-19 (2-D), :38-54 (3-D)
        if element.k != 3:
            return tab
        for i, (a, b) in enumerate(SUB_EDGES[dim]):
            edge = cell.node_tags[[a, b]]
            v_edge = mesh.cells[cell.sub_cells_ids[1][i]].node_tags
            oriented = np.all(edge == v_edge) if dim == 2 else np.all(edge == np.sort(v_edge))
            if oriented:
                continue
            d0 = dim + 1 + 2 * i
            tab[:, :, [d0, d0 + 1], :] = tab[:, :, [d0 + 1, d0], :]
        return tab
    T = entity_transformations(element.family, dim, element.k)
    edofs = entity_dofs(element.family, dim, element.k)
    if dim == 2:
        for i, (a, b) in enumerate(((1, 2), (0, 2), (0, 1))):
            edge = cell.node_tags[[a, b]]
            v_edge = mesh.cells[cell.sub_cells_ids[1][i]].node_tags
            if np.all(edge == v_edge):
                continue
            dofs = edofs[1][i]
            for d in range(tab.shape[0]):
                for c in range(tab.shape[3]):
                    tab[d, :, dofs, c] = T["interval"][0] @ tab[d, :, dofs, c]
        return tab
    bc_entities = [mesh.cells[f].material_id is not None for f in cell.sub_cells_ids[2]]
    case = {0: (0, 0), 1: (0, 1), 2: (2, 1), 3: (1, 0), 4: (2, 0), 5: (1, 1)}
    for i, fc in enumerate(((1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2))):
        face = cell.node_tags[list(fc)]
        face_ref = np.sort(face)
        oriented = bool(np.all(face_ref == face))
        rot = ref = 0
        if not oriented:
            perms = list(permutations(list(face_ref)))
            pos = [k for k, p in enumerate(perms) if p == tuple(face)][0]
            rot, ref = case[pos]
        if bc_entities[i] and i in (1, 3):
            oriented, ref = False, 1
        if oriented:
            continue
        rotate_t, reflect_t = T["triangle"]
        M = np.identity(len(rotate_t))
        for _ in range(rot):
            M = rotate_t @ M
        for _ in range(ref):
            M = reflect_t @ M
        dofs = edofs[2][i]
        for d in range(tab.shape[0]):
            for c in range(tab.shape[3]):
                tab[d, :, dofs, c] = M @ tab[d, :, dofs, c]
    # edges carry no dofs for RT/BDM on tetrahedra
    return tab
