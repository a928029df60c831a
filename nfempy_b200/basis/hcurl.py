"""H(curl) reference element: first-kind Nedelec of degree 1 ("N1E", the Whitney edge element; basix convention: lowest
degree is 1) on triangles and tetrahedra, restated in exact rational arithmetic like the H(div) elements of ``hdiv.py``.

* polynomial space: ``[P_0]^d + {p in [Ptilde_1]^d : p . x = 0}`` (2-D: ``(-y, x)``; 3-D: ``x cross e_j``);
* dual basis: one tangential moment per edge ``e = (a, b)`` in the sub-entity order the reference hard-codes
  (R/src/mesh/mesh.py:187-203): ``l_e(phi) = int_0^1 phi(v_a + s (v_b - v_a)) . (v_b - v_a) ds`` (non-normalised tangent);
* basis = dual basis, i.e. the Whitney functions ``lam_a grad lam_b - lam_b grad lam_a``;
* physical element: covariant Piola map ``phi = J^{-T} phi_hat`` (``covariant_piola``); reversing an edge flips the sign
  of its function (``entity_transformations``: ``interval = [[-1]]``).

Second-kind Nedelec of degree 1 ("N2E") is the full ``[P_1]^d`` with two tangential moments per edge (``build_hcurl``).

The reference names the families (R/src/basis/element_family.py:9, R/src/basis/finite_element.py:29) and drives them through
``L2ProjectorWeakForm`` in its projector test (R/tests/test_projectors/test_projectors.py:187-216); its ``evaluate_basis``
still carries a TODO for the covariant map (R/src/basis/finite_element.py:178).  Consumers here: ``DiscreteSpace``,
``L2ProjectorWeakForm`` and ``l2_error`` through ``csrc/nfa_hcurl.cu``.
"""
from fractions import Fraction

import numpy as np

from .hdiv import _compose_affine, _padd, _pmul, _simplex_integral, _solve_exact
from .reference_element import REF_VERTICES, SUB_ENTITIES, monomial_exponents


def _space_functions(dim):
    fns = []
    zero = (0,) * dim
    for c in range(dim):
        fns.append([({zero: Fraction(1)} if cc == c else {}) for cc in range(dim)])
    unit = lambda i: tuple(1 if d == i else 0 for d in range(dim))
    if dim == 2:
        fns.append([{unit(1): Fraction(-1)}, {unit(0): Fraction(1)}])
    else:
        for j in range(3):  # x cross e_j: component c = eps(c, a, j) x_a
            vec = [{} for _ in range(3)]
            for c in range(3):
                for a in range(3):
                    sgn = ((c, a, j) in ((0, 1, 2), (1, 2, 0), (2, 0, 1))) - ((c, a, j) in ((0, 2, 1), (2, 1, 0), (1, 0, 2)))
                    if sgn:
                        vec[c] = _padd(vec[c], {unit(a): Fraction(1)}, sgn)
            fns.append(vec)
    return fns


def _full_p1_functions(dim):
    """Basis of ``[P_1]^d``: every monomial of degree <= 1 in every component."""
    fns = []
    for c in range(dim):
        for e in monomial_exponents(dim, 1):
            key = tuple(int(v) for v in e)
            fns.append([({key: Fraction(1)} if cc == c else {}) for cc in range(dim)])
    return fns


def build_hcurl(el):
    """N1E of degree 1 (one tangential moment per edge) and N2E of degree 1 (the full ``[P_1]^d``, two tangential moments
    per edge against the Legendre functions ``q_0 = 1``, ``q_1 = sqrt(3) (2 s - 1)`` of the edge parameter -- basix'
    legendre variant, which makes the reversal of an edge a diagonal transformation: the even moment changes sign with the
    tangent, the odd one does not)."""
    dim, k = el.dim, el.degree
    if dim not in (2, 3) or k != 1 or el.family not in ("N1E", "N2E"):
        raise NotImplementedError("N1E / N2E are restated for degree 1 on triangles and tetrahedra")
    second_kind = el.family == "N2E"
    verts = [[Fraction(int(v)) for v in row] for row in REF_VERTICES[dim].tolist()]
    fns = _full_p1_functions(dim) if second_kind else _space_functions(dim)
    moments = [{(0,): Fraction(1)}] + ([{(0,): Fraction(-1), (1,): Fraction(2)}] if second_kind else [])  # 1, 2 s - 1
    funcs = []
    for a, b in SUB_ENTITIES[dim][1]:
        t = [verts[b][c] - verts[a][c] for c in range(dim)]
        x_of_s = [_padd({(0,): verts[a][c]}, {(1,): t[c]}) for c in range(dim)]
        for q in moments:

            def functional(vec, t=t, x_of_s=x_of_s, q=q):
                total = {}
                for c in range(dim):
                    if t[c] != 0 and vec[c]:
                        comp = _compose_affine(vec[c], x_of_s, 1)
                        total = _padd(total, {e: v * t[c] for e, v in _pmul(comp, q).items()})
                return _simplex_integral(total) if total else Fraction(0)

            funcs.append(functional)
    D = [[fn(s) for s in fns] for fn in funcs]
    Dinv = _solve_exact(D)
    el.value_size = dim
    el.exps = monomial_exponents(dim, 1)
    index = {tuple(int(v) for v in e): m for m, e in enumerate(el.exps)}
    n = len(funcs)
    coeffs = [[[Fraction(0)] * len(el.exps) for _ in range(dim)] for _ in range(n)]
    for j in range(n):
        for s, vec in enumerate(fns):
            w = Dinv[s][j]
            if w == 0:
                continue
            for c in range(dim):
                for e, v in vec[c].items():
                    coeffs[j][c][index[e]] += w * v
    el.coeffs = np.array([[[float(v) for v in comp] for comp in fn] for fn in coeffs])
    el._exact_coeffs = coeffs  # dual to the moments against 1 and 2 s - 1 (rational)
    per_edge = len(moments)
    if second_kind:
        el.coeffs[1::2] /= np.sqrt(3.0)  # dual to the ORTHONORMAL odd moment sqrt(3) (2 s - 1)
    el._entity_dofs = [[[] for _ in SUB_ENTITIES[dim][d]] for d in range(dim + 1)]
    for i in range(n):
        el._entity_dofs[1][i // per_edge].append(i)
    el._interval_T = el._triangle_T = None


def covariant_piola(values, jac):
    """``phi = J^{-T} phi_hat`` for reference values ``[..., n_phi, dim]`` and the cell Jacobian ``J[3 or dim, dim]``
    (``J^{-T}`` is the transposed Moore-Penrose inverse for cells embedded in 3-space)."""
    jinv_t = np.linalg.pinv(np.asarray(jac, dtype=float)).T  # [rows of x, dim]
    return np.einsum("xk,...k->...x", jinv_t, values)
