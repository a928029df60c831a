"""H(div) elements on simplices: Raviart-Thomas (degree 1..3, lowest = 1 as in basix) and BDM degree 1..3, restated
from their published definition in the conventions of basix (un-vendored, unpinned dependency of the reference:
R/requirements.txt:2; called at R/src/basis/finite_element.py:85-91 with ``lagrange_variant=gll_centroid``).

Construction (Ciarlet triple, exact rational arithmetic, then one rounding to double):

* polynomial space: RT_k = [P_{k-1}]^d + x * Ptilde_{k-1};  BDM_k = [P_k]^d;
* dual basis, facet part: for facet f with reference vertices (v0, v1[, v2]) in the sub-entity order the reference
  hard-codes (R/src/mesh/mesh.py:187-203), ``l(phi) = int_{ref facet} phi(x(s)) . n_f q_i(s) ds`` with the
  NON-normalised normal ``n_f = (-t_y, t_x)`` (2-D, t = v1 - v0) or ``(v1 - v0) x (v2 - v0)`` (3-D) and ``q_i`` the
  nodal Lagrange basis of degree k-1 (RT) / k (BDM) on the reference facet;
* dual basis, interior part (RT, k >= 2): ``l(phi) = int_cell phi_d q_i`` against the nodal Lagrange basis of degree
  k-2, dof index ``i * dim + d`` (basis function major, direction minor);
* dual basis, interior part (BDM_2): ``l(phi) = int_cell phi . w_e`` against the lowest-order Nedelec functions
  ``w_e = lam_a grad lam_b - lam_b grad lam_a``, one per cell edge (a, b) in the sub-entity order;
* dual basis, interior part (BDM_3): moments against a basis of the degree-2 Nedelec space (``nedelec2_test_functions``);
  the facet moments of BDM_3 use the equispaced cubic nodal family (``cubic_lattice``);
* basis = dual basis: ``l_i(phi_j) = delta_ij``.

PARITY UNPINNED: these conventions are written from basix's documented construction, not checked against a basix
binary (none is installable here).  What the repository's tests do pin: the space (degrees of freedom are
unisolvent, divergence theorem, normal-moment biorthogonality) and H(div)-conformity of the assembled global basis
under the reference's orientation table (R/src/basis/permute_and_transform.py:57-140).
"""
import itertools
from fractions import Fraction
from math import factorial

import numpy as np

from .reference_element import REF_VERTICES, SUB_ENTITIES, monomial_exponents


# ---- tiny exact polynomial algebra: {exponent tuple: Fraction} ------------------------------------------
def _pmul(p, q):
    out = {}
    for ea, ca in p.items():
        for eb, cb in q.items():
            e = tuple(x + y for x, y in zip(ea, eb))
            out[e] = out.get(e, 0) + ca * cb
    return out


def _padd(p, q, s=1):
    out = dict(p)
    for e, c in q.items():
        out[e] = out.get(e, 0) + s * c
    return out


def _ppow(p, n, nvar):
    out = {(0,) * nvar: Fraction(1)}
    for _ in range(n):
        out = _pmul(out, p)
    return out


def _simplex_integral(p):
    """Integral of a polynomial over the reference simplex of its variable count."""
    total = Fraction(0)
    for e, c in p.items():
        num = 1
        for a in e:
            num *= factorial(a)
        total += c * Fraction(num, factorial(sum(e) + len(e)))
    return total


def _lagrange_nodal_polys(dim, degree):
    """Nodal Lagrange basis (equispaced; identical to gll_centroid for degree <= 2) on the reference simplex of
    dimension ``dim`` as exact polynomials, basix dof order (vertices, edges, faces, interior)."""
    if degree > 3:
        raise NotImplementedError("facet/interior moment spaces above degree 3 (RT > 4, BDM > 3) are not restated")
    if degree == 3 and dim > 0:
        return _nodal_polys_from_nodes(dim, 3, cubic_lattice(dim))
    if dim == 0:
        return [{(): Fraction(1)}]
    nvar = dim
    one = {(0,) * nvar: Fraction(1)}
    lam = [dict(one)]
    for d in range(dim):
        e = tuple(1 if i == d else 0 for i in range(dim))
        lam.append({e: Fraction(1)})
        lam[0] = _padd(lam[0], {e: Fraction(1)}, -1)
    if degree == 0:
        return [one]
    if degree == 1:
        return lam
    out = [_pmul(l, _padd({k: 2 * v for k, v in l.items()}, one, -1)) for l in lam]  # l (2 l - 1)
    for a, b in SUB_ENTITIES[dim][1]:
        out.append({k: 4 * v for k, v in _pmul(lam[a], lam[b]).items()})
    return out


def cubic_lattice(dim):
    """Equispaced cubic lattice on the reference interval / triangle in the dof order vertices, edges (two nodes per
    edge at 1/3 and 2/3 from the edge's first vertex, edges in the sub-entity order), interior (triangle centroid).
    basix's ``gll_centroid`` variant moves the edge nodes to Gauss-Lobatto points (irrational); the moment SPACE is the
    same, the individual facet dofs of BDM_3 differ by a fixed invertible map -- part of the unpinned conventions."""
    verts = [[Fraction(int(v)) for v in row] for row in REF_VERTICES[dim].tolist()]
    nodes = [v[:] for v in verts]
    for a, b in SUB_ENTITIES[dim][1]:
        for t in (Fraction(1, 3), Fraction(2, 3)):
            nodes.append([(1 - t) * verts[a][c] + t * verts[b][c] for c in range(dim)])
    if dim == 2:
        nodes.append([Fraction(1, 3), Fraction(1, 3)])
    elif dim != 1:
        raise NotImplementedError("cubic moment family on cells of dimension %d" % dim)
    return nodes


def _nodal_polys_from_nodes(dim, degree, nodes):
    """Exact nodal basis of P_degree for the given (rational) nodes: inverse of the Vandermonde matrix."""
    exps = [tuple(int(v) for v in e) for e in monomial_exponents(dim, degree)]
    if len(exps) != len(nodes):
        raise AssertionError("node count does not match the polynomial space")
    V = []
    for x in nodes:
        row = []
        for e in exps:
            v = Fraction(1)
            for c, p in enumerate(e):
                v *= x[c] ** p
            row.append(v)
        V.append(row)
    Vinv = _solve_exact(V)  # column i = monomial coefficients of the nodal function of node i
    return [{e: Vinv[m][i] for m, e in enumerate(exps) if Vinv[m][i] != 0} for i in range(len(nodes))]


def nedelec2_test_functions(dim):
    """A basis of the first-kind Nedelec space of degree 2, ``[P_1]^d + {p in [Ptilde_2]^d : p . x = 0}``, as exact vector
    polynomials: the 2-D rotated / 3-D cross products ``x x (x_i e_j)`` span the second summand (in 3-D they satisfy one
    relation, ``x x x = 0``: the pair (z, e_z) is left out).  BDM_3's interior dofs are the moments against them."""
    fns = []
    mono = lambda *e: {tuple(e): Fraction(1)}
    for c in range(dim):
        for e in monomial_exponents(dim, 1):
            fns.append([({tuple(int(v) for v in e): Fraction(1)} if cc == c else {}) for cc in range(dim)])
    if dim == 2:
        for e in ((1, 0), (0, 1)):  # (-y, x) * x, (-y, x) * y
            m = mono(*e)
            fns.append([{k: -v for k, v in _pmul(m, mono(0, 1)).items()}, _pmul(m, mono(1, 0))])
    else:
        unit = lambda i: tuple(1 if d == i else 0 for d in range(3))
        for i in range(3):
            for j in range(3):
                if (i, j) == (2, 2):
                    continue
                # x x (x_i e_j): component c = eps(c, a, j) x_a x_i
                vec = [{} for _ in range(3)]
                for c in range(3):
                    for a in range(3):
                        sgn = ((c, a, j) in ((0, 1, 2), (1, 2, 0), (2, 0, 1))) - ((c, a, j) in ((0, 2, 1), (2, 1, 0), (1, 0, 2)))
                        if sgn:
                            vec[c] = _padd(vec[c], _pmul(mono(*unit(a)), mono(*unit(i))), sgn)
                fns.append(vec)
    return fns


def _space_functions(family, dim, k):
    """Spanning set of the polynomial space: list of vector polynomials [dim] of exact polys in x."""
    zero = {}
    fns = []
    base_deg = k - 1 if family == "RT" else k
    for c in range(dim):
        for e in monomial_exponents(dim, base_deg):
            fns.append([({tuple(int(v) for v in e): Fraction(1)} if cc == c else zero) for cc in range(dim)])
    if family == "RT":
        for e in monomial_exponents(dim, k - 1):
            if sum(e) != k - 1:
                continue
            m = {tuple(int(v) for v in e): Fraction(1)}
            fns.append([_pmul(m, {tuple(1 if i == c else 0 for i in range(dim)): Fraction(1)}) for c in range(dim)])
    return fns


def _compose_affine(poly, x_of_s, nvar_s):
    """poly(x) with x_c = affine polynomial in the facet parameters s."""
    out = {}
    for e, c in poly.items():
        term = {(0,) * nvar_s: Fraction(c)}
        for d, p in enumerate(e):
            if p:
                term = _pmul(term, _ppow(x_of_s[d], p, nvar_s))
        out = _padd(out, term)
    return out


def _functionals(family, dim, k):
    """List of (entity_dim, entity_index, functional) with functional(vector poly) -> Fraction."""
    verts = [[Fraction(int(v)) for v in row] for row in REF_VERTICES[dim].tolist()]
    out = []
    fdim = dim - 1
    q_facet = _lagrange_nodal_polys(fdim, k - 1 if family == "RT" else k)
    for f, lv in enumerate(SUB_ENTITIES[dim][fdim]):
        v0 = verts[lv[0]]
        edges = [[verts[lv[j]][c] - v0[c] for c in range(dim)] for j in range(1, dim)]
        if dim == 2:
            t = edges[0]
            normal = [-t[1], t[0]]
        else:
            a, b = edges
            normal = [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]
        x_of_s = []
        for c in range(dim):
            p = {(0,) * fdim: v0[c]}
            for j in range(fdim):
                e = tuple(1 if i == j else 0 for i in range(fdim))
                p = _padd(p, {e: edges[j][c]})
            x_of_s.append(p)
        for q in q_facet:

            def functional(vec, q=q, normal=normal, x_of_s=x_of_s):
                total = {}
                for c in range(dim):
                    if normal[c] != 0 and vec[c]:
                        comp = _compose_affine(vec[c], x_of_s, fdim)
                        total = _padd(total, {e: v * normal[c] for e, v in comp.items()})
                return _simplex_integral(_pmul(total, q)) if total else Fraction(0)

            out.append((fdim, f, functional))
    if family == "RT" and k >= 2:
        for q in _lagrange_nodal_polys(dim, k - 2):
            for d in range(dim):

                def functional(vec, q=q, d=d):
                    return _simplex_integral(_pmul(vec[d], q)) if vec[d] else Fraction(0)

                out.append((dim, 0, functional))
    if family == "BDM" and k == 2:
        # interior moments against the lowest-order Nedelec (first kind) functions lam_i grad lam_j - lam_j grad lam_i,
        # one per edge of the cell in the sub-entity order
        one = {(0,) * dim: Fraction(1)}
        lam = [dict(one)]
        grad = [[Fraction(-1)] * dim]
        for d in range(dim):
            e = tuple(1 if i == d else 0 for i in range(dim))
            lam.append({e: Fraction(1)})
            lam[0] = _padd(lam[0], {e: Fraction(1)}, -1)
            grad.append([Fraction(int(c == d)) for c in range(dim)])
        for a, b in SUB_ENTITIES[dim][1]:
            w = [_padd({e: v * grad[b][c] for e, v in lam[a].items()}, {e: v * grad[a][c] for e, v in lam[b].items()}, -1)
                 for c in range(dim)]

            def functional(vec, w=w):
                total = Fraction(0)
                for c in range(dim):
                    if vec[c]:
                        total += _simplex_integral(_pmul(vec[c], w[c]))
                return total

            out.append((dim, 0, functional))
    elif family == "BDM" and k == 3:
        for w in nedelec2_test_functions(dim):

            def functional(vec, w=w):
                total = Fraction(0)
                for c in range(dim):
                    if vec[c] and w[c]:
                        total += _simplex_integral(_pmul(vec[c], w[c]))
                return total

            out.append((dim, 0, functional))
    elif family == "BDM" and k > 3:
        raise NotImplementedError("BDM degree >= 4 is not restated")
    return out


def _solve_exact(D):
    """Inverse of a square Fraction matrix by Gauss-Jordan."""
    n = len(D)
    A = [row[:] + [Fraction(int(i == j)) for j in range(n)] for i, row in enumerate(D)]
    for col in range(n):
        piv = next(r for r in range(col, n) if A[r][col] != 0)
        A[col], A[piv] = A[piv], A[col]
        inv = 1 / A[col][col]
        A[col] = [v * inv for v in A[col]]
        for r in range(n):
            if r != col and A[r][col] != 0:
                f = A[r][col]
                A[r] = [a - f * b for a, b in zip(A[r], A[col])]
    return [row[n:] for row in A]


def build_hdiv(el):
    family, dim, k = el.family, el.dim, el.degree
    if dim not in (2, 3):
        raise NotImplementedError("H(div) elements live on triangles and tetrahedra")
    if family == "RT" and not 1 <= k <= 3:
        raise NotImplementedError("RT degree %d is not restated (1..3 are)" % k)
    if family == "BDM" and k not in (1, 2, 3):
        raise NotImplementedError("BDM degree %d is not restated (1, 2 and 3 are)" % k)
    fns = _space_functions(family, dim, k)
    funcs = _functionals(family, dim, k)
    if len(fns) != len(funcs):
        raise AssertionError("H(div) element is not square: %d functions, %d functionals" % (len(fns), len(funcs)))
    D = [[fn(s) for s in fns] for (_, _, fn) in funcs]  # D[i][s] = l_i(space fn s)
    Dinv = _solve_exact(D)  # phi_j = sum_s Dinv[s][j] space_s
    max_deg = k
    el.value_size = dim
    el.exps = monomial_exponents(dim, max_deg)
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
    el._exact_coeffs = coeffs
    el._entity_dofs = [[[] for _ in SUB_ENTITIES[dim][d]] for d in range(dim + 1)]
    for i, (d, e, _) in enumerate(funcs):
        el._entity_dofs[d][e].append(i)
    el._interval_T = el._triangle_T = None
