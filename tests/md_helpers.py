"""Shared set-up of the mixed-dimensional tests: the fractured Kuhn box (tetrahedra + fracture triangles on z = 1/2) with a
mixed space on each dimension, described once for the product and once for the oracle."""
import numpy as np

from nfempy_b200.mesh import fractured_box_tet_mesh

OUTER = [2, 3, 4, 5, 6, 7]
KAPPA_C0 = lambda x, y, z: 1.0 + 0.5 * x  # noqa: E731
KAPPA_C1 = lambda x, y, z: 0.3 + 0.1 * y  # noqa: E731
COUPLING = {"kappa_normal": lambda x, y, z: 2.0 + x, "delta": lambda x, y, z: 0.1 * (1.0 + y)}
P_EXACT = lambda x, y, z: np.array([np.sin(x) + y * z])  # noqa: E731
RHS = lambda x, y, z: np.array([1.0 + x * y - z])  # noqa: E731


def md_spaces(mesh, space_class, k=1):
    s0 = space_class({"u": (3, 1, "RT", k, mesh), "p": (3, 1, "Lagrange", k - 1, mesh)})
    s0.make_subspaces_discontinuous({"p": True})
    s0.build_structures({"u": [1], "p": [1]}, {"u": OUTER + [50]})
    s1 = space_class({"u": (2, 1, "RT", k, mesh), "p": (2, 1, "Lagrange", k - 1, mesh)})
    s1.make_subspaces_discontinuous({"p": True})
    s1.build_structures({"u": [10], "p": [10]}, {"u": [30]})
    s1.dof_shift = s0.n_dof  # R/examples/md_elliptic_scalar/md_elliptic_scalar.py:205-209
    return [s0, s1]


def md_problem(n=2, k=1, jitter=0.15):
    from nfempy_b200.spaces import ProductSpace
    from nfempy_b200.weak_forms import interface_triples
    from oracle.ref_mesh import RefMesh
    from oracle.ref_spaces import RefProductSpace

    mesh, interface = fractured_box_tet_mesh(n, jitter=jitter)
    spaces = md_spaces(mesh, ProductSpace, k)
    ref_mesh = RefMesh(3, mesh.points, mesh._blocks)
    ref_mesh.build_conformal_mesh()
    ref_spaces = md_spaces(ref_mesh, RefProductSpace, k)
    triples = interface_triples(spaces[0], spaces[1], interface)
    u0 = spaces[0].discrete_spaces["u"]
    # the outer boundary facets only: the clones along the fracture carry the coupling, not a boundary condition
    # (R/examples/md_elliptic_scalar/md_elliptic_scalar.py:309-318)
    outer = [i for i, cid in enumerate(u0.bc_element_ids) if mesh.material[cid] != 50]
    return mesh, spaces, ref_spaces, triples, outer
