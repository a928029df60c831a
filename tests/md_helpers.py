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


# ---- config 5 in full: the two-flux transport model (Sundus) on the fractured box, fracture included ---------------------
SUNDUS = {
    "rhs_f": lambda x, y, z, t: np.array([np.sin(x + t) + y]), "rhs_r": lambda x, y, z, t: np.array([x * y + t * z]),
    "kappa": lambda x, y, z: 2.0 + x, "delta": lambda x, y, z: 1.5, "eta": lambda x, y, z: 0.7 + 0.1 * y, "delta_t": 0.1,
}
SUNDUS_BC = {"p": lambda x, y, z, t: np.array([x + t]), "c": lambda x, y, z, t: np.array([y * z - t])}
FIELDS = ("mp", "mc", "p", "c")


def md_sundus_spaces(mesh, space_class, k=1):
    out = []
    for dim, tag, btags in ((3, 1, OUTER + [50]), (2, 10, [30])):
        s = space_class({"mp": (dim, 1, "RT", k, mesh), "mc": (dim, 1, "RT", k, mesh),
                         "p": (dim, 1, "Lagrange", k - 1, mesh), "c": (dim, 1, "Lagrange", k - 1, mesh)})
        s.make_subspaces_discontinuous({"p": True, "c": True})
        s.build_structures({n: [tag] for n in FIELDS}, {"mp": btags, "mc": btags})
        out.append(s)
    out[1].dof_shift = out[0].n_dof
    return out


def md_sundus_forms(spaces, triples, outer, alphas, t):
    """``form_to_input_list`` of one Newton iteration of the mixed-dimensional transport step: time-dependent volume and
    boundary forms on both dimensions + the Robin coupling of each flux with the fracture's matching potential."""
    from nfempy_b200.weak_forms import CouplingWeakForm, SundusDualWeakForm, SundusDualWeakFormBCDirichlet

    forms = {}
    for tag, space in (("c0", spaces[0]), ("c1", spaces[1])):
        wf, bf = SundusDualWeakForm(space), SundusDualWeakFormBCDirichlet(space)
        wf.functions, bf.functions = SUNDUS, SUNDUS_BC
        mp = space.discrete_spaces["mp"]
        forms["vol_" + tag] = ["time_dependent_form", range(len(mp.elements)), wf, alphas, t]
        forms["bc_" + tag] = ["time_dependent_bc_form", outer if tag == "c0" else range(len(mp.bc_elements)), bf, alphas, t]
    for flux, pot in (("mp", "p"), ("mc", "c")):
        cpl = type("Coupling_%s_%s" % (flux, pot), (CouplingWeakForm,), {"flux_name": flux, "potential_name": pot})(spaces)
        cpl.functions = COUPLING
        forms["coupling_" + flux] = ["interface_form", triples, cpl, alphas[0]]
    return forms
