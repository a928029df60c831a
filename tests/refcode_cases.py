"""The cases pinned to the reference's own source (tools/make_reference_pins.py writes ``tests/golden/refcode_<name>.npz``,
tests/test_cpu_reference_pins.py and tests/test_gpu_zz_reference_pins.py read them).  Built with the constructors of
tests/parity_helpers.py, so product, oracle and reference code see the same mesh, spaces, coefficient functions and lists."""
from tests import parity_helpers as ph


def _gmsh_fixture(dim, level=0):
    from tests.golden_cases import fixture_mesh

    return fixture_mesh(dim, level)


def hcurl_projector_case(dim):
    """N1E of degree 1 through L2ProjectorWeakForm, as tests/test_gpu_projector.py::test_hcurl_projector_matches_oracle."""
    import numpy as np

    from nfempy_b200.mesh import box_tet_mesh, square_tri_mesh
    from oracle import ref_hcurl

    mesh = box_tet_mesh(2, jitter=0.2, shuffle_local_nodes=True) if dim == 3 else square_tri_mesh(3, jitter=0.2)
    functions = {"e": lambda x, y, z: np.array([[np.sin(x) + y, x * y - z, 1.0 + z * x]])}

    def gpu_forms(space, a):
        from nfempy_b200.weak_forms import L2ProjectorWeakForm

        wf = L2ProjectorWeakForm(space)
        wf.functions = functions
        return {"proj": ["form", range(len(space.discrete_spaces["e"].elements)), wf, a]}

    def ref_forms_fn(space):
        return [("form", ref_hcurl.l2_projector, functions, len(space.discrete_spaces["e"].element_ids))]

    return ph.Case(mesh, lambda g: {"e": (dim, 1, "N1E", 1, g)}, {"e": [1]}, {}, gpu_forms, ref_forms_fn)


DRIVER_MATERIAL = {"lambda": 2.5, "mu": 0.75}


def driver_displacement(m_lambda, m_mu, dim=2):
    """Takes the place of strong_solution_elasticity_example_1.displacement while the reference's driver runs."""
    import numpy as np

    return lambda x, y, z: np.array([x * (1 - x) * y, np.sin(np.pi * x) * z + m_mu * y, 0.5 * z * z + y / (1.0 + m_lambda)])[:dim]


def driver_rhs(m_lambda, m_mu, dim=2):
    import numpy as np

    return lambda x, y, z: np.array([np.sin(np.pi * x) * np.cos(np.pi * y), m_lambda * np.cos(np.pi * y) * np.sin(np.pi * z),
                                     x * y + m_mu * z])[:dim]


def driver_case():
    """What R/examples/cauchy_elasticity/cg_elasticity_example_1.py ``primal_approximation`` assembles on the jittered
    2 x 2 x 2 Kuhn box: vector-P2 elasticity, volume tags [1, 2, 3, 4], Dirichlet penalty on the boundary tags [5, 6, 7, 8]
    (three of the box's six sides carry them), through the driver's own scatter closure -- entries with
    ``~isclose(v, 1e-16)`` and ``row <= col`` only ("sbaij"), explicit zeros on the diagonal (:92-145)."""
    import numpy as np

    from nfempy_b200.mesh import box_tet_mesh
    from oracle import ref_forms

    lam, mu = DRIVER_MATERIAL["lambda"], DRIVER_MATERIAL["mu"]
    functions = {"rhs": driver_rhs(lam, mu, 3), "lambda": lambda x, y, z: lam * np.ones_like(x),
                 "mu": lambda x, y, z: mu * np.ones_like(x)}
    bc_functions = {"u": driver_displacement(lam, mu, 3)}

    def gpu_forms(space, a):
        from nfempy_b200.weak_forms import LEPrimalWeakForm, LEPrimalWeakFormBCDirichlet

        wf, bf = LEPrimalWeakForm(space), LEPrimalWeakFormBCDirichlet(space)
        wf.functions, bf.functions = functions, bc_functions
        u = space.discrete_spaces["u"]
        return {"volume": ["form", range(len(u.elements)), wf, a], "bc": ["bc_form", range(len(u.bc_elements)), bf, a]}

    def ref_forms_fn(space):
        u = space.discrete_spaces["u"]
        return [("form", ref_forms.le_primal, functions, len(u.element_ids)),
                ("bc_form", ref_forms.le_primal_bc_dirichlet, bc_functions, len(u.bc_element_ids))]

    mesh = box_tet_mesh(2, jitter=0.2, shuffle_local_nodes=True)
    return ph.Case(mesh, lambda g: {"u": (3, 3, "Lagrange", 2, g)}, {"u": [1, 2, 3, 4]}, {"u": [5, 6, 7, 8]}, gpu_forms,
                   ref_forms_fn)


def driver_pressure(m_lambda, m_mu, dim=2):
    """Takes the place of strong_solution_elasticity_example_1.pressure while the reference's driver runs."""
    import numpy as np

    return lambda x, y, z: np.array([m_lambda * (x - y) + m_mu * np.cos(z)])


def taylor_hood_driver_case(dim=2):
    """What R/examples/cauchy_elasticity/th_elasticity_example_1.py ``mixed_approximation`` assembles with its method at
    k_order = 2 (u in P_2, p in discontinuous P_0; volume tags [1, 2, 3, 4], Dirichlet penalty on [5, 6, 7, 8]) through its
    own closures: entries with ``~isclose(v, 1e-16)`` (full storage, "filtered"), explicit diagonal zeros (:99-145).  2-D:
    the reference's Taylor-Hood form refuses tetrahedra (``assert False``, le_taylor_hood_weak_form.py:133-135); one side of
    the jittered square carries a boundary tag of the driver's list."""
    import numpy as np

    from nfempy_b200.mesh import square_tri_mesh
    from oracle import ref_h1_multi

    lam, mu = DRIVER_MATERIAL["lambda"], DRIVER_MATERIAL["mu"]
    functions = {"rhs": driver_rhs(lam, mu, dim), "lambda": lambda x, y, z: lam * np.ones_like(x),
                 "mu": lambda x, y, z: mu * np.ones_like(x)}
    bc_functions = {"u": driver_displacement(lam, mu, dim), "p": driver_pressure(lam, mu, dim)}

    def gpu_forms(space, a):
        from nfempy_b200.weak_forms import LETaylorHoodWeakForm, LETaylorHoodWeakFormBCDirichlet

        wf, bf = LETaylorHoodWeakForm(space), LETaylorHoodWeakFormBCDirichlet(space)
        wf.functions, bf.functions = functions, bc_functions
        u = space.discrete_spaces["u"]
        return {"volume": ["form", range(len(u.elements)), wf, a], "bc": ["bc_form", range(len(u.bc_elements)), bf, a]}

    def ref_forms_fn(space):
        u = space.discrete_spaces["u"]
        return [("form", ref_h1_multi.le_taylor_hood, functions, len(u.element_ids)),
                ("bc_form", ref_h1_multi.le_taylor_hood_bc_dirichlet, bc_functions, len(u.bc_element_ids))]

    mesh = square_tri_mesh(3, jitter=0.2)
    return ph.Case(mesh, lambda g: {"u": (dim, dim, "Lagrange", 2, g), "p": (dim, 1, "Lagrange", 0, g)},
                   {"u": [1, 2, 3, 4], "p": [1, 2, 3, 4]}, {"u": [5, 6, 7, 8]}, gpu_forms, ref_forms_fn, disc={"p": True})


def mixed_elasticity_driver_cases():
    """(saddle-point case, Riesz-map case) of what R/examples/cauchy_elasticity/mfe_elasticity_example_1.py
    ``three_field_approximation`` assembles with its method at k_order = 0 (s in BDM_1 rows, u, t in discontinuous P_0) on the
    jittered Kuhn box of six tetrahedra: constant Lame parameters from its material table, volume tags [1, 2, 3, 4], natural
    boundary form on the tags [5, 6, 7, 8]; ``A`` from LEDualWeakForm through the driver's closure (filtered, upper, explicit
    diagonal zeros), ``P`` from LERieszMapWeakForm (filtered, upper)."""
    import numpy as np

    from nfempy_b200.mesh import box_tet_mesh

    mesh = box_tet_mesh(1, jitter=0.2, shuffle_local_nodes=True)
    out = []
    for kind in ("le", "le_riesz"):
        case = ph.tensor_mixed_case(kind=kind, dim=3, n=1, alpha=False, mesh=mesh)
        case.tags = {name: [1, 2, 3, 4] for name in ("s", "u", "t")}
        case.btags = {"s": [5, 6, 7, 8]}
        functions = case.ref_forms_fn(case.ref_space())[0][2]  # the dict both form lists of the case share
        functions["lambda"] = lambda x, y, z: DRIVER_MATERIAL["lambda"] * np.ones_like(x)
        functions["mu"] = lambda x, y, z: DRIVER_MATERIAL["mu"] * np.ones_like(x)
        out.append(case)
    return tuple(out)


def cosserat_driver_cases():
    """(saddle-point case, Riesz-map case) of what R/examples/cosserat_elasticity/mfe_cosserat_elasticity_example_3.py
    ``four_field_scaled_approximation`` assembles with its method "wc_rt" at k_order = 0 (BDM_1 x RT_1 x P_0 x P_0) on the
    jittered Kuhn box of six tetrahedra: ``A`` from LCEScaledDualWeakForm through the driver's closure (filtered, upper,
    explicit diagonal zeros; the natural boundary form only feeds the right-hand side) and the preconditioner ``P`` from
    LCEScaledRieszMapWeakForm (filtered, upper).  Coefficients: tests/parity_helpers.tensor_mixed_case's."""
    from nfempy_b200.mesh import box_tet_mesh

    mesh = box_tet_mesh(1, jitter=0.2, shuffle_local_nodes=True)
    return (ph.tensor_mixed_case(kind="lce_scaled", dim=3, n=1, alpha=False, mesh=mesh),
            ph.tensor_mixed_case(kind="lce_scaled_riesz", dim=3, n=1, alpha=False, mesh=mesh))


def transport_driver_problem():
    """Mesh, coefficient / source functions and boundary data of what
    R/examples/parabolic_diffusive_problem/mixed_transport_model.py ``four_fields_formulation`` solves in 3-D (:104-216):
    manufactured p = c = t b(x) b(y) b(z), b(s) = s (1 - s), fluxes -grad p, unit kappa / delta / eta, delta_t = 0.1 up to
    t = 1, on the jittered Kuhn box of six tetrahedra.  Returns ``(mesh, functions, bc_functions, delta_t, t_end)``."""
    import numpy as np

    from nfempy_b200.mesh import box_tet_mesh

    b, db = (lambda s: s * (1.0 - s)), (lambda s: 1.0 - 2.0 * s)
    scalar = lambda x, y, z, t: np.array([t * b(x) * b(y) * b(z)])  # noqa: E731
    flux = lambda x, y, z, t: np.array([[-t * db(x) * b(y) * b(z), -t * b(x) * db(y) * b(z), -t * b(x) * b(y) * db(z)]])  # noqa: E731
    source = lambda x, y, z, t: np.array([2.0 * t * (b(x) * b(y) + b(x) * b(z) + b(y) * b(z)) + b(x) * b(y) * b(z)])  # noqa: E731
    delta_t, t_end, eta = 0.1, 1.0, 1.0
    functions = {"rhs_f": source, "rhs_r": lambda x, y, z, t: source(x, y, z, t) + (1.0 + eta * scalar(x, y, z, t) ** 2),
                 "kappa": lambda x, y, z: 1.0, "delta": lambda x, y, z: 1.0, "eta": lambda x, y, z: eta, "delta_t": delta_t}
    bc_functions = {"mp": flux, "p": scalar, "mc": flux, "c": scalar}
    return box_tet_mesh(1, jitter=0.2, shuffle_local_nodes=True), functions, bc_functions, delta_t, t_end


def transport_space(space_class, mesh):
    space = space_class({"mp": (3, 1, "RT", 1, mesh), "mc": (3, 1, "RT", 1, mesh), "p": (3, 1, "Lagrange", 0, mesh),
                         "c": (3, 1, "Lagrange", 0, mesh)})
    space.make_subspaces_discontinuous({"p": True, "c": True})
    sides = [2, 3, 4, 5, 6, 7]
    space.build_structures({f: [1] for f in ("mp", "mc", "p", "c")}, {"mp": sides, "mc": sides})
    return space


def projector_case(dim, fields):
    from tests.test_gpu_projector import _case

    return _case(dim, fields)


def cases():
    """name -> zero-argument function returning a tests.parity_helpers.Case (the same constructors the GPU parity tests
    use, so product, oracle and reference code see the same mesh, spaces, coefficient functions and form lists)."""
    out = {
        "le_p2_3d": lambda: ph.elasticity_case(n=2, with_bc=True, k=2),
        "le_p2_3d_varying_alpha": lambda: ph.elasticity_case(n=1, with_bc=True, varying=True, alpha=True, k=2),
        "le_p1_2d_neumann": lambda: ph.elasticity_case(n=3, with_bc=True, k=1, dim=2, neumann=True),
        "le_p2_3d_normal": lambda: ph.elasticity_case(n=1, with_bc=True, k=2, normal_bc=True),
        "poisson_rt1_3d": lambda: ph.mixed_poisson_case(dim=3, k=1, n=2),
        "poisson_rt2_3d": lambda: ph.mixed_poisson_case(dim=3, k=2, n=1, varying=True),
        "poisson_rt3_2d": lambda: ph.mixed_poisson_case(dim=2, k=3, n=2),
        "poisson_bdm2_3d": lambda: ph.mixed_poisson_case(dim=3, k=2, n=1, family="BDM"),
        "poisson_bdm3_3d": lambda: ph.mixed_poisson_case(dim=3, k=3, n=1, family="BDM"),
        "poisson_bdm3_3d_varying": lambda: ph.mixed_poisson_case(dim=3, k=3, n=1, family="BDM", varying=True),
        # on the reference's own unstructured gmsh fixture (example_3_3d_l_0.msh, 298 tets: arbitrary face orientations)
        "poisson_rt2_gmsh3d": lambda: ph.mixed_poisson_case(dim=3, k=2, varying=True, mesh=_gmsh_fixture(3)),
        "sundus_rt1_gmsh3d": lambda: ph.SundusCase(dim=3, k=1, mesh=_gmsh_fixture(3)),
        "poisson_bdm1_2d": lambda: ph.mixed_poisson_case(dim=2, k=1, n=3, family="BDM"),
        "degenerate_rt1_3d": lambda: ph.degenerate_case(kind="degenerate", dim=3, k=1, n=2),
        "degenerate_rt2_2d": lambda: ph.degenerate_case(kind="degenerate", dim=2, k=2, n=3),
        "unscaled_rt1_3d": lambda: ph.degenerate_case(kind="unscaled", dim=3, k=1, n=2),
        "unscaled_rt2_3d": lambda: ph.degenerate_case(kind="unscaled", dim=3, k=2, n=1),
        # tensor forms through ``evaluate_form_vectorized`` with alpha = 0, as the example drivers call them (that entry
        # point returns the load vector only); names ending in "_ad": the SAME classes through ``evaluate_form`` (quadrature
        # loop + AD) at a non-zero state -- the oracle restates the vectorized blocks and adds J.alpha, so these agree to
        # round-off, not bit for bit
        "scaled_to_physical_projection_3d": lambda: ph.degenerate_case(kind="projection", dim=3, k=1, n=2),
        "scaled_to_physical_projection_2d": lambda: ph.degenerate_case(kind="projection", dim=2, k=2, n=3),
        "le_dual_bdm1_3d": lambda: ph.tensor_mixed_case(kind="le", dim=3, k=1, n=1, alpha=False),
        "lce_dual_bdm1_rt1_3d": lambda: ph.tensor_mixed_case(kind="lce", dim=3, k=1, n=1, alpha=False),
        "lce_dual_bdm2_2d_neumann": lambda: ph.tensor_mixed_case(kind="lce", dim=2, s_family="BDM", m_family="BDM", k=2, n=2, neumann=True, alpha=False),
        "lce_scaled_bdm1_rt1_2d": lambda: ph.tensor_mixed_case(kind="lce_scaled", dim=2, k=1, n=2, alpha=False),
        "le_riesz_3d": lambda: ph.tensor_mixed_case(kind="le_riesz", dim=3, k=1, n=1, alpha=False),
        "lce_riesz_3d": lambda: ph.tensor_mixed_case(kind="lce_riesz", dim=3, k=1, n=1, alpha=False),
        "lce_scaled_riesz_2d": lambda: ph.tensor_mixed_case(kind="lce_scaled_riesz", dim=2, k=2, n=2, alpha=False),
        "lce_dual_bdm1_rt1_3d_ad": lambda: ph.tensor_mixed_case(kind="lce", dim=3, k=1, n=1),
        "lce_scaled_bdm1_rt1_2d_ad": lambda: ph.tensor_mixed_case(kind="lce_scaled", dim=2, k=1, n=2),
        "lce_scaled_bdm1_rt1_3d_ad": lambda: ph.tensor_mixed_case(kind="lce_scaled", dim=3, k=1, n=1),
        "lce_scaled_config4_order_ad": lambda: ph.tensor_mixed_case(kind="lce_scaled", dim=3, s_family="BDM", m_family="RT", k=3, n=1),
        # config 5's element and form (two-flux transport, implicit Euler, nonlinear reaction; residual and tangent at a
        # non-zero iterate through the time-dependent scatter); the k = 2 case takes the two-stage CUDA path
        "sundus_rt1_3d": lambda: ph.SundusCase(dim=3, k=1, n=2),
        "sundus_rt1_2d": lambda: ph.SundusCase(dim=2, k=1, n=3),
        "sundus_rt2_3d": lambda: ph.SundusCase(dim=3, k=2, n=1),
        # L2ProjectorWeakForm (the only path the reference's own tests exercise); the continuous cubic Lagrange cases go
        # through the reference's edge exchange (permute_and_transform with transformations["interval"][0])
        "projector_mixed_2d": lambda: projector_case(2, {"u": ("RT", 2, False), "p": ("Lagrange", 1, True), "q": ("Lagrange", 2, False)}),
        "projector_mixed_3d": lambda: projector_case(3, {"u": ("RT", 2, False), "p": ("Lagrange", 1, True), "q": ("Lagrange", 2, False)}),
        "projector_bdm1_rt3_3d": lambda: projector_case(3, {"u": ("BDM", 1, False), "v": ("RT", 3, False)}),
        # H(curl): N1E of degree 1; the stand-in tabulates the textbook Whitney functions on the reference simplex, the
        # reference's FiniteElement maps them (covariant) and its permute_and_transform applies the edge signs; the oracle
        # builds the same functions in physical coordinates
        "projector_n1e_2d": lambda: hcurl_projector_case(2),
        "projector_n1e_3d": lambda: hcurl_projector_case(3),
        "projector_p3_2d": lambda: projector_case(2, {"u": ("Lagrange", 3, False)}),
        "projector_p3_3d": lambda: projector_case(3, {"u": ("Lagrange", 3, False)}),
        "lce_primal_3d": lambda: ph.lce_primal_case(dim=3, k=2, n=1, varying=True, alpha=True),
        "lce_primal_2d": lambda: ph.lce_primal_case(dim=2, k=2, n=2),
        "taylor_hood_3d": lambda: ph.taylor_hood_case(n=1, varying=True, alpha=True),
        "taylor_hood_3d_n3": lambda: ph.taylor_hood_case(n=3, varying=True, alpha=True),
        "stress_constraint_3d": lambda: ph.stress_constraint_case(dim=3, k=2, n=1, alpha=True),
    }
    return out
