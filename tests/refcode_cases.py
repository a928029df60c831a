"""The cases pinned to the reference's own source (tools/make_reference_pins.py writes ``tests/golden/refcode_<name>.npz``,
tests/test_cpu_reference_pins.py and tests/test_gpu_zz_reference_pins.py read them).  Built with the constructors of
tests/parity_helpers.py, so product, oracle and reference code see the same mesh, spaces, coefficient functions and lists."""
from tests import parity_helpers as ph


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
        "poisson_bdm1_2d": lambda: ph.mixed_poisson_case(dim=2, k=1, n=3, family="BDM"),
        "degenerate_rt1_3d": lambda: ph.degenerate_case(kind="degenerate", dim=3, k=1, n=2),
        "degenerate_rt2_2d": lambda: ph.degenerate_case(kind="degenerate", dim=2, k=2, n=3),
        "unscaled_rt1_3d": lambda: ph.degenerate_case(kind="unscaled", dim=3, k=1, n=2),
        "unscaled_rt2_3d": lambda: ph.degenerate_case(kind="unscaled", dim=3, k=2, n=1),
        # tensor forms through ``evaluate_form_vectorized`` with alpha = 0, as the example drivers call them (that entry
        # point returns the load vector only); names ending in "_ad": the SAME classes through ``evaluate_form`` (quadrature
        # loop + AD) at a non-zero state -- the oracle restates the vectorized blocks and adds J.alpha, so these agree to
        # round-off, not bit for bit
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
        "projector_p3_2d": lambda: projector_case(2, {"u": ("Lagrange", 3, False)}),
        "projector_p3_3d": lambda: projector_case(3, {"u": ("Lagrange", 3, False)}),
        "lce_primal_3d": lambda: ph.lce_primal_case(dim=3, k=2, n=1, varying=True, alpha=True),
        "lce_primal_2d": lambda: ph.lce_primal_case(dim=2, k=2, n=2),
        "taylor_hood_3d": lambda: ph.taylor_hood_case(n=1, varying=True, alpha=True),
        "taylor_hood_3d_n3": lambda: ph.taylor_hood_case(n=3, varying=True, alpha=True),
        "stress_constraint_3d": lambda: ph.stress_constraint_case(dim=3, k=2, n=1, alpha=True),
    }
    return out
