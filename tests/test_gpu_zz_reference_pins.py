"""The CUDA path against fixtures produced by the REFERENCE'S OWN SOURCE (``tests/golden/refcode_*.npz``, written by
``tools/make_reference_pins.py``: /root/reference/src imported unmodified over stand-ins for meshio / basix / auto_diff;
see ``tests/test_cpu_reference_pins.py`` for what that pins and what it leaves open).

Most cases are form lists the oracle-parity tests of this suite already run (same constructors, same arguments); here the
expected CSR and right-hand side come from the reference code instead of the oracle.  New combinations: the cg example
driver's tag lists under the ``filtered_upper`` policy, the continuous cubic Lagrange projector with a linear function, and
the transport driver's ten-step Newton / time loop.  (The file sorts last on purpose: written after the round's GPU budget was
spent, it has not run on a B200 yet, and ``pytest -x`` reaches every other GPU test first.)"""
import os

import numpy as np
import pytest

from tests.parity_helpers import compare_csr
from tests.refcode_cases import cases

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

# name -> the oracle-parity test that runs the same case
SAME_AS = {
    "le_p2_3d": "test_gpu_parity.py::test_elasticity_p2_3d (BASELINE config 2's form list)",
    "poisson_rt1_3d": "test_gpu_hdiv.py::test_mixed_poisson_rt_3d[1] (config 3, k = 1; row-owner path)",
    "poisson_bdm3_3d_varying": "test_gpu_hdiv.py::test_mixed_poisson_bdm[3-3]",
    "degenerate_rt1_3d": "test_gpu_hdiv.py::test_degenerate_elliptic_family[3-1-degenerate]",
    "unscaled_rt1_3d": "test_gpu_hdiv.py::test_degenerate_elliptic_family[3-1-unscaled]",
    "lce_scaled_config4_order_ad": "test_gpu_tensor_forms.py::test_config_4_order_bdm3[lce_scaled-RT] (config 4's order)",
    "sundus_rt1_3d": "__graft_entry__.smoke() / test_gpu_hdiv.py::test_sundus_time_dependent_forms[3-1] (config 5's element; row-owner path)",
    "sundus_rt1_2d": "test_gpu_hdiv.py::test_sundus_time_dependent_forms[2-1]",
    "projector_n1e_2d": "test_gpu_projector.py::test_hcurl_projector_matches_oracle[2]",
    "projector_n1e_3d": "test_gpu_projector.py::test_hcurl_projector_matches_oracle[3]",
    "projector_mixed_2d": "test_gpu_projector.py::test_projector_matches_oracle_mixed_fields[2]",
    "projector_mixed_3d": "test_gpu_projector.py::test_projector_matches_oracle_mixed_fields[3]",
    # continuous cubic Lagrange: the product scatters through exchanged destinations where the reference exchanges basis
    # functions -- the assembled systems must coincide (same space as test_gpu_cubic.py's projector test)
    "projector_p3_2d": "test_gpu_cubic.py::test_projection_of_a_cubic_onto_continuous_cubic_lagrange_is_exact[2]",
    "projector_p3_3d": "test_gpu_cubic.py::test_projection_of_a_cubic_onto_continuous_cubic_lagrange_is_exact[3]",
    "taylor_hood_3d_n3": "test_gpu_h1_multifield.py::test_taylor_hood_2d[True-True]",
}


@pytest.mark.parametrize("name", sorted(SAME_AS))
def test_cuda_path_assembles_what_the_reference_code_assembled(name):
    fx = np.load(os.path.join(GOLDEN, "refcode_%s.npz" % name))
    case = cases()[name]()
    compare_csr(case.run_gpu(), (fx["indptr"], fx["indices"], fx["data"], fx["rhs"]), 1e-12)


@pytest.mark.parametrize("name,k,seed", [("md_elliptic_k1", 1, 5), ("md_elliptic_k2", 2, None)])
def test_cuda_md_system_equals_the_reference_codes(name, k, seed):
    """The stacked mixed-dimensional system (bulk + fracture + Robin coupling, ``GPUAssembler([space_c0, space_c1])``) of
    test_gpu_md_coupling.py, expected values from the reference's forms and its example-local ``CouplingWeakForm``."""
    from tests.test_gpu_md_coupling import _systems

    fx = np.load(os.path.join(GOLDEN, "refcode_%s.npz" % name))
    got = _systems(k, alpha_seed=seed)[0]
    compare_csr(got, (fx["indptr"], fx["indices"], fx["data"], fx["rhs"]), 1e-12)


def test_cuda_filtered_upper_policy_equals_what_the_reference_driver_inserted():
    """The example drivers' scatter semantic (``GPUAssembler(pattern_policy="filtered_upper", explicit_diagonal=True)``, as
    test_gpu_parity.py::test_example_driver_scatter_policy_primal_elasticity) against the "sbaij" matrix the reference's own
    driver closure filled (R/examples/cauchy_elasticity/cg_elasticity_example_1.py:92-145, run by
    tools/make_reference_pins.py --driver): pattern bit-exact, values to 1e-12."""
    from tests.refcode_cases import driver_case

    fx = np.load(os.path.join(GOLDEN, "refcode_driver_cg_elasticity.npz"))
    got = driver_case().run_gpu(pattern_policy="filtered_upper", explicit_diagonal=True)
    compare_csr(got, (fx["indptr"], fx["indices"], fx["data"], got[3]), 1e-12)


def test_cuda_filtered_upper_policy_equals_what_the_reference_cosserat_driver_inserted():
    """BASELINE config 4's driver (mfe_cosserat_elasticity_example_3.py ``four_field_scaled_approximation``, "wc_rt" at
    k_order = 0, run by tools/make_reference_pins.py --cosserat-driver): its saddle-point matrix against
    ``GPUAssembler(pattern_policy="filtered_upper", explicit_diagonal=True)`` on the same form list -- the case of
    test_gpu_tensor_forms.py::test_example_driver_scatter_policies_cosserat[filtered_upper-True]."""
    from tests.refcode_cases import cosserat_driver_cases

    fx = np.load(os.path.join(GOLDEN, "refcode_driver_mfe_cosserat.npz"))
    case, _ = cosserat_driver_cases()
    got = case.run_gpu(pattern_policy="filtered_upper", explicit_diagonal=True)
    compare_csr(got, (fx["a_indptr"], fx["a_indices"], fx["a_data"], got[3]), 1e-12)


def test_cuda_time_loop_reproduces_the_reference_transport_driver():
    """BASELINE config 5's driver (mixed_transport_model.py ``four_fields_formulation``, run by
    tools/make_reference_pins.py --transport-driver): ten implicit-Euler steps of Newton iterations, every Jacobian and
    residual assembled by the CUDA path through the drop-in call (time-dependent form types, row-owner kernels), linear solves
    on the host -- the final coefficients must be the driver's."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl

    from nfempy_b200.assembly import GPUAssembler
    from nfempy_b200.spaces import ProductSpace
    from nfempy_b200.weak_forms import SundusDualWeakForm, SundusDualWeakFormBCDirichlet
    from tests.refcode_cases import transport_driver_problem, transport_space

    fx = np.load(os.path.join(GOLDEN, "refcode_driver_transport.npz"))
    mesh, functions, bc_functions, delta_t, t_end = transport_driver_problem()
    space = transport_space(ProductSpace, mesh)
    n = space.n_dof
    wf, bf = SundusDualWeakForm(space), SundusDualWeakFormBCDirichlet(space)
    wf.functions, bf.functions = functions, bc_functions
    mp = space.discrete_spaces["mp"]
    previous = np.zeros(n)
    for t in np.arange(delta_t, t_end + delta_t, delta_t):
        iterate = previous.copy()
        for _ in range(20):
            asm = GPUAssembler(space, None, np.zeros(n))
            asm.form_to_input_list = {
                "volume": ["time_dependent_form", range(len(mp.elements)), wf, (iterate, previous), t],
                "bc": ["time_dependent_bc_form", range(len(mp.bc_elements)), bf, (iterate, previous), t],
            }
            indptr, indices, data, res = asm.assemble()
            if np.linalg.norm(res) < 1.0e-10:
                break
            iterate = iterate + spl.spsolve(sp.csr_matrix((data, indices, indptr), shape=(n, n)).tocsc(), -res)
        else:
            raise AssertionError("Newton did not converge at t = %g" % t)
        previous = iterate
    assert np.abs(previous - fx["alpha"]).max() <= 1e-9 * np.abs(fx["alpha"]).max()
