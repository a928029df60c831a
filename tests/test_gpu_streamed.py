"""Streamed assembly (assembly/streamed.py + csrc/nfa_stream.cu) against the two-stage GPUAssembler and the oracle."""
import numpy as np
import pytest

from tests.parity_helpers import compare_csr, tensor_mixed_case

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _streamed(case, chunk):
    from nfempy_b200.assembly.streamed import StreamedAssembler

    space = case.space()
    alpha = np.zeros(space.n_dof) if case.alpha_fn is None else case.alpha_fn(space.n_dof)
    asm = StreamedAssembler(space, chunk_cells=chunk)
    asm.form_to_input_list = case.gpu_forms(space, alpha)
    csr, rhs = asm.assemble()
    ip, idx, vals = csr.rows(0, csr.n_dof)
    assert asm.stats["n_chunks"] == -(-asm.stats["n_cells"] // chunk)
    assert ip.dtype == np.int64  # 64-bit row pointers; these test matrices fit 32 bits, which is what compare_csr expects
    return ip.astype(np.int32), idx, vals, rhs.cpu().numpy()


@pytest.mark.parametrize("kind,dim,k,n,chunk", [("lce_scaled", 3, 1, 2, 7), ("lce", 3, 2, 1, 4), ("le", 2, 2, 3, 5),
                                                 ("lce_scaled", 2, 1, 4, 1000)])
def test_streamed_equals_two_stage(kind, dim, k, n, chunk):
    case = tensor_mixed_case(kind=kind, dim=dim, k=k, n=n)
    ref = case.run_gpu()
    got = _streamed(case, chunk)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
    compare_csr(got, ref, 1e-13)


def test_streamed_config_4_order_against_the_oracle():
    """BDM_3 x RT_3 x P_2 x P_2 (348 local dofs) on six tetrahedra, three chunks."""
    case = tensor_mixed_case(kind="lce_scaled", dim=3, s_family="BDM", m_family="RT", k=3, n=1)
    compare_csr(_streamed(case, 2), case.run_oracle(), RTOL)
