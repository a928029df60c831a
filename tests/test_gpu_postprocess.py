"""l2_projector / l2_error / div_error (R/src/postprocess/projectors.py, l2_error_post_processor.py) on the GPU:
error norms against the oracle's cell-by-cell restatement for arbitrary coefficient vectors, and the reference's own
end-to-end property -- the projection of an in-space polynomial has (machine-)zero error
(R/tests/test_projectors/test_projectors.py:86-216)."""
import numpy as np
import pytest

from nfempy_b200.postprocess import (div_error, div_norm, div_scaled_error, l2_error, l2_error_projected, l2_norm,
                                     l2_projector)
from oracle import ref_postprocess
from tests.parity_helpers import tensor_mixed_case

pytestmark = pytest.mark.gpu


def _fields(dim):
    tc = 3 if dim == 3 else 1
    return {
        "s": lambda x, y, z: np.array([[x + y, y * z, 1.0 + z], [np.sin(x), y, x * z], [z, x - y, np.cos(y)]])[:dim, :dim],
        "m": lambda x, y, z: np.array([[y, x, z * x], [1.0 + x, z, y], [x * y, 2.0, z]])[:tc, :dim],
        "u": lambda x, y, z: np.array([x * y, np.sin(z) + x, y - z])[:dim],
        "t": lambda x, y, z: np.array([z + 1.0, x * x, y])[:tc],
        "div_s": lambda x, y, z: np.array([1.0 + z, np.cos(x) + 1.0 + x, 1.0 - np.sin(y)])[:dim],
        "div_m": lambda x, y, z: np.array([x, 0.0 * x, y + 1.0])[:tc],
    }


@pytest.mark.parametrize("dim,k,fam", [(3, 1, "BDM"), (2, 2, "RT"), (3, 2, "RT")])
def test_error_norms_match_the_oracle(dim, k, fam):
    case = tensor_mixed_case(kind="lce", dim=dim, s_family=fam, m_family="RT", k=k, n=2 if dim == 2 else 1)
    space, rspace = case.space(), case.ref_space()
    alpha = np.cos(np.arange(space.n_dof) * 0.31) + 0.2
    fn = _fields(dim)
    got, ref = l2_error(dim, space, fn, alpha), ref_postprocess.l2_error(dim, rspace, fn, alpha)
    assert len(got) == 4 and np.allclose(got, ref, rtol=1e-12, atol=0.0)
    got_d, ref_d = div_error(dim, space, fn, alpha), ref_postprocess.div_error(dim, rspace, fn, alpha)
    assert len(got_d) == 2 and np.allclose(got_d, ref_d, rtol=1e-12, atol=0.0)
    assert np.allclose(l2_error(dim, space, fn, alpha, skip_fields=["m", "t"]), [ref[0], ref[2]], rtol=1e-12)
    # norms of the discrete fields themselves and of the exact functions (solution_norms_post_processor.py)
    assert np.allclose(l2_error_projected(dim, space, alpha), ref_postprocess.l2_error_projected(dim, rspace, alpha), rtol=1e-12)
    shifted = np.concatenate([np.full(5, 7.0), alpha])
    assert np.allclose(l2_error_projected(dim, space, shifted, dof_shift=5),
                       ref_postprocess.l2_error_projected(dim, rspace, shifted, dof_shift=5), rtol=1e-12)
    assert np.allclose(l2_norm(dim, space, fn), ref_postprocess.l2_norm(dim, rspace, fn), rtol=1e-12)
    assert np.allclose(div_norm(dim, space, fn), ref_postprocess.div_norm(dim, rspace, fn), rtol=1e-12)
    # scaled divergence (the scaled Cosserat forms): skipped without a scale, else l div s + grad l . s
    assert div_scaled_error(dim, space, fn, alpha) == []
    fn_l = dict(fn, l=lambda x, y, z: 0.5 + 0.25 * x + 0.1 * y * y,
                grad_l=lambda x, y, z: np.array([0.25 * np.ones_like(x), 0.2 * y, 0.0 * x]))
    got_s, ref_s = div_scaled_error(dim, space, fn_l, alpha), ref_postprocess.div_scaled_error(dim, rspace, fn_l, alpha)
    assert len(got_s) == 2 and np.allclose(got_s, ref_s, rtol=1e-12, atol=0.0)


@pytest.mark.parametrize("dim", [2, 3])
def test_projection_of_an_in_space_field_has_no_error(dim):
    """Lagrange P2 (vector) and RT2 fields: a quadratic displacement and a linear flux are reproduced exactly."""
    from nfempy_b200.mesh import box_tet_mesh, square_tri_mesh
    from nfempy_b200.spaces import ProductSpace

    mesh = box_tet_mesh(2, jitter=0.2, shuffle_local_nodes=True) if dim == 3 else square_tri_mesh(3, jitter=0.2)
    space = ProductSpace({"u": (dim, dim, "Lagrange", 2, mesh), "q": (dim, 1, "RT", 2, mesh)})
    space.build_structures({"u": [1], "q": [1]}, {})
    f_u = lambda x, y, z: np.array([x * y + z, 1.0 - x * x, y * z + x])[:dim]
    f_q = lambda x, y, z: np.array([[1.0 + x - y, 2.0 * y + z * (dim - 2), x + y + 0.0 * z]])  # [n_comp, 3(, points)]
    alpha = l2_projector(space, {"u": f_u, "q": f_q})
    errs = l2_error(dim, space, {"u": f_u, "q": f_q}, alpha)
    assert max(errs) < 1e-12, errs
