"""Primal linear elasticity on a vector Lagrange field ``u`` -- GPU versions of the classes in
R/src/weak_forms/le_primal_weak_form.py (same names, same ``functions`` keys, same sign conventions).

* ``LEPrimalWeakForm``            R/...:13-129   r = +int phi.f + int sigma(u_h):grad phi, functions rhs/lambda/mu
* ``LEPrimalWeakFormBCDirichlet`` R/...:131-176  penalty beta = 1e12, functions u
* ``LEPrimalWeakFormBCNeumann``   R/...:179-221  traction load, functions t
* ``LEPrimalWeakFormBCNormalDirichlet`` R/...:224-286  normal-only penalty, functions u
"""
from .. import _lib
from .h1_common import H1FacetForm, H1VolumeForm


class LEPrimalWeakForm(H1VolumeForm):
    field_name = "u"
    coefficient_names = ("lambda", "mu")
    load_name = "rhs"
    load_sign = 1.0
    fused_kernel = "nfa_assemble_h1_elasticity_affine"
    p2tet_kernel = "nfa_assemble_p2tet_elasticity"

    def launch_local(self, dev, data, alpha_dev, dest_dev, Ke, re):
        lam, mu = data.coefs["lambda"], data.coefs["mu"]
        _lib.call(
            "nfa_local_h1_elasticity", data.dim, data.n_phi, data.n_q, _lib.ptr(data.tab), _lib.ptr(data.w),
            _lib.ptr(dev.coords), _lib.ptr(data.cell_nodes), data.n_items,
            _lib.ptr(lam[0]), lam[1], lam[2], _lib.ptr(mu[0]), mu[1], mu[2],
            _lib.ptr(data.f), _lib.ptr(alpha_dev), _lib.ptr(dest_dev), _lib.ptr(Ke), _lib.ptr(re), _lib.stream_ptr(),
        )


class LEPrimalWeakFormBCDirichlet(H1FacetForm):
    field_name = "u"
    data_name = "u"
    beta = 1.0e12
    r_sign = -1.0
    with_matrix = True


class LEPrimalWeakFormBCNeumann(H1FacetForm):
    field_name = "u"
    data_name = "t"
    beta = 1.0
    r_sign = 1.0
    with_matrix = False


class LEPrimalWeakFormBCNormalDirichlet(H1FacetForm):
    """Normal-displacement penalty, R/src/weak_forms/le_primal_weak_form.py:224-286: per component c the basis is
    phi n_c, so J_cc = beta n_c^2 int phi phi^T (no cross-component blocks, as coded there) and
    r_c = -beta int (u_D . n) n_c phi."""

    field_name = "u"
    data_name = "u"
    beta = 1.0e12
    r_sign = -1.0
    with_matrix = True
    normal_only = True
