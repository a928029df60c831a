"""Primal Laplace/Poisson on a scalar Lagrange field ``p`` -- GPU versions of the classes in
R/src/weak_forms/laplace_primal_weak_form.py (with the one-token repair ``quadrature[dim]``).

* ``LaplacePrimalWeakForm``            R/...:10-74    r = -int phi f + int kappa grad u_h . grad phi
* ``LaplacePrimalWeakFormBCDirichlet`` R/...:77-157   penalty beta = 1e12 on the trace of the neighbour's basis,
  which for Lagrange elements is the facet's own basis in the facet's vertex order.
"""
from .. import _lib
from .h1_common import H1FacetForm, H1VolumeForm


class LaplacePrimalWeakForm(H1VolumeForm):
    field_name = "p"
    coefficient_names = ("kappa",)
    load_name = "rhs"
    load_sign = -1.0
    fused_kernel = "nfa_assemble_h1_laplace_affine"

    def launch_local(self, dev, data, alpha_dev, dest_dev, Ke, re):
        kappa = data.coefs["kappa"]
        _lib.call(
            "nfa_local_h1_laplace", data.dim, data.n_phi, data.n_q, _lib.ptr(data.tab), _lib.ptr(data.w),
            _lib.ptr(dev.coords), _lib.ptr(data.cell_nodes), data.n_items,
            _lib.ptr(kappa[0]), kappa[1], kappa[2],
            _lib.ptr(data.f), _lib.ptr(alpha_dev), _lib.ptr(dest_dev), _lib.ptr(Ke), _lib.ptr(re), _lib.stream_ptr(),
        )


class LaplacePrimalWeakFormBCDirichlet(H1FacetForm):
    field_name = "p"
    data_name = "p"
    beta = 1.0e12
    r_sign = -1.0
    with_matrix = True
