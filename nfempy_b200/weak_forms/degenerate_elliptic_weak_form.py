"""Degenerate elliptic (Arbogast-Taicher scaled) mixed form -- GPU versions of the classes in
R/src/weak_forms/degenerate_elliptic_weak_form.py (same names, fields ``v`` (H(div)) and ``q`` (L2), ``functions``
keys ``rhs``, ``porosity``, ``d_phi``, ``grad_d_phi`` / ``q``, ``d_phi``, ``porosity``).

Volume form (:41-144), per quadrature point with ``phi = porosity``, ``delta = d_phi``:
``D_i = (delta div psi_i + grad(delta) . psi_i) / sqrt(phi)`` where ``phi`` is positive and not ``isclose`` to zero,
else ``D_i = 0`` (:125-130); ``J = [[int psi.psi, -int w D], [int D w, int w w]]``, ``r = J alpha - [0; int f w]``.
Boundary form (:147-216): ``int_F q_D' (psi . n)`` with ``q_D' = q_D delta / sqrt(phi)`` where phi is positive.
"""
import numpy as np

from .. import coefficients
from .hdiv_general import HdivL2GeneralForm, HdivNaturalBC


def _positive(phi):
    return np.logical_and(np.logical_not(np.isclose(phi, 0.0)), phi > 0.0)


class DegenerateEllipticWeakForm(HdivL2GeneralForm):
    flux_name, scalar_name = "v", "q"

    def point_coefficients(self, x, dim):
        n, nq, _ = x.shape
        phi = coefficients.evaluate(self.functions["porosity"], x)
        delta = coefficients.evaluate(self.functions["d_phi"], x)
        grad = coefficients.evaluate(self.functions["grad_d_phi"], x, dim)
        f = coefficients.evaluate(self.functions["rhs"], x, 1)[:, :, 0]
        ok = _positive(phi)
        inv_sqrt = np.where(ok, 1.0 / np.sqrt(np.where(ok, phi, 1.0)), 0.0)
        return {
            "c_uu": np.ones((n, nq)), "c_d": delta * inv_sqrt, "g": grad * inv_sqrt[:, :, None],
            "c_pp": np.ones((n, nq)), "load_p": f,
        }


class DegenerateEllipticWeakFormBCDirichlet(HdivNaturalBC):
    flux_name, scalar_name = "v", "q"

    def boundary_values(self, x, y, z):
        q_d = np.asarray(self.functions["q"](x, y, z), dtype=np.float64)
        phi = np.asarray(self.functions["porosity"](x, y, z), dtype=np.float64) * np.ones_like(x)
        delta = np.asarray(self.functions["d_phi"](x, y, z), dtype=np.float64) * np.ones_like(x)
        ok = _positive(phi)
        scale = np.where(ok, delta / np.sqrt(np.where(ok, phi, 1.0)), 1.0)
        return q_d * scale
