"""Unscaled (physical-variable) degenerate elliptic mixed form -- GPU versions of the classes in
R/src/weak_forms/unscaled_elliptic_weak_form.py (same names, fields ``u``, ``p``; ``functions`` keys ``rhs``,
``porosity``, ``d_phi``, ``grad_d_phi`` (unused by the reference, :125-150) / ``p``).

Volume form (:41-153): porosity and delta are replaced by ``phi_threshold`` where they are ``isclose`` to zero or
negative (:135-139); ``J = [[int delta^-2 psi.psi, -int w div psi], [int div psi w, int phi w w]]``,
``r = J alpha - [0; int sqrt(porosity) f w]`` (the load uses the un-thresholded porosity, :117-123).
Boundary form (:156-218): plain natural condition ``int_F p_D (psi . n)``.
"""
import numpy as np

from .. import coefficients
from .hdiv_general import HdivL2GeneralForm, HdivNaturalBC


class UnscaledEllipticWeakForm(HdivL2GeneralForm):
    flux_name, scalar_name = "u", "p"

    def __init__(self, space=None):
        super().__init__(space)
        self._phi_threshold = 1.0e-16

    @property
    def phi_threshold(self):
        return self._phi_threshold

    @phi_threshold.setter
    def phi_threshold(self, value):
        if value <= 0:
            raise ValueError("phi_threshold must be positive")
        self._phi_threshold = value

    def point_coefficients(self, x, dim):
        phi_raw = coefficients.evaluate(self.functions["porosity"], x)
        delta = coefficients.evaluate(self.functions["d_phi"], x)
        f = coefficients.evaluate(self.functions["rhs"], x, 1)[:, :, 0]
        phi = np.where(np.logical_or(np.isclose(phi_raw, 0.0), phi_raw < 0.0), self._phi_threshold, phi_raw)
        delta = np.where(np.logical_or(np.isclose(delta, 0.0), delta < 0.0), self._phi_threshold, delta)
        with np.errstate(invalid="ignore"):
            load = np.sqrt(phi_raw) * f
        return {"c_uu": 1.0 / (delta * delta), "c_pp": phi, "load_p": load}


class UnscaledEllipticWeakFormBCDirichlet(HdivNaturalBC):
    flux_name, scalar_name = "u", "p"

    def boundary_values(self, x, y, z):
        return np.asarray(self.functions["p"](x, y, z), dtype=np.float64)
