"""``ScaledToPhysicalL2Projection`` -- GPU version of R/src/weak_forms/scaled_to_physical_l2_projection.py:27-105:
L2 projection of the scaled unknowns ``(v, q)`` of the degenerate elliptic problem back to physical variables,
``J = diag(int psi.psi, int w w)``, ``r = J alpha - [int s_u (v_h . psi); int s_p q_h w]`` with ``s_u = d_phi`` and
``s_p = 1/sqrt(porosity)`` where the porosity is positive (:90-92), 1 elsewhere.

``evaluate_form(i, alpha_l, alpha_scaled_l)`` keeps the reference's three-argument signature; for whole-mesh assembly
set ``form.alpha_scaled`` (global vector) and list the form as an ordinary ``"form"``.
"""
import numpy as np
import torch

from .. import coefficients
from ..device import space_device
from .hdiv_general import HdivL2GeneralForm


class ScaledToPhysicalL2Projection(HdivL2GeneralForm):
    flux_name, scalar_name = "v", "q"
    alpha_scaled = None

    def point_coefficients(self, x, dim):
        n, nq, _ = x.shape
        phi = coefficients.evaluate(self.functions["porosity"], x)
        delta = coefficients.evaluate(self.functions["d_phi"], x)
        pos = phi > 0.0
        return {
            "c_uu": np.ones((n, nq)), "c_d": np.zeros((n, nq)), "c_pp": np.ones((n, nq)),
            "s_u": np.where(pos, delta, 1.0), "s_p": np.where(pos, 1.0 / np.sqrt(np.where(pos, phi, 1.0)), 1.0),
        }

    def launch_local(self, dev, data, alpha_dev, dest_dev, Ke, re, beta_dev=None):
        if beta_dev is None and self.alpha_scaled is not None:
            beta_dev = torch.as_tensor(np.ascontiguousarray(self.alpha_scaled, dtype=np.float64)).to(dev.device)
            if dest_dev is None:
                dest_dev = self._dest_for(dev, data)
        super().launch_local(dev, data, alpha_dev, dest_dev, Ke, re, beta_dev)

    def _dest_for(self, dev, data):
        full = dev.dest()
        if data.items is None:
            return full
        return full.index_select(0, torch.as_tensor(np.asarray(data.items)).to(full.device)).contiguous()

    def evaluate_form(self, element_index, alpha, alpha_scaled):
        self._check()
        dev = space_device(self.space)
        data = self.prepare(dev, np.array([int(element_index)], dtype=np.int64))
        nl = self.n_loc()
        up = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64).reshape(-1)).to(dev.device)
        dest = torch.arange(nl, dtype=torch.int32, device=dev.device).reshape(1, nl)
        Ke = torch.empty((1, nl, nl), dtype=torch.float64, device=dev.device)
        re = torch.empty((1, nl), dtype=torch.float64, device=dev.device)
        HdivL2GeneralForm.launch_local(self, dev, data, up(alpha), dest, Ke, re, up(alpha_scaled))
        return re[0].cpu().numpy(), Ke[0].cpu().numpy()
