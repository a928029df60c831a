"""Stand-in for PTNobel/AutoDiff (absent here): forward-mode differentiation of numpy expressions, restricted to what the
reference's weak forms use (R/src/weak_forms/*.py: ``with ad.AutoDiff(alpha) as alpha`` ... ``el_form.val``, ``el_form.der``).

``AutoDiff(alpha)`` for a 1-D ``alpha`` of length n hands out ``val = alpha[None, :]`` and ``der[0, i, j] = delta_ij``
(``der.shape = val.shape + (n,)``): the forms index ``alpha[:, c:n:nc]``, multiply ``alpha @ grad_phi`` with
``grad_phi[n_phi, dim]`` and end with ``el_form.der.reshape((n, n))``, which fixes these shapes."""
import numpy as np

from .vecvalder import VecValDer


class AutoDiff:
    def __init__(self, *xs):
        self.xs = xs

    def __enter__(self):
        out, offset = [], 0
        total = sum(np.asarray(x).size for x in self.xs)
        for x in self.xs:
            x = np.asarray(x, dtype=float)
            val = x.reshape(1, -1) if x.ndim == 1 else x
            der = np.zeros(val.shape + (total,))
            flat = der.reshape(val.size, total)
            flat[np.arange(val.size), offset + np.arange(val.size)] = 1.0
            offset += val.size
            out.append(VecValDer(val.copy(), der))
        self._true_zeros = np.zeros

        def zeros(shape, *args, **kwargs):
            val = self._true_zeros(shape, *args, **kwargs)
            return VecValDer(val, self._true_zeros(val.shape + (total,)))

        np.zeros = zeros
        return out[0] if len(out) == 1 else tuple(out)

    def __exit__(self, *exc):
        np.zeros = self._true_zeros
        return False
