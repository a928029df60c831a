"""Stand-in for PTNobel/AutoDiff (absent here): forward-mode differentiation of numpy expressions, restricted to what the
reference's weak forms use (R/src/weak_forms/*.py: ``with ad.AutoDiff(alpha) as alpha`` ... ``el_form.val``, ``el_form.der``).

``AutoDiff(alpha)`` for a 1-D ``alpha`` of length n hands out ``val = alpha[None, :]`` and ``der[0, i, j] = delta_ij``
(``der.shape = val.shape + (n,)``): the forms index ``alpha[:, c:n:nc]``, multiply ``alpha @ grad_phi`` with
``grad_phi[n_phi, dim]`` and end with ``el_form.der.reshape((n, n))``, which fixes these shapes.

While the context is open ``np.zeros`` and ``np.array`` are replaced (the real package must do the same, since the forms rely on
it): they write ``el_form = np.zeros(n_dof)`` / ``integrand = np.zeros((1, n_dof))`` INSIDE the ``with`` block, assign
differentiated slices into them and read ``el_form.val`` / ``el_form.der`` afterwards -- which only works when those arrays are
born as ``VecValDer`` with a zero derivative; and they build matrices as nested lists of differentiated scalars
(``Gamma_outer``, lce_dual_weak_form.py:366-372), which must come out as one ``VecValDer``."""
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

        self._true_array = np.array

        def array(obj, *args, **kwargs):
            """``np.array([[0.0 * t, -t2, +t1], ...])`` of differentiated scalars / arrays gives a ``VecValDer``."""
            if not _contains(obj):
                return self._true_array(obj, *args, **kwargs)
            return _assemble(obj, total, self._true_zeros)

        np.zeros, np.array = zeros, array
        return out[0] if len(out) == 1 else tuple(out)

    def __exit__(self, *exc):
        np.zeros, np.array = self._true_zeros, self._true_array
        return False


def _contains(obj):
    if isinstance(obj, VecValDer):
        return True
    if isinstance(obj, (list, tuple)):
        return any(_contains(o) for o in obj)
    return False


def _assemble(obj, total, zeros):
    if isinstance(obj, VecValDer):
        return obj
    if isinstance(obj, (list, tuple)):
        parts = [_assemble(o, total, zeros) for o in obj]
        return VecValDer(np.stack([p.val for p in parts]), np.stack([p.der for p in parts]))
    val = np.asarray(obj, dtype=float)
    return VecValDer(val, zeros(val.shape + (total,)))
