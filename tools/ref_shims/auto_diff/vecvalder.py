"""``VecValDer(val, der)``: value array + derivative array of shape ``val.shape + (n,)``.  ``__array_ufunc__`` routes numpy's
own operators to the methods below, so ``ndarray @ v``, ``np.float64 * v`` and ``ndarray += v`` (whose result rebinds the
name to a ``VecValDer``, as the forms rely on: ``el_form = np.zeros(n); el_form += ...; el_form.val``) all land here."""
import numpy as np


def _is(v):
    return isinstance(v, VecValDer)


class VecValDer:
    __slots__ = ("val", "der")

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        if method != "__call__" or kwargs:
            return NotImplemented
        name = ufunc.__name__
        if name == "negative":
            return -inputs[0]
        if name not in _BINARY:
            return NotImplemented
        a, b = inputs
        forward, reflected = _BINARY[name]
        return getattr(a, forward)(b) if _is(a) else getattr(b, reflected)(a)

    def __init__(self, val, der):
        self.val = np.asarray(val, dtype=float)
        self.der = np.asarray(der, dtype=float)
        assert self.der.shape[: self.val.ndim] == self.val.shape, (self.val.shape, self.der.shape)

    @property
    def n(self):
        return self.der.shape[-1]

    @property
    def shape(self):
        return self.val.shape

    # ---- structure ---------------------------------------------------------------------------------------
    def __getitem__(self, idx):
        idx = idx if isinstance(idx, tuple) else (idx,)
        return VecValDer(self.val[idx], self.der[idx + (slice(None),)])

    def __setitem__(self, idx, value):
        idx = idx if isinstance(idx, tuple) else (idx,)
        if _is(value):
            self.val[idx] = value.val
            self.der[idx + (slice(None),)] = value.der
        else:
            self.val[idx] = value
            self.der[idx + (slice(None),)] = 0.0

    def __len__(self):
        return len(self.val)

    @property
    def T(self):
        axes = tuple(range(self.val.ndim))[::-1]
        return VecValDer(self.val.T, np.transpose(self.der, axes + (self.val.ndim,)))

    def reshape(self, *shape):
        shape = shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list)) else shape
        val = self.val.reshape(shape)
        return VecValDer(val, self.der.reshape(val.shape + (self.n,)))

    def ravel(self):
        return self.reshape((-1,))

    def trace(self):
        return VecValDer(self.val.trace(), self.der.trace())

    # ---- linear arithmetic -------------------------------------------------------------------------------
    def __pos__(self):
        return self

    def __neg__(self):
        return VecValDer(-self.val, -self.der)

    def __add__(self, other):
        if _is(other):
            val = self.val + other.val
            return VecValDer(val, np.broadcast_to(self.der, val.shape + (self.n,)) + other.der)
        other = np.asarray(other, dtype=float)
        val = self.val + other
        return VecValDer(val, np.broadcast_to(self.der, val.shape + (self.n,)).copy())

    __radd__ = __add__

    def __sub__(self, other):
        return self + (-other)

    def __rsub__(self, other):
        return (-self) + other

    def __mul__(self, other):
        if _is(other):
            return VecValDer(self.val * other.val, self.der * other.val[..., None] + other.der * self.val[..., None])
        other = np.asarray(other, dtype=float)
        return VecValDer(self.val * other, self.der * other[..., None])

    __rmul__ = __mul__

    def __truediv__(self, other):
        if _is(other):
            val = self.val / other.val
            return VecValDer(val, (self.der - val[..., None] * other.der) / other.val[..., None])
        other = np.asarray(other, dtype=float)
        return VecValDer(self.val / other, self.der / other[..., None])

    def __rtruediv__(self, other):
        other = np.asarray(other, dtype=float)
        val = other / self.val
        return VecValDer(val, -(val / self.val)[..., None] * self.der)

    def __matmul__(self, other):
        if _is(other):
            return VecValDer(self.val @ other.val, _mm_left(self, other.val) + _mm_right(self.val, other))
        return VecValDer(self.val @ np.asarray(other, dtype=float), _mm_left(self, np.asarray(other, dtype=float)))

    def __rmatmul__(self, other):
        other = np.asarray(other, dtype=float)
        return VecValDer(other @ self.val, _mm_right(other, self))


_BINARY = {"add": ("__add__", "__radd__"), "subtract": ("__sub__", "__rsub__"), "multiply": ("__mul__", "__rmul__"),
           "true_divide": ("__truediv__", "__rtruediv__"), "divide": ("__truediv__", "__rtruediv__"),
           "matmul": ("__matmul__", "__rmatmul__")}


def _mm_left(v, b):
    """d(v @ b) for constant b: the derivative axis rides along as a batch axis."""
    moved = np.moveaxis(v.der, -1, 0)  # [n, *val.shape]
    if v.val.ndim == 1:
        out = np.einsum("nk,k...->n...", moved, b)
    else:
        out = moved @ b
    return np.moveaxis(out, 0, -1)


def _mm_right(a, v):
    """d(a @ v) for constant a."""
    moved = np.moveaxis(v.der, -1, 0)  # [n, *val.shape]
    if v.val.ndim == 1:
        out = np.einsum("...k,nk->n...", a, moved)
    else:
        out = a @ moved
    return np.moveaxis(out, 0, -1)
