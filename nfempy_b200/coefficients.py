"""Host evaluation of the user's coefficient callables over all quadrature points at once.

The reference calls ``functions[name](x, y, z)`` cell by cell, either vectorised over the cell's quadrature
points (R/src/weak_forms/le_primal_weak_form.py:39-41) or point by point with scalars
(R/src/weak_forms/le_primal_weak_form.py:166-168).  Python callables cannot run on the device, so they are
evaluated once on the host over ``[n_items * n_q]`` points and uploaded.
"""
import numpy as np
import torch

from . import _host, _lib


def physical_points(dev, dim, points, cell_nodes_dev):
    """``x[n, n_q, 3]`` (numpy) of the reference points mapped into every cell, computed on the device."""
    n = int(cell_nodes_dev.shape[0])
    points = np.asarray(points, dtype=np.float64).reshape(-1, max(dim, 1))
    nq = len(points)
    shapes = dev.p1_shapes(dim, points) if dim > 0 else np.ones((1, 1))
    shape1 = torch.as_tensor(np.ascontiguousarray(shapes, dtype=np.float64)).to(dev.device)
    x = torch.empty((n, nq, 3), dtype=torch.float64, device=dev.device)
    _lib.call(
        "nfa_map_points",
        dim,
        nq,
        _lib.ptr(shape1),
        _lib.ptr(dev.coords),
        _lib.ptr(cell_nodes_dev),
        n,
        _lib.ptr(x),
        _lib.stream_ptr(),
    )
    return x.cpu().numpy()


_EVAL_CHUNK = 1 << 19  # points per threaded call of a coefficient callable


class _NoBroadcast(Exception):
    pass


def _evaluate_vectorised(fn, flat, n_comp, extra):
    """One vectorised call over ``flat[m, 3]``: ``[m]`` (scalar) or ``[m, n_comp]``; ``_NoBroadcast`` when the callable
    does not take arrays."""
    m = len(flat)
    try:
        out = np.asarray(fn(flat[:, 0], flat[:, 1], flat[:, 2], *extra), dtype=np.float64)
        if n_comp is None:
            if out.ndim == 0:
                out = np.full(m, float(out))
            out = out.reshape(-1)
            if out.shape != (m,):
                raise ValueError
            return out
        if out.ndim == 1 and out.shape[0] != m:  # constant vector
            out = np.repeat(out[:, None], m, axis=1)
        out = out.reshape(out.shape[0], -1)
        if out.shape[1] != m or out.shape[0] < n_comp:
            raise ValueError
        return np.ascontiguousarray(out[:n_comp].T)
    except (ValueError, TypeError, IndexError):
        raise _NoBroadcast()


def evaluate(fn, x, n_comp=None, extra=()):
    """Evaluates ``fn(x, y, z, *extra)`` over ``x[n, n_q, 3]`` (``extra`` = e.g. the time of a time-dependent form).

    Returns ``[n, n_q]`` for scalar coefficients (``n_comp is None``) or ``[n, n_q, n_comp]`` for vector data.
    Tries vectorised calls -- over chunks of points on a few host threads when there are millions of them (numpy
    releases the GIL inside its loops; a time loop re-evaluates its loads at every new time) -- and falls back to the
    reference's point-by-point calling convention when the callable does not broadcast."""
    n, nq, _ = x.shape
    flat = x.reshape(-1, 3)
    m = n * nq
    shape = (n, nq) if n_comp is None else (n, nq, n_comp)
    try:
        pool = _host.pool() if m >= 2 * _EVAL_CHUNK else None
        if pool is None:
            return _evaluate_vectorised(fn, flat, n_comp, extra).reshape(shape)
        parts = _host.parts(m, _EVAL_CHUNK)
        head = _evaluate_vectorised(fn, flat[parts[0][0]:parts[0][1]], n_comp, extra)  # settles the calling convention
        rest = list(pool.map(lambda ab: _evaluate_vectorised(fn, flat[ab[0]:ab[1]], n_comp, extra), parts[1:]))
        return np.concatenate([head] + rest, axis=0).reshape(shape)
    except _NoBroadcast:
        pass
    if n_comp is None:
        vals = np.array([float(fn(p[0], p[1], p[2], *extra)) for p in flat])
        return vals.reshape(n, nq)
    vals = np.array([np.asarray(fn(p[0], p[1], p[2], *extra), dtype=np.float64).reshape(-1)[:n_comp] for p in flat])
    return vals.reshape(n, nq, n_comp)


def classify(values):
    """Smallest faithful representation of a scalar coefficient ``[n, n_q]``:
    returns ``(array, stride_cell, stride_q, kind)`` with kind in {"constant", "cell", "point"}."""
    n, nq = values.shape
    if n == 0:
        return np.zeros(1), 0, 0, "constant"
    first = values[:, :1]
    if np.all(values == first):
        if np.all(first == first[0, 0]):
            return np.array([first[0, 0]]), 0, 0, "constant"
        return np.ascontiguousarray(first[:, 0]), 1, 0, "cell"
    return np.ascontiguousarray(values), nq, 1, "point"
