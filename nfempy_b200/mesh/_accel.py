"""Sort / unique / search primitives of the mesh and space numbering, on the device when there is one.

The numbering of R/src/mesh/mesh.py:239-339 and R/src/spaces/dof_map.py:38-129 is restated with whole-array operations
(``mesh.py``, ``spaces/discrete_space.py``); at 10^6..10^7 tetrahedra they are dominated by a dozen ``unique`` /
``argsort`` / ``searchsorted`` calls over 10^7..10^8 integer keys -- seconds to a minute of numpy per call chain, and
replicated on every rank of a sharded run.  These wrappers run the same primitives with radix sorts on the GPU
(torch tensors as buffers) when CUDA is available and the array is large, and with numpy otherwise: identical results
(stable orders, first occurrences), so the CPU tests pin both.  SURVEY.md section 8 (f4): numbering on the device.
"""
import numpy as np

ACCEL_MIN = 1 << 18  # below this the transfers cost more than the sort
_state = {"enabled": None}


def _device():
    if _state["enabled"] is False:
        return None
    try:
        import torch

        if torch.cuda.is_available():
            _state["enabled"] = True
            return torch.device("cuda", torch.cuda.current_device())
    except Exception:
        pass
    _state["enabled"] = False
    return None


def enable(flag):
    """Force the device path on (None = auto) or off (False)."""
    _state["enabled"] = None if flag else False


def _use(a):
    return a.size >= ACCEL_MIN and a.dtype.kind in "iu" and _device() is not None


def unique(keys, return_index=False, return_inverse=False):
    """``np.unique`` of a 1-D integer array (ascending values; ``return_index`` = first occurrence)."""
    keys = np.asarray(keys)
    if keys.ndim != 1 or not _use(keys):
        return np.unique(keys, return_index=return_index, return_inverse=return_inverse)
    import torch

    k = torch.as_tensor(np.ascontiguousarray(keys, dtype=np.int64)).to(_device())
    sk, perm = torch.sort(k, stable=True)
    head = torch.ones_like(sk, dtype=torch.bool)
    head[1:] = sk[1:] != sk[:-1]
    out = [sk[head].cpu().numpy().astype(keys.dtype, copy=False)]
    if return_index:
        out.append(perm[head].cpu().numpy())
    if return_inverse:
        inv = torch.empty_like(perm)
        inv[perm] = torch.cumsum(head.to(torch.int64), 0) - 1
        out.append(inv.cpu().numpy())
    return out[0] if len(out) == 1 else tuple(out)


def unique_rows2(a, b):
    """First-occurrence index and inverse of the distinct pairs ``(a[i], b[i])`` in lexicographic order -- what
    ``np.unique(np.stack([a, b], 1), axis=0, return_index=True, return_inverse=True)`` returns after the values."""
    a, b = np.asarray(a), np.asarray(b)
    if not _use(a):
        _, first, inv = np.unique(np.stack([a, b], axis=1), axis=0, return_index=True, return_inverse=True)
        return first, inv.ravel()
    import torch

    dev = _device()
    ta = torch.as_tensor(np.ascontiguousarray(a, dtype=np.int64)).to(dev)
    tb = torch.as_tensor(np.ascontiguousarray(b, dtype=np.int64)).to(dev)
    _, p1 = torch.sort(tb, stable=True)
    _, p2 = torch.sort(ta[p1], stable=True)
    perm = p1[p2]  # lexicographic by (a, b), original order among equals
    sa, sb = ta[perm], tb[perm]
    head = torch.ones_like(sa, dtype=torch.bool)
    head[1:] = (sa[1:] != sa[:-1]) | (sb[1:] != sb[:-1])
    inv = torch.empty_like(perm)
    inv[perm] = torch.cumsum(head.to(torch.int64), 0) - 1
    return perm[head].cpu().numpy(), inv.cpu().numpy()


def argsort(a):
    """Stable argsort of a 1-D integer array."""
    a = np.asarray(a)
    if a.ndim != 1 or not _use(a):
        return np.argsort(a, kind="stable")
    import torch

    return torch.sort(torch.as_tensor(np.ascontiguousarray(a, dtype=np.int64)).to(_device()), stable=True)[1].cpu().numpy()


def searchsorted(sorted_keys, values):
    sorted_keys, values = np.asarray(sorted_keys), np.asarray(values)
    if not _use(values) or sorted_keys.dtype.kind not in "iu":
        return np.searchsorted(sorted_keys, values)
    import torch

    dev = _device()
    s = torch.as_tensor(np.ascontiguousarray(sorted_keys, dtype=np.int64)).to(dev)
    v = torch.as_tensor(np.ascontiguousarray(values, dtype=np.int64)).to(dev)
    return torch.searchsorted(s, v).cpu().numpy().reshape(values.shape)


def sort_last(a):
    """``np.sort(a, axis=-1)`` (rows of 2..4 node ids)."""
    a = np.asarray(a)
    if not _use(a):
        return np.sort(a, axis=-1)
    import torch

    return torch.sort(torch.as_tensor(np.ascontiguousarray(a, dtype=np.int64)).to(_device()), dim=-1)[0].cpu().numpy()
