"""Device-resident copies of the mesh/space arrays the kernels read (torch tensors used purely as buffers)."""
import numpy as np
import torch

from . import _lib
from .basis.reference_element import create_element


def require_cuda(device=None):
    if not torch.cuda.is_available():
        raise _lib.NfaError("nfempy_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    _lib.load()
    return torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())


def to_dev(array, dtype, device):
    return torch.as_tensor(np.ascontiguousarray(array, dtype=dtype)).to(device, non_blocking=False)


class SpaceDevice:
    """Cache of device buffers for one ``ProductSpace``: coordinates, connectivity, destination tables."""

    def __init__(self, space, device=None):
        self.space = space
        self.device = require_cuda(device)
        self._cache = {}
        first = next(iter(space.discrete_spaces.values()))
        self.mesh = first.mesh
        self.coords = to_dev(self.mesh.points, np.float64, self.device)

    def _get(self, key, build):
        if key not in self._cache:
            self._cache[key] = build()
        return self._cache[key]

    def cell_nodes(self, field):
        s = self.space.discrete_spaces[field]
        return self._get(("cn", field), lambda: to_dev(s.cell_nodes(), np.int32, self.device))

    def bc_cell_nodes(self, field):
        s = self.space.discrete_spaces[field]
        return self._get(("bcn", field), lambda: to_dev(s.bc_cell_nodes(), np.int32, self.device))

    def dest(self):
        return self._get(
            ("dest", self.space.dof_shift), lambda: to_dev(self.space.destination_table(), np.int32, self.device)
        )

    def bc_dest(self, field=None):
        return self._get(
            ("bcdest", field, self.space.dof_shift),
            lambda: to_dev(self.space.bc_destination_table(field), np.int32, self.device),
        )

    def tensor(self, key, build):
        """Cache an arbitrary host array on the device (basis tables, weights...)."""
        return self._get(key, lambda: to_dev(build(), np.float64, self.device))

    @staticmethod
    def p1_shapes(dim, points):
        ref = create_element("Lagrange", dim, 1)
        return ref.tabulate(0, points)[0, :, :, 0]


def space_device(space, device=None):
    dev = getattr(space, "_nfa_device", None)
    if dev is None or (device is not None and torch.device(device) != dev.device):
        dev = SpaceDevice(space, device)
        space._nfa_device = dev
    return dev
