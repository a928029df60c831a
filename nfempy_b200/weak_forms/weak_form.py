"""``WeakForm`` base classes.

``WeakForm`` keeps the reference's surface (R/src/weak_forms/weak_from.py:6-29): ``space``, ``functions`` and
``evaluate_form(element_index, alpha) -> (r_el, j_el)``.  ``BatchedWeakForm`` adds what the GPU assembler needs:
whole-sequence evaluation in one kernel launch.  ``evaluate_form`` of a batched form runs the same CUDA kernel
on a batch of one cell -- there is no host implementation.
"""
from abc import ABC, abstractmethod

import numpy as np
import torch

from .. import _lib
from ..device import space_device


class WeakForm(ABC):
    def __init__(self, space):
        self.space = space
        self._functions = None

    @property
    def space(self):
        return self._space

    @space.setter
    def space(self, space):
        self._space = space

    @property
    def functions(self):
        return self._functions

    @functions.setter
    def functions(self, functions):
        self._functions = functions

    @abstractmethod
    def evaluate_form(self, element_index, alpha):
        pass


class FormData:
    """Device-resident inputs of one form over one item sequence (output of ``prepare``)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def get(self, name, default=None):
        return self.__dict__.get(name, default)


class BatchedWeakForm(WeakForm):
    form_kind = "form"  # "form": items index space.elements; "bc_form": items index bc_elements

    # -- to be provided by concrete forms ---------------------------------------------------------
    def n_items_total(self):
        raise NotImplementedError

    def dest_table(self):
        """numpy ``[n_items_total, n_loc]`` destination indices of every item of this form."""
        raise NotImplementedError

    def prepare(self, dev, items=None):
        raise NotImplementedError

    def launch_local(self, dev, data, alpha_dev, dest_dev, Ke, re):
        raise NotImplementedError

    def fused_spec(self, data):
        """Description of the fused (no local matrices) path for this data, or None."""
        return None

    # -- shared ------------------------------------------------------------------------------------
    def _check(self):
        if self.space is None or self.functions is None:
            raise ValueError("weak form needs .space and .functions")

    def n_loc(self):
        return int(self.dest_table().shape[1])

    def local_blocks(self, dev, data, alpha_dev=None, dest_dev=None, want_matrix=True):
        n, nl = data.n_items, self.n_loc()
        Ke = torch.empty((n, nl, nl), dtype=torch.float64, device=dev.device) if want_matrix else None
        re = torch.empty((n, nl), dtype=torch.float64, device=dev.device)
        self.launch_local(dev, data, alpha_dev, dest_dev, Ke, re)
        return Ke, re

    def evaluate_form(self, element_index, alpha):
        self._check()
        dev = space_device(self.space)
        items = np.array([int(element_index)], dtype=np.int64)
        data = self.prepare(dev, items)
        nl = self.n_loc()
        alpha_dev = dest_dev = None
        if alpha is not None and np.any(np.asarray(alpha) != 0.0):
            alpha_dev = torch.as_tensor(np.ascontiguousarray(alpha, dtype=np.float64).reshape(-1)).to(dev.device)
            dest_dev = torch.arange(nl, dtype=torch.int32, device=dev.device).reshape(1, nl)
        Ke, re = self.local_blocks(dev, data, alpha_dev, dest_dev)
        return re[0].cpu().numpy(), Ke[0].cpu().numpy()


def take_rows(tensor, items):
    """Rows ``items`` of a device table (or the table itself for ``items is None``)."""
    if items is None:
        return tensor
    idx = torch.as_tensor(np.asarray(items, dtype=np.int64)).to(tensor.device)
    return tensor.index_select(0, idx).contiguous()
