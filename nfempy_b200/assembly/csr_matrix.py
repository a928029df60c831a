"""``CSRMatrix``: the object the GPU assembler fills in place of the reference's ``PETSc.Mat``.

The reference creates ``PETSc.Mat().createAIJ([n, n])`` and lets ``setValue(..., addv=True)`` grow the pattern
(R/examples/cauchy_elasticity/cg_elasticity_example_1.py:59-61, R/src/assembly/ScatterFormData.py:5-11), then
hands ``A`` to a KSP or reads ``ai, aj, av = A.getValuesCSR()``
(R/examples/cosserat_elasticity/md_primal_cosserat.py:415-417).  petsc4py is not available here, so this shim
answers the part of the ``Mat`` protocol the assembly path uses and converts to scipy / PETSc on request.
Values live on the device; index arrays are int32 (PETSc's index type), columns sorted per row, explicit zeros
kept.
"""
import numpy as np
import torch


class CSRMatrix:
    def __init__(self, n_dof=None):
        self.n = n_dof
        self.plan_key = None
        self.indptr_dev = None
        self.indices_dev = None
        self.values_dev = None
        self._indptr = None
        self._indices = None
        self.is_zero = True
        # True: getValuesCSR() returns a view of ONE pinned host buffer that the next call overwrites (a solver loop that
        # consumes the arrays before reassembling); False (PETSc's behaviour): a fresh array per call
        self.reuse_host_buffer = False
        self._values_host = None

    # -- filled by the assembler -------------------------------------------------------------------
    def bind_pattern(self, plan_key, n_dof, indptr_dev, indices_dev):
        self.plan_key = plan_key
        self.n = n_dof
        self.indptr_dev, self.indices_dev = indptr_dev, indices_dev
        self._indptr = self._indices = None
        self._values_host = None
        self.values_dev = torch.zeros(indices_dev.shape[0], dtype=torch.float64, device=indices_dev.device)
        self.is_zero = True

    @property
    def nnz(self):
        return 0 if self.indices_dev is None else int(self.indices_dev.shape[0])

    # -- PETSc Mat protocol (subset) ---------------------------------------------------------------
    def createAIJ(self, size, *args, **kwargs):
        self.n = int(size[0])
        return self

    def setType(self, *_):
        return self

    def setUp(self):
        return self

    def assemble(self):
        return self

    def setValue(self, row, col, value, addv=True):
        raise NotImplementedError(
            "CSRMatrix is filled by GPUAssembler.scatter_forms(); entry-by-entry insertion is the reference's "
            "host path and is not provided"
        )

    def scale(self, a):
        if self.values_dev is not None:
            if float(a) == 0.0:  # the reassembly idiom jac_g.scale(0.0): a write-only fill instead of read-multiply-write
                self.values_dev.zero_()
                self.is_zero = True
            else:
                self.values_dev.mul_(float(a))
        return self

    def zeroEntries(self):
        return self.scale(0.0)

    def getSize(self):
        return (self.n, self.n)

    def getInfo(self):
        return {"nz_allocated": float(self.nnz), "nz_used": float(self.nnz), "nz_unneeded": 0.0}

    def getValuesCSR(self):
        if self.indices_dev is None:
            n = self.n or 0
            return np.zeros(n + 1, dtype=np.int32), np.zeros(0, dtype=np.int32), np.zeros(0)
        if self._indptr is None:
            self._indptr = self.indptr_dev.cpu().numpy()
            self._indices = self.indices_dev.cpu().numpy()
        if not self.reuse_host_buffer:
            return self._indptr, self._indices, self.values_dev.cpu().numpy()
        if self._values_host is None or self._values_host.shape != self.values_dev.shape:
            self._values_host = torch.empty(self.values_dev.shape, dtype=torch.float64).pin_memory()
        self._values_host.copy_(self.values_dev, non_blocking=False)
        return self._indptr, self._indices, self._values_host.numpy()

    def createVecLeft(self):
        return _Vec(self.n)

    def createVecRight(self):
        return _Vec(self.n)

    def destroy(self):
        self.values_dev = self.indptr_dev = self.indices_dev = None

    # -- hand-off ------------------------------------------------------------------------------------
    def to_scipy(self):
        import scipy.sparse as sp

        indptr, indices, data = self.getValuesCSR()
        return sp.csr_matrix((data, indices, indptr), shape=(self.n, self.n))

    def to_petsc(self):
        from petsc4py import PETSc  # noqa: raises ImportError where PETSc is absent

        indptr, indices, data = self.getValuesCSR()
        return PETSc.Mat().createAIJWithArrays([self.n, self.n], (indptr, indices, data))


class _Vec:
    def __init__(self, n):
        self.array = np.zeros(n)
