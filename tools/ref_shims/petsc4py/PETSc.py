"""See the package docstring."""
import numpy as np

COMM_WORLD = None
ASSEMBLED = []


class Vec:
    def __init__(self, n):
        self.array = np.zeros(n)

    @staticmethod
    def destroy(vec):
        return None


class Mat:
    def __init__(self):
        self.shape, self.type, self.entries = None, "aij", {}

    def createAIJ(self, size, *args, **kwargs):
        self.shape = tuple(int(v) for v in size)
        return self

    def setType(self, name):
        self.type = name

    def setValue(self, row, col, value, addv=True):
        key = (int(row), int(col))
        self.entries[key] = (self.entries.get(key, 0.0) + float(value)) if addv else float(value)

    def assemble(self):
        ASSEMBLED.append(self)

    def scale(self, factor):
        for key in self.entries:
            self.entries[key] *= float(factor)

    def getInfo(self):
        nz = float(len(self.entries))
        return {"nz_allocated": nz, "nz_used": nz, "nz_unneeded": 0.0}

    def getValuesCSR(self):
        keys = sorted(self.entries)
        rows = np.array([k[0] for k in keys], dtype=np.int64)
        indptr = np.zeros(self.shape[0] + 1, dtype=np.int64)
        np.add.at(indptr, rows + 1, 1)
        return (np.cumsum(indptr).astype(np.int32), np.array([k[1] for k in keys], dtype=np.int32),
                np.array([self.entries[k] for k in keys]))

    def createVecLeft(self):
        return Vec(self.shape[0])

    def createVecRight(self):
        return Vec(self.shape[1])

    @staticmethod
    def destroy(mat):
        return None


class _Permissive:
    """Solver-configuration calls (types, tolerances, field splits, factor packages ...) are accepted and ignored: the
    stand-in always solves directly."""

    def __getattr__(self, name):
        return lambda *args, **kwargs: None


class IS(_Permissive):
    pass


class PC(_Permissive):
    class CompositeType:
        ADDITIVE = "additive"

    def getFieldSplitSubKSP(self):
        return KSP(), KSP()


_PC = PC


class KSP(_Permissive):
    def create(self, comm=None):
        return self

    def setOperators(self, A, P=None):
        self.A = A

    def setType(self, name):
        return None

    def getPC(self):
        return _PC()

    def setTolerances(self, **kwargs):
        return None

    def setFromOptions(self):
        return None

    def solve(self, b, x):
        import scipy.sparse as sp
        import scipy.sparse.linalg as spl

        indptr, indices, data = self.A.getValuesCSR()
        K = sp.csr_matrix((data, indices, indptr), shape=self.A.shape)
        if self.A.type == "sbaij":  # only the upper triangle was inserted
            K = K + sp.triu(K, 1).T
        x.array[:] = spl.spsolve(K.tocsc(), b.array)

    def getConvergenceHistory(self):
        return np.zeros(0)

    @staticmethod
    def destroy(ksp):
        return None
