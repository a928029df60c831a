"""nfempy_b200: B200-native (sm_100a) finite-element assembly behind the nfempy API.

Public surface mirrors the reference's ``src`` packages for the assembly hot path:
``mesh.Mesh``, ``spaces.ProductSpace``, ``weak_forms.*``, ``assembly.GPUAssembler`` (= ``SequentialAssembler``).
"""
from . import _lib  # noqa: F401
from .assembly import CSRMatrix, GPUAssembler, SequentialAssembler  # noqa: F401
from .mesh import Mesh  # noqa: F401
from .spaces import ProductSpace  # noqa: F401

__version__ = "0.1.0"
