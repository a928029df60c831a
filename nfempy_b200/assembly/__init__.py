from .csr_matrix import CSRMatrix
from .gpu_assembler import GPUAssembler, SequentialAssembler
