"""Stand-in for petsc4py (absent here), just enough for the reference's example drivers to run from assembly to solve
(R/examples/cauchy_elasticity/cg_elasticity_example_1.py:58-180): ``PETSc.Mat`` is a dict of keys with ADD_VALUES,
``getValuesCSR`` returns the stored entries row by row with ascending columns (what PETSc's AIJ / SBAIJ hand back),
``KSP.solve`` is a direct sparse solve (an "sbaij" matrix is mirrored first).  Every matrix that gets ``assemble()``d is kept
in ``PETSc.ASSEMBLED`` so that tools/make_reference_pins.py can read what the driver's own scatter closure inserted."""
from . import PETSc  # noqa: F401
