"""``l2_projector`` -- GPU version of R/src/postprocess/projectors.py:8-91: assembles ``L2ProjectorWeakForm`` over all
elements with the driver's scatter policy (entries with ``~isclose(v, 1e-16)``, upper triangle for the symmetric
storage), solves ``A x = -r`` and returns the coefficient vector.  The reference factorises with PETSc/MUMPS
(``preonly`` + Cholesky / LU); petsc4py is absent here, so the assembled CSR goes to ``scipy.sparse.linalg.spsolve``."""
import numpy as np

from ..assembly import CSRMatrix, GPUAssembler
from ..weak_forms import L2ProjectorWeakForm


def l2_projector(fe_space, functions, dof_shift=0, symmetric_solver_q=True):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl

    # the reference resets the space's own shift and adds `dof_shift` to every destination index (R/...:13-15, 33-34); the
    # projection is local to the space, so the shift moves rows and columns together and the returned LOCAL coefficient
    # vector does not depend on it -- it is accepted for call compatibility (mixed-dimensional drivers pass it,
    # R/examples/md_elliptic_scalar/md_elliptic_scalar.py:384-390)
    if int(dof_shift) != dof_shift:
        raise TypeError("l2_projector: dof_shift must be an integer")
    n_dof = fe_space.n_dof
    saved_shift = fe_space.dof_shift
    fe_space.dof_shift = 0  # local projection (R/...:13-15)
    try:
        weak_form = L2ProjectorWeakForm(fe_space)
        weak_form.functions = functions
        name = list(fe_space.discrete_spaces.keys())[0]
        n_els = len(fe_space.discrete_spaces[name].elements)
        rg = np.zeros(n_dof)
        asm = GPUAssembler(fe_space, CSRMatrix(n_dof), rg,
                           pattern_policy="filtered_upper" if symmetric_solver_q else "filtered")
        asm.form_to_input_list = {"projector": ["form", range(n_els), weak_form, np.zeros(n_dof)]}
        asm.scatter_forms()
        indptr, indices, data = asm.jacobian.getValuesCSR()
    finally:
        fe_space.dof_shift = saved_shift
    A = sp.csr_matrix((data, indices, indptr), shape=(n_dof, n_dof))
    if symmetric_solver_q:
        A = A + sp.triu(A, k=1).T
    return spl.spsolve(A.tocsc(), -rg)
