"""ORACLE (test infrastructure only -- never imported by the product path).

Restates the reference scatter and the PETSc object it fills:

* ``scatter_form`` / ``scatter_bc_form`` -- R/src/assembly/ScatterFormData.py:5-25, 36-56: residual added dof by dof,
  Jacobian added entry by entry (all ``n_loc^2`` pairs, exact zeros included, row-major order);
* ``DokMatrix``    -- ``PETSc.Mat().createAIJ([n, n])`` + ``setValue(..., addv=True)`` + ``assemble()`` +
  ``getValuesCSR()`` as described in SURVEY.md A.6: additions happen in call order, columns come out sorted per
  row, explicitly inserted zeros stay in the pattern, indices are 32-bit.  petsc4py is absent, so this is a
  dictionary of keys.
"""
import numpy as np


class DokMatrix:
    def __init__(self, n):
        self.n = n
        self.data = {}

    def setValue(self, row, col, value, addv=True):
        key = (int(row), int(col))
        if addv:
            self.data[key] = self.data.get(key, 0.0) + float(value)
        else:
            self.data[key] = float(value)

    def scale(self, a):
        for k in self.data:
            self.data[k] *= a

    def getValuesCSR(self):
        keys = sorted(self.data.keys())
        indptr = np.zeros(self.n + 1, dtype=np.int32)
        for r, _ in keys:
            indptr[r + 1] += 1
        indptr = np.cumsum(indptr, dtype=np.int32)
        indices = np.array([c for _, c in keys], dtype=np.int32)
        data = np.array([self.data[k] for k in keys], dtype=np.float64)
        return indptr, indices, data


def scatter_lhs(jac_g, jac_l, dest):
    data = jac_l.ravel()
    row = np.repeat(dest, len(dest))
    col = np.tile(dest, len(dest))
    for k in range(data.shape[0]):
        jac_g.setValue(row=row[k], col=col[k], value=data[k], addv=True)


def scatter_rhs(res_g, res_l, dest):
    for k, s in enumerate(dest):
        res_g[s] += res_l[k]


def scatter_form(space, form, functions, idx, res_g, jac_g, alpha):
    dest = space.destination_indexes(idx)
    r_l, j_l = form(space, functions, idx, alpha[dest])
    scatter_rhs(res_g, r_l, dest)
    scatter_lhs(jac_g, j_l, dest)


def scatter_bc_form(space, form, functions, idx, res_g, jac_g, alpha):
    dest = space.bc_destination_indexes(idx)
    r_l, j_l = form(space, functions, idx, alpha[dest])
    scatter_rhs(res_g, r_l, dest)
    scatter_lhs(jac_g, j_l, dest)


def scatter_time_dependent_form(space, form, functions, idx, res_g, jac_g, alpha_n, alpha, t):
    """scatter_time_dependent_form_data, R/src/assembly/ScatterFormData.py:28-33 (note the argument order it hands on)."""
    dest = space.destination_indexes(idx)
    alpha_l, alpha_l_n = alpha[dest], alpha_n[dest]
    r_l, j_l = form(space, functions, idx, alpha_l_n, alpha_l, t)
    scatter_rhs(res_g, r_l, dest)
    scatter_lhs(jac_g, j_l, dest)


def scatter_time_dependent_bc_form(space, form, functions, idx, res_g, jac_g, alpha_n, alpha, t):
    """scatter_time_dependent_bc_form_data, R/src/assembly/ScatterFormData.py:43-49."""
    dest = space.bc_destination_indexes(idx)
    r_l, j_l = form(space, functions, idx, alpha_n[dest], t)
    scatter_rhs(res_g, r_l, dest)
    scatter_lhs(jac_g, j_l, dest)


def assemble_time_dependent(space, forms, alphas, t):
    """``forms``: ``(kind, form_fn, functions, n_items)`` with kind in {"time_dependent_form", "time_dependent_bc_form"};
    ``alphas = (alpha_n, alpha)`` exactly as listed in ``form_to_input_list``."""
    n = space.n_dof
    alpha_n, alpha = alphas
    res_g, jac_g = np.zeros(n), DokMatrix(n)
    for kind, form, functions, n_items in forms:
        scatter = scatter_time_dependent_form if kind == "time_dependent_form" else scatter_time_dependent_bc_form
        for i in (range(n_items) if isinstance(n_items, int) else n_items):
            scatter(space, form, functions, i, res_g, jac_g, alpha_n, alpha, t)
    indptr, indices, data = jac_g.getValuesCSR()
    return indptr, indices, data, res_g


def assemble(space, forms, alpha=None):
    """``forms``: list of ``(kind, form_fn, functions, n_items)`` with kind in {"form", "bc_form"}.
    Returns ``(indptr, indices, data, rhs)`` exactly as the reference + PETSc would hand them to a solver."""
    n = space.n_dof
    alpha = np.zeros(n) if alpha is None else alpha
    res_g = np.zeros(n)
    jac_g = DokMatrix(n)
    for kind, form, functions, n_items in forms:
        scatter = scatter_form if kind == "form" else scatter_bc_form
        for i in (range(n_items) if isinstance(n_items, int) else n_items):  # a count or an explicit sequence
            scatter(space, form, functions, i, res_g, jac_g, alpha)
    indptr, indices, data = jac_g.getValuesCSR()
    return indptr, indices, data, res_g


def assemble_filtered(space, forms, alpha=None, upper=False, explicit_diagonal=False):
    """The inline scatter closures of the example drivers (R/examples/cauchy_elasticity/cg_elasticity_example_1.py:
    92-145, R/examples/cosserat_elasticity/mfe_cosserat_elasticity_example_3.py:150-243): only entries with
    ``~np.isclose(value, 1e-16)`` are inserted, optionally only ``row <= col`` ("sbaij"), and every local dof of a
    volume form gets an explicit ``(i, i, 0.0)``."""
    n = space.n_dof
    alpha = np.zeros(n) if alpha is None else alpha
    res_g, jac_g = np.zeros(n), DokMatrix(n)
    for kind, form, functions, n_items in forms:
        for i in (range(n_items) if isinstance(n_items, int) else n_items):
            dest = space.destination_indexes(i) if kind == "form" else space.bc_destination_indexes(i)
            r_l, j_l = form(space, functions, i, alpha[dest])
            scatter_rhs(res_g, r_l, dest)
            data = j_l.ravel()
            row, col = np.repeat(dest, len(dest)), np.tile(dest, len(dest))
            for idx in np.where(np.logical_not(np.isclose(data, 1.0e-16)))[0]:
                if upper and row[idx] > col[idx]:
                    continue
                jac_g.setValue(row=row[idx], col=col[idx], value=data[idx], addv=True)
            if explicit_diagonal and kind == "form":
                for idx in dest:
                    jac_g.setValue(row=idx, col=idx, value=0.0, addv=True)
    indptr, indices, data = jac_g.getValuesCSR()
    return indptr, indices, data, res_g
