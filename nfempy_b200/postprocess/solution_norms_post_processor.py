"""Norms of the exact solution -- GPU versions of ``l2_norm`` and ``div_norm``
(R/src/postprocess/solution_norms_post_processor.py:7-39, 42-78): the integrands are the error integrands of
``l2_error`` / ``div_error`` with a vanishing discrete field, so the same kernel (``nfa_field_error``) integrates them.
``devia_l2_norm`` (:81-) unpacks ``fe_space.quadrature`` as a pair, which the current reference spaces no longer allow;
it is not restated."""
import numpy as np

from .l2_error_post_processor import _field_errors


def l2_norm(dim, fe_space, functions, skip_fields=[]):
    return _field_errors(dim, fe_space, functions, np.zeros(fe_space.n_dof), skip_fields, False)


def div_norm(dim, fe_space, functions, skip_fields=[]):
    return _field_errors(dim, fe_space, functions, np.zeros(fe_space.n_dof), skip_fields, True)
