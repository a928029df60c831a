"""Error norms of a discrete solution -- GPU versions of ``l2_error``, ``l2_error_projected`` and ``div_error``
(R/src/postprocess/l2_error_post_processor.py:7-63, 66-117, 185-237): same signatures, same return value (one norm per
field, in field order).  ``grad_error`` (:120-182) and ``devia_l2_error`` (:308-) read attributes the current reference
spaces no longer have (``el_data.quadrature``, ``fe_space.quadrature`` as a pair) and are not restated.  The exact functions are evaluated on the host over all quadrature points at once, the discrete field
is evaluated and compared on the device (``nfa_field_error``), cell contributions are summed in cell order."""
import numpy as np
import torch

from .. import _lib, coefficients
from ..basis.orientation import case_matrices, edge_dof_signs
from ..device import space_device
from ..weak_forms.laplace_dual_weak_form import _all_cases


def _exact_values(fn, x, shape_tail):
    """``fn(x, y, z)`` over all points -> ``[n, n_q, *shape_tail]`` (the reference indexes the leading axes of what the
    callable returns: components first, points last)."""
    n, nq, _ = x.shape
    flat = x.reshape(-1, 3)
    try:
        val = np.asarray(fn(flat[:, 0], flat[:, 1], flat[:, 2]), dtype=np.float64)
        if val.shape[-1] != n * nq:
            raise ValueError
    except Exception:
        val = np.stack([np.asarray(fn(p[0], p[1], p[2]), dtype=np.float64) for p in flat], axis=-1)
    val = np.moveaxis(val, -1, 0)  # [points, ...]
    idx = tuple(slice(0, k) for k in shape_tail)
    return np.ascontiguousarray(val[(slice(None),) + idx]).reshape((n, nq) + tuple(shape_tail))


def _field_errors(dim, fe_space, functions, alpha, skip_fields, divergence):
    dev = space_device(fe_space)
    points, weights = fe_space.quadrature[dim]
    up = lambda a, dt=np.float64: torch.as_tensor(np.ascontiguousarray(a, dtype=dt)).to(dev.device)
    alpha_dev = up(alpha)
    dest = dev.dest()
    strides, out = {}, []
    off = 0
    for name, s in fe_space.discrete_spaces.items():
        strides[name] = off
        off += s.ref_element.n_phi * s.n_comp
    w = up(weights)
    for name, s in fe_space.discrete_spaces.items():
        if name in skip_fields:
            continue
        vector_family = s.family in ("RT", "BDM")
        if s.family in ("N1E", "N2E") and not divergence and s.dimension == dim:
            # H(curl) field: covariant map + edge signs (csrc/nfa_hcurl.cu)
            ref = s.ref_element
            cell_nodes = dev.cell_nodes(name)
            n = int(cell_nodes.shape[0])
            x = coefficients.physical_points(dev, dim, points, cell_nodes)
            flat = x.reshape(-1, 3)
            val = np.asarray(functions[name](flat[:, 0], flat[:, 1], flat[:, 2]), dtype=np.float64)
            exact = np.ascontiguousarray(val.reshape(-1, 3, flat.shape[0])[0].T).reshape(n, len(weights), 3)
            per_cell = torch.empty(n, dtype=torch.float64, device=dev.device)
            tab, sg, ex = up(ref.tabulate(0, points)[0]), up(edge_dof_signs(s.mesh, dim, s.mesh_topology.selected_local, ref), np.int8), up(exact)
            _lib.call(
                "nfa_hcurl_field_error", dim, ref.n_phi, len(weights), _lib.ptr(tab), _lib.ptr(w), _lib.ptr(sg),
                _lib.ptr(dev.coords), _lib.ptr(cell_nodes), n, _lib.ptr(alpha_dev), _lib.ptr(dest), int(dest.shape[1]),
                int(strides[name]), _lib.ptr(ex), _lib.ptr(per_cell), _lib.stream_ptr())
            out.append(float(np.sqrt(per_cell.cpu().numpy().sum())))
            continue
        if divergence and not vector_family:
            continue
        if s.dimension != dim:
            # the reference sums over the field's elements of dimension `dim` and appends sqrt(0.0) when there are none
            # (R/src/postprocess/l2_error_post_processor.py:15-19, 61): the list keeps one entry per non-skipped field
            if not divergence:
                out.append(0.0)
            continue
        if divergence == "scaled" and (functions.get("l") is None or functions.get("grad_l") is None):
            continue  # the reference skips the field when no scale is given (:296-298)
        ref = s.ref_element
        cell_nodes = dev.cell_nodes(name)
        n = int(cell_nodes.shape[0])
        x = coefficients.physical_points(dev, dim, points, cell_nodes)
        per_cell = torch.empty(n, dtype=torch.float64, device=dev.device)
        if not vector_family:
            exact = _exact_values(functions[name], x, (min(s.n_comp, dim),))
            if s.n_comp > dim:
                raise NotImplementedError("l2_error: Lagrange fields with more components than dimensions")
            tab = up(ref.tabulate(0, points)[0, :, :, 0])
            args = (0, ref.n_phi, s.n_comp, len(weights), 0, _lib.ptr(tab), None, _lib.ptr(w), None, None)
            keep = [tab]
        else:
            t = ref.tabulate(1, points)
            val, dv = up(t[0]), up(sum(t[1 + d][:, :, d] for d in range(dim)))
            tc = up(case_matrices(ref))
            cases = torch.as_tensor(_all_cases(s)).to(dev.device).contiguous()
            nfd = len(ref.entity_dofs[dim - 1][0])
            if divergence:
                exact = _exact_values(functions["div_" + name], x, (s.n_comp,))
            else:
                exact = _exact_values(functions[name], x, (s.n_comp, dim))
            kind = 3 if divergence == "scaled" else (2 if divergence else 1)
            args = (kind, ref.n_phi, s.n_comp, len(weights), nfd, _lib.ptr(val), _lib.ptr(dv),
                    _lib.ptr(w), _lib.ptr(tc), _lib.ptr(cases))
            keep = [val, dv, tc, cases]
        ex = up(exact)
        scale = gscale = None
        if divergence == "scaled":
            scale = up(coefficients.evaluate(functions["l"], x))
            gscale = up(coefficients.evaluate(functions["grad_l"], x, dim))
        _lib.call(
            "nfa_field_error_scaled", dim, *args, _lib.ptr(dev.coords), _lib.ptr(cell_nodes), n, _lib.ptr(alpha_dev),
            _lib.ptr(dest), int(dest.shape[1]), int(strides[name]), _lib.ptr(ex), _lib.ptr(scale), _lib.ptr(gscale),
            _lib.ptr(per_cell), _lib.stream_ptr(),
        )
        out.append(float(np.sqrt(per_cell.cpu().numpy().sum())))
        del keep
    return out


def l2_error(dim, fe_space, functions, alpha, skip_fields=[]):
    return _field_errors(dim, fe_space, functions, np.asarray(alpha, dtype=np.float64), skip_fields, False)


def div_error(dim, fe_space, functions, alpha, skip_fields=[]):
    return _field_errors(dim, fe_space, functions, np.asarray(alpha, dtype=np.float64), skip_fields, True)


def _zero_like_field(s, dim):
    if s.family in ("RT", "BDM"):
        return lambda x, y, z: np.zeros((s.n_comp, 3) + np.shape(x))
    return lambda x, y, z: np.zeros((max(s.n_comp, dim),) + np.shape(x))


def l2_error_projected(dim, fe_space, alpha, skip_fields=[], dof_shift=0):
    """L2 norm of every discrete field itself (the reference's "error" against zero,
    R/src/postprocess/l2_error_post_processor.py:66-117); ``dof_shift`` offsets the destination indices into ``alpha``."""
    zero = {name: _zero_like_field(s, dim) for name, s in fe_space.discrete_spaces.items()}
    return _field_errors(dim, fe_space, zero, np.asarray(alpha, dtype=np.float64)[dof_shift:], skip_fields, False)


def div_scaled_error(dim, fe_space, functions, alpha, skip_fields=[]):
    """Error of the scaled divergence ``l div s + grad l . s`` of every H(div) field against ``functions["div_" + name]``
    (R/src/postprocess/l2_error_post_processor.py:240-305); fields are skipped when ``l`` / ``grad_l`` are missing."""
    return _field_errors(dim, fe_space, functions, np.asarray(alpha, dtype=np.float64), skip_fields, "scaled")
