"""ORACLE (test infrastructure only -- never imported by the product path).

``l2_error`` / ``l2_error_projected`` / ``div_error`` restated cell by cell from
R/src/postprocess/l2_error_post_processor.py:7-63, 66-117, 185-237; ``l2_norm`` / ``div_norm`` from
R/src/postprocess/solution_norms_post_processor.py:7-39, 42-78."""
import numpy as np

from .ref_forms import RefElement


def _cell_tables(space, field_space, i):
    el = RefElement(field_space, field_space.element_ids[i])
    dim = el.cell.dimension
    points, weights = space.quadrature[dim]
    x, jac, det_jac, inv_jac = el.evaluate_mapping(points)
    return dim, np.asarray(weights), x, det_jac, el.evaluate_basis(points, jac, det_jac, inv_jac)


def l2_error(dim, space, functions, alpha, skip_fields=()):
    out = []
    for name, fs in space.discrete_spaces.items():
        if name in skip_fields:
            continue
        total = 0.0
        for i in range(len(fs.element_ids)):
            _, weights, x, det_jac, tab = _cell_tables(space, fs, i)
            alpha_l = alpha[space.destination_indexes(i, name)]
            n_phi = tab.shape[2]
            alpha_star = np.array(np.split(alpha_l, n_phi))
            exact = functions[name]
            if fs.family == "Lagrange":
                f_e = np.asarray(exact(x[:, 0], x[:, 1], x[:, 2]))[0:dim, :]
                f_h = (tab[0, :, :, 0] @ alpha_star[:, 0:dim]).T
                total += np.sum(det_jac * weights * (f_e - f_h) * (f_e - f_h))
            else:
                f_e = np.array([exact(xv[0], xv[1], xv[2]) for xv in x])[:, :, 0:dim]
                f_h = np.array([np.vstack([tab[0, k, :, 0:dim].T @ alpha_star[:, c] for c in range(fs.n_comp)])
                                for k in range(len(weights))])
                diff = f_e - f_h
                total += np.sum(det_jac * weights * np.array([np.trace(e.T @ e) for e in diff]))
        out.append(np.sqrt(total))
    return out


def div_error(dim, space, functions, alpha, skip_fields=()):
    out = []
    for name, fs in space.discrete_spaces.items():
        if fs.family not in ("RT", "BDM") or name in skip_fields:
            continue
        exact = functions["div_" + name]
        total = 0.0
        for i in range(len(fs.element_ids)):
            _, weights, x, det_jac, tab = _cell_tables(space, fs, i)
            alpha_l = alpha[space.destination_indexes(i, name)]
            alpha_star = np.array(np.split(alpha_l, tab.shape[2]))
            div_star = np.trace(tab[1 : 1 + dim, :, :, 0:dim], axis1=0, axis2=3).T * (1 / det_jac)
            f_e = np.array([exact(xv[0], xv[1], xv[2]) for xv in x])
            f_h = div_star.T @ alpha_star
            total += np.sum(((f_e - f_h) * (f_e - f_h)).T @ (det_jac * weights))
        out.append(np.sqrt(total))
    return out


def l2_error_projected(dim, space, alpha, skip_fields=(), dof_shift=0):
    out = []
    for name, fs in space.discrete_spaces.items():
        if name in skip_fields:
            continue
        total = 0.0
        for i in range(len(fs.element_ids)):
            _, weights, x, det_jac, tab = _cell_tables(space, fs, i)
            alpha_l = alpha[space.destination_indexes(i, name) + dof_shift]
            alpha_star = np.array(np.split(alpha_l, tab.shape[2]))
            if fs.family == "Lagrange":
                f_h = (tab[0, :, :, 0] @ alpha_star[:, 0:dim]).T
                total += np.sum(det_jac * weights * f_h * f_h)
            else:
                f_h = np.array([np.vstack([tab[0, k, :, 0:dim].T @ alpha_star[:, c] for c in range(fs.n_comp)])
                                for k in range(len(weights))])
                total += np.sum(det_jac * weights * np.array([np.trace(e.T @ e) for e in f_h]))
        out.append(np.sqrt(total))
    return out


def l2_norm(dim, space, functions, skip_fields=()):
    out = []
    for name, fs in space.discrete_spaces.items():
        if name in skip_fields:
            continue
        total = 0.0
        exact = functions[name]
        for i in range(len(fs.element_ids)):
            _, weights, x, det_jac, _ = _cell_tables(space, fs, i)
            if fs.family == "Lagrange":
                f_e = np.asarray(exact(x[:, 0], x[:, 1], x[:, 2]))[0:dim, :]
                total += np.sum(det_jac * weights * f_e * f_e)
            else:
                f_e = np.array([exact(xv[0], xv[1], xv[2]) for xv in x])[:, :, 0:dim]
                total += np.sum(det_jac * weights * np.array([np.trace(e.T @ e) for e in f_e]))
        out.append(np.sqrt(total))
    return out


def div_norm(dim, space, functions, skip_fields=()):
    out = []
    for name, fs in space.discrete_spaces.items():
        if fs.family not in ("RT", "BDM") or name in skip_fields:
            continue
        exact = functions["div_" + name]
        total = 0.0
        for i in range(len(fs.element_ids)):
            _, weights, x, det_jac, _ = _cell_tables(space, fs, i)
            f_e = np.array([exact(xv[0], xv[1], xv[2]) for xv in x])
            total += np.sum((f_e * f_e).T @ (det_jac * weights))
        out.append(np.sqrt(total))
    return out


def div_scaled_error(dim, space, functions, alpha, skip_fields=()):
    """R/src/postprocess/l2_error_post_processor.py:240-305."""
    out = []
    for name, fs in space.discrete_spaces.items():
        if fs.family not in ("RT", "BDM") or name in skip_fields:
            continue
        exact = functions["div_" + name]
        scale, grad_scale = functions.get("l"), functions.get("grad_l")
        if scale is None or grad_scale is None:
            continue
        total = 0.0
        for i in range(len(fs.element_ids)):
            _, weights, x, det_jac, tab = _cell_tables(space, fs, i)
            alpha_l = alpha[space.destination_indexes(i, name)]
            n_phi = tab.shape[2]
            alpha_star = np.array(np.split(alpha_l, n_phi))
            gamma = scale(x[:, 0], x[:, 1], x[:, 2]) * np.ones(len(weights))
            grad_gamma = np.asarray(grad_scale(x[:, 0], x[:, 1], x[:, 2]))
            div_v_star = np.array([[np.trace(np.outer(grad_gamma[:, q], tab[0, q, j, 0:dim])) for q in range(len(weights))]
                                   for j in range(n_phi)])
            div_v_star = div_v_star + np.trace(tab[1 : 1 + dim, :, :, 0:dim], axis1=0, axis2=3).T * (gamma / det_jac)
            f_e = np.array([exact(xv[0], xv[1], xv[2]) for xv in x])
            f_h = div_v_star.T @ alpha_star
            total += np.sum(((f_e - f_h) * (f_e - f_h)).T @ (det_jac * weights))
        out.append(np.sqrt(total))
    return out
