"""Symmetry sets of BOP object models (reference: cosypose/lib3d/symmetries.py:7-35, itself after
bop_toolkit_lib/misc.py `get_symmetry_transformations`): the identity plus the discrete symmetries of
`models_info.json`, each composed with `n_symmetries_continuous` rotations about the continuous-symmetry
axis.  numpy only (the reference goes through pinocchio's SE3 and transforms3d quaternions for the same
matrices)."""
import numpy as np


def _axis_rotation(axis, angle):
    """Rotation by `angle` about coordinate axis `axis` (a one-hot vector: the reference asserts axis.sum() == 1
    and builds euler = axis * angle in 'sxyz', i.e. a rotation about that single axis)."""
    k = int(np.argmax(axis))
    c, s = np.cos(angle), np.sin(angle)
    M = np.eye(4)
    i, j = [(1, 2), (2, 0), (0, 1)][k]
    M[i, i], M[i, j], M[j, i], M[j, j] = c, -s, s, c
    return M


def make_bop_symmetries(dict_symmetries, n_symmetries_continuous=8, scale=0.001):
    """-> [n_sym, 4, 4] float64; translations of the discrete symmetries are scaled (mm -> m)."""
    sym_discrete = dict_symmetries.get('symmetries_discrete', [])
    sym_continuous = dict_symmetries.get('symmetries_continuous', [])
    all_discrete = [np.eye(4)]
    for sym_n in sym_discrete:
        M = np.array(sym_n, dtype=np.float64).reshape(4, 4).copy()
        M[:3, -1] *= scale
        all_discrete.append(M)
    all_continuous = []
    for sym_n in sym_continuous:
        assert np.allclose(sym_n['offset'], 0)
        axis = np.array(sym_n['axis'])
        assert axis.sum() == 1
        for n in range(n_symmetries_continuous):
            all_continuous.append(_axis_rotation(axis, 2 * np.pi * n / n_symmetries_continuous))
    out = []
    for sym_d in all_discrete:
        if all_continuous:
            out.extend(sym_c @ sym_d for sym_c in all_continuous)
        else:
            out.append(sym_d)
    return np.array(out)
