"""Symmetry sets of BOP object models, as [n_sym, 4, 4] matrices.

Semantics of the reference's `make_bop_symmetries` (cosypose/lib3d/symmetries.py:7-35, after bop_toolkit_lib/misc.py
`get_symmetry_transformations`): the set is {C_j D_i}, i-major, where D_0 = I and D_1.. are the discrete
symmetries listed in `models_info.json` (translations in model units, scaled to metres) and C_j are
`n_symmetries_continuous` equally spaced rotations about each continuous-symmetry axis (a coordinate axis, zero
offset); without continuous symmetries the set is just {D_i}.  Built with numpy broadcasting; no pinocchio /
transforms3d."""
import numpy as np

# index pairs (i, j) of the plane each coordinate axis rotates: x -> (y, z), y -> (z, x), z -> (x, y)
_PLANE = ((1, 2), (2, 0), (0, 1))


def _rotations_about(axis_index, n):
    """[n, 4, 4]: rotations by 2 pi k / n, k = 0..n-1, about coordinate axis `axis_index`."""
    ang = 2.0 * np.pi * np.arange(n) / n
    i, j = _PLANE[axis_index]
    R = np.tile(np.eye(4), (n, 1, 1))
    R[:, i, i] = R[:, j, j] = np.cos(ang)
    R[:, j, i] = np.sin(ang)
    R[:, i, j] = -R[:, j, i]
    return R


def make_bop_symmetries(dict_symmetries, n_symmetries_continuous=8, scale=0.001):
    discrete = [np.eye(4)]
    for flat in dict_symmetries.get('symmetries_discrete', []):
        D = np.asarray(flat, dtype=np.float64).reshape(4, 4).copy()
        D[:3, 3] *= scale
        discrete.append(D)
    discrete = np.stack(discrete)                                                   # [nd, 4, 4]

    continuous = []
    for entry in dict_symmetries.get('symmetries_continuous', []):
        axis = np.asarray(entry['axis'])
        if not np.allclose(entry['offset'], 0) or axis.sum() != 1:
            raise AssertionError('continuous symmetries must be about a coordinate axis through the origin')
        continuous.append(_rotations_about(int(np.argmax(axis)), n_symmetries_continuous))
    if not continuous:
        return discrete
    continuous = np.concatenate(continuous, axis=0)                                 # [nc, 4, 4]
    # all products C_j @ D_i, discrete-major
    return np.einsum('jab,ibc->ijac', continuous, discrete).reshape(-1, 4, 4)
