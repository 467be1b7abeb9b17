"""ctypes binding of libcosyb200.so (C-ABI declared in include/cosyb200.h).

There is no CPU or PyTorch fallback: if the shared library is missing, or a call fails, the
product raises.  Build it with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C cosypose_b200/csrc`.
"""
import ctypes
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p
from pathlib import Path

import os

LIB_PATH = Path(os.environ.get('COSYB200_LIB') or Path(__file__).resolve().parent / 'libcosyb200.so')

EINVAL, ECUDA, ESTATE, ENOMEM = -1, -2, -3, -4
SLOT_COARSE, SLOT_REFINER = 0, 1


class EngineError(RuntimeError):
    pass


_lib = None

_P = c_void_p
_SIGS = {
    'cosyb200_version': ([], c_int),
    'cosyb200_create': ([POINTER(c_void_p), c_int, c_int], c_int),
    'cosyb200_destroy': ([_P], c_int),
    'cosyb200_effnet_block': ([c_int, POINTER(c_int32)], c_int),
    'cosyb200_launch_plan': ([c_int, c_int, POINTER(c_int32)], c_int),
    'cosyb200_pw2_plan': ([c_int, c_int, c_int, c_int, POINTER(c_int32)], c_int),
    'cosyb200_load_pose_model': ([_P, c_int, c_int, POINTER(c_char_p), POINTER(c_void_p), POINTER(c_int64)], c_int),
    'cosyb200_set_meshes': ([_P, c_int, c_int, _P, c_int, _P, c_int, _P, _P, _P], c_int),
    'cosyb200_tco_init': ([_P, c_int, c_int, _P, _P, _P, _P, _P], c_int),
    'cosyb200_prepare_iter': ([_P, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P], c_int),
    'cosyb200_roi_crop': ([_P, c_int, _P, c_int, c_int, c_int, _P, _P, _P, _P], c_int),
    'cosyb200_net_forward': ([_P, c_int, c_int, _P, _P, _P, POINTER(c_void_p), _P], c_int),
    'cosyb200_update_pose': ([_P, c_int, _P, _P, _P, _P, _P], c_int),
    'cosyb200_refine_iter': ([_P, c_int, c_int, _P, c_int, c_int, c_int, _P, _P, _P, c_int, _P, _P, _P, _P, _P], c_int),
    'cosyb200_refine_n': ([_P, c_int, c_int, c_int, _P, c_int, c_int, c_int, _P, _P, _P, _P, c_int, _P,
                           _P, _P, _P, _P, _P, _P], c_int),
    'cosyb200_set_render_meshes': ([_P, c_int, c_int64, _P, _P, c_int64, _P, _P], c_int),
    'cosyb200_render': ([_P, c_int, _P, _P, _P, _P, c_int, _P, _P], c_int),
    'cosyb200_set_option': ([_P, c_char_p, c_int], c_int),
    'cosyb200_nccl_unique_id': ([_P], c_int),
    'cosyb200_nccl_comm_init': ([_P, c_int, c_int, _P], c_int),
    'cosyb200_nccl_comm_destroy': ([_P], c_int),
    'cosyb200_allgather_candidates': ([_P, _P, _P, c_int64, _P], c_int),
    'cosyb200_debug_pointwise': ([_P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, c_int, _P, c_int, _P, _P], c_int),
    'cosyb200_debug_trace': ([_P, _P], c_int),
    'cosyb200_debug_dump': ([_P, c_int, _P, _P, _P], c_int),
    'cosyb200_profile_enable': ([_P, c_int], c_int),
    'cosyb200_profile_read': ([_P, c_int, _P, _P], c_int),
    'cosyb200_profile_read_blocks': ([_P, c_int, _P], c_int),
    'cosyb200_ransac_infos': ([c_int, _P, _P, c_int, c_int, POINTER(c_int64), POINTER(c_int64), _P, _P], c_int),
    'cosyb200_ransac_models': ([_P, c_int64, _P, _P, _P, _P, _P], c_int),
    'cosyb200_ransac_score': ([_P, c_int64, _P, _P, _P, _P, _P, _P], c_int),
    'cosyb200_symmetric_distance': ([_P, c_int64, _P, _P, _P, _P, _P, _P], c_int),
    'cosyb200_compose_inv': ([_P, c_int64, _P, _P, _P, _P, _P, _P], c_int),
    'cosyb200_ba_linearize': ([_P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_float,
                               _P, _P, _P, _P, _P, _P, _P, _P], c_int),
    'cosyb200_ba_linearize_f64': ([_P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_float,
                                   _P, _P, _P, _P, _P, _P], c_int),
    'cosyb200_lm_solve': ([_P, c_int, _P, _P, c_double, _P, _P, _P], c_int),
    'cosyb200_ransac_inliers': ([c_int64, _P, _P, c_int64, _P, _P, _P, _P, c_float, c_int, _P, _P,
                                 POINTER(c_int64), _P, POINTER(c_int64)], c_int),
    'cosyb200_ransac_inliers_dev': ([_P, c_int64, c_int64, _P, c_int64, _P, _P, _P, _P, c_float, c_int, _P, _P, _P, _P, _P], c_int),
    'cosyb200_pose_errors': ([_P, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P], c_int),
    'cosyb200_scatter_argmin': ([c_int64, _P, _P, c_int64, _P], c_int),
    'cosyb200_expand_ids_for_symmetry': ([c_int64, _P, _P, POINTER(c_int64), _P, _P], c_int),
}
EXPORTS = sorted(list(_SIGS) + ['cosyb200_last_error'])


def lib():
    """The loaded library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f'{LIB_PATH} is missing: the CUDA engine has not been built '
                f'(run `make -C {LIB_PATH.parent / "csrc"}`); cosypose_b200 has no CPU fallback.')
        L = ctypes.CDLL(str(LIB_PATH))
        L.cosyb200_last_error.argtypes = []
        L.cosyb200_last_error.restype = c_char_p
        for name, (args, res) in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = res
        _lib = L
    return _lib


def check(rc, what=''):
    """Maps C status codes onto the reference's exception conventions
    (assert -> AssertionError, unsupported value -> ValueError)."""
    if rc == 0:
        return
    msg = lib().cosyb200_last_error().decode('utf-8', 'replace')
    if rc == EINVAL:
        raise AssertionError(f'{what}: {msg}')
    raise EngineError(f'{what}: rc={rc}: {msg}')
