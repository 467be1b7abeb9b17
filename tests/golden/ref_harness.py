"""Imports the *unmodified* reference (/root/reference) on CPU so golden vectors can
be generated from it.  Only usable in the build container: the GPU box has no
/root/reference, and nothing in tests/, bench.py or smoke() imports this module.

What is needed to import the reference's hot path:
  * a writable project root holding a symlink to the read-only package,
    `config_yann.yaml` and an empty `local_data/` (cosypose/config.py:33-53);
  * CONDA_PREFIX set to anything (cosypose/config.py:45);
  * empty stand-ins for pinocchio / eigenpy / transforms3d / trimesh, imported at
    module scope by lib3d but never executed on this path;
  * `np.int` (removed in numpy 1.24, used by multiview/ransac.py:94,102,122);
  * the reference extension `cosypose_cext`, built by oracle/Makefile into oracle/_ref;
  * `Tensor.cuda()` mapped to identity when no GPU is present (hard-coded at
    multiview/bundle_adjustment.py:221, integrated/multiview_predictor.py:16-18,80).
"""
import os
import shutil
import sys
import tempfile
import types
from pathlib import Path

REFERENCE_ROOT = Path('/root/reference')
REPO_ROOT = Path(__file__).resolve().parents[2]


def available():
    return (REFERENCE_ROOT / 'cosypose').exists()


def import_reference():
    """Returns the imported `cosypose` package of the reference."""
    if 'cosypose' in sys.modules:
        return sys.modules['cosypose']
    assert available(), 'reference tree not mounted'
    import numpy as np
    import torch

    root = Path(tempfile.mkdtemp(prefix='cosypose_ref_'))
    os.symlink(REFERENCE_ROOT / 'cosypose', root / 'cosypose')
    shutil.copy(REFERENCE_ROOT / 'config_yann.yaml', root / 'config_yann.yaml')
    (root / 'local_data').mkdir()
    os.environ.setdefault('CONDA_PREFIX', str(root))
    sys.dont_write_bytecode = True

    for name in ('pinocchio', 'eigenpy', 'transforms3d', 'trimesh'):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            sys.modules[name] = mod
    sys.modules['eigenpy'].switchToNumpyArray = lambda: None
    if not hasattr(np, 'int'):
        np.int = int
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self

    sys.path.insert(0, str(REPO_ROOT / 'oracle' / '_ref'))
    sys.path.insert(0, str(root))
    n_threads = torch.get_num_threads()
    import cosypose  # sets OMP/MKL env vars (cosypose/__init__.py:2-3); harmless after torch import
    torch.set_num_threads(n_threads)
    if not torch.cuda.is_available():
        # TensorCollection.cuda() is `.to('cuda')` (utils/tensor_collection.py:83-84), used by
        # integrated/multiview_predictor.py:80; identity on a CPU-only box (runtime patch, the
        # reference sources are untouched)
        from cosypose.utils import tensor_collection as ref_tc
        ref_tc.TensorCollection.cuda = lambda self: self
    return cosypose
