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
REPO_ROOT = Path(__file__).resolve().parents[1]


INSTALLED_ROOT = REPO_ROOT / 'baseline' / '_ref'   # pip --target install of the reference (__graft_entry__.build)


def available():
    return (REFERENCE_ROOT / 'cosypose').exists()


def installed():
    return (INSTALLED_ROOT / 'cosypose').exists() and (INSTALLED_ROOT / 'config_yann.yaml').exists()


def import_reference(force_cpu=False, use_installed=False):
    """Returns the imported `cosypose` package of the reference: the mounted tree (/root/reference) or, with
    `use_installed`, the copy pip-installed into baseline/_ref (the only one that exists on the GPU box).
    force_cpu: map `.cuda()` to identity even when a GPU is visible (the CPU arm of bench.py)."""
    if 'cosypose' in sys.modules:
        return sys.modules['cosypose']
    import numpy as np
    import torch

    root = Path(tempfile.mkdtemp(prefix='cosypose_ref_'))
    if use_installed:
        assert installed(), 'baseline/_ref is missing: run __graft_entry__.build() where /root/reference is mounted'
        os.symlink(INSTALLED_ROOT / 'cosypose', root / 'cosypose')
        shutil.copy(INSTALLED_ROOT / 'config_yann.yaml', root / 'config_yann.yaml')
        sys.path.insert(0, str(INSTALLED_ROOT))            # cosypose_cext built by the reference's own setup.py
    else:
        assert available(), 'reference tree not mounted'
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
    if force_cpu or not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self

    sys.path.insert(0, str(REPO_ROOT / 'oracle' / '_ref'))
    sys.path.insert(0, str(root))
    n_threads = torch.get_num_threads()
    import cosypose  # sets OMP/MKL env vars (cosypose/__init__.py:2-3); harmless after torch import
    torch.set_num_threads(n_threads)
    if force_cpu or not torch.cuda.is_available():
        # TensorCollection.cuda() is `.to('cuda')` (utils/tensor_collection.py:83-84), used by
        # integrated/multiview_predictor.py:80; identity on a CPU-only box (runtime patch, the
        # reference sources are untouched)
        from cosypose.utils import tensor_collection as ref_tc
        ref_tc.TensorCollection.cuda = lambda self: self
    return cosypose
