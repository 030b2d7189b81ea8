"""TEST INFRASTRUCTURE -- imports the reference's OWN Python (read-only, /root/reference) on CPU.

Follows SURVEY.md Appendix C: `tinycudann` is replaced in sys.modules by oracle/tcnn_shim.py and
`mmengine` by a stub, after which src/slam/coslam/model/scene_rep.py:JointEncodingNaruto constructs
and runs unmodified.  Only usable in the build container (the GPU box has no /root/reference); it is
used by oracle/make_golden.py to produce tests/golden/*.npz and by tests that are skipped when the
reference tree is absent.  Never imported by naruto_b200/.
"""
import contextlib
import io
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = '/root/reference'


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, 'src', 'slam'))


def _install_stubs():
    from oracle import tcnn_shim
    sys.modules['tinycudann'] = tcnn_shim
    if 'mmengine' not in sys.modules:
        mm = types.ModuleType('mmengine')

        class _AttrDict(dict):
            __getattr__ = dict.get

        mm.Config = _AttrDict
        mm.ConfigDict = _AttrDict
        sys.modules['mmengine'] = mm
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


@contextlib.contextmanager
def _in_ref_dir():
    cwd = os.getcwd()
    os.chdir(REF_ROOT)            # configs use relative inherit_from paths (src/utils/config_utils.py:46-49)
    try:
        yield
    finally:
        os.chdir(cwd)


def load_reference_config(rel_path='configs/Replica/office0/coslam.yaml'):
    _install_stubs()
    with _in_ref_dir():
        from src.utils.config_utils import load_config
        return load_config(rel_path)


def build_reference_model(cfg, uncert_voxel=0.1, quiet=True):
    """JointEncodingNaruto(cfg, bound) on CPU with the uncertainty grid attached the way
    get_uncert_grid (src/slam/coslam/model/scene_rep.py:49-56) would, minus its hard-coded device="cuda"."""
    _install_stubs()
    with _in_ref_dir():
        from src.slam.coslam.model.scene_rep import JointEncodingNaruto
        bound = torch.tensor(cfg['mapping']['bound'], dtype=torch.float32)
        sink = io.StringIO()
        with (contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext()):
            model = JointEncodingNaruto(cfg, bound)
    dims = [round((bound[i, 1] - bound[i, 0]).item() / uncert_voxel + 0.0005) + 1 for i in range(3)]
    model.uncert_grid = nn.Parameter(torch.ones(dims, dtype=torch.float32) * 3)
    return model
