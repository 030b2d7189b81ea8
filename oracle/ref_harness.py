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


def _stub_module(name):
    """A permissive stand-in for a package the Mapper imports but never calls on the paths under test."""
    class _Any:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, n):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

    m = types.ModuleType(name)

    def _ga(n):
        if n.startswith('__'):
            raise AttributeError(n)
        return _Any

    m.__getattr__ = _ga
    sys.modules[name] = m
    return m


def build_reference_slam(model_cls=None, tmp_dir=None, active_ray=True):
    """The reference's OWN `CoSLAMNaruto` (src/slam/coslam/coslam.py:32-147), constructed on the CPU following SURVEY Appendix C:
    stubs for marching_cubes / trimesh / matplotlib / pytorch3d / open3d, namespace shims for third_parties/coslam's
    `datasets` / `model` / `tools` / `optimization` packages, a temporary data directory holding only a `traj.txt`.
    model_cls: class to swap in for `JointEncoding` at src/slam/coslam/coslam.py:22; None keeps
    the reference's.  Returns (slam, coslam_module)."""
    import tempfile
    _install_stubs()
    for name in ('marching_cubes', 'trimesh', 'matplotlib', 'matplotlib.pyplot', 'pytorch3d', 'pytorch3d.transforms', 'open3d'):
        if name not in sys.modules:
            _stub_module(name)
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.modules['pytorch3d'].transforms = sys.modules['pytorch3d.transforms']
    tp = os.path.join(REF_ROOT, 'third_parties', 'coslam')
    for name in ('datasets', 'model', 'tools', 'optimization'):       # the venv's HuggingFace `datasets` would win otherwise
        m = sys.modules.get(name)
        if m is None or getattr(m, '__path__', None) != [os.path.join(tp, name)]:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(tp, name)]
            sys.modules[name] = m
    if tp not in sys.path:
        sys.path.insert(0, tp)
    tmp = tmp_dir or tempfile.mkdtemp(prefix='nrt_ref_slam_')
    data = os.path.join(tmp, 'data')
    os.makedirs(os.path.join(data, 'results'), exist_ok=True)
    with open(os.path.join(data, 'traj.txt'), 'w') as f:
        f.write('\n'.join(' '.join(str(float(i == j)) for i in range(4) for j in range(4)) for _ in range(10)) + '\n')
    AD = sys.modules['mmengine'].Config
    main_cfg = AD(slam=AD(room_cfg='configs/Replica/office0/coslam.yaml', voxel_size=0.1, SLAMData_dir=data,
                          enable_active_planning=True, enable_active_ray=active_ray, act_ray_num_uncert_sample=500,
                          act_ray_oversample_mul=4),
                  dirs=AD(result_dir=os.path.join(tmp, 'out')), visualizer=AD())
    with _in_ref_dir():
        import src.slam.coslam.coslam as coslam_mod
        from src.utils.general_utils import InfoPrinter
        if model_cls is not None:
            coslam_mod.JointEncoding = model_cls
        sink = io.StringIO()
        # the reference's get_uncert_grid hard-codes device="cuda" (src/slam/coslam/model/scene_rep.py:54): drop the argument
        # while the object is being constructed on the CPU
        real_ones = torch.ones

        def ones_here(*a, **k):
            k.pop('device', None)
            return real_ones(*a, **k)

        torch.ones = ones_here
        try:
            with contextlib.redirect_stdout(sink):
                slam = coslam_mod.CoSLAMNaruto(main_cfg, InfoPrinter('test'))
        finally:
            torch.ones = real_ones
    return slam, coslam_mod
