"""CPU: the C-ABI library loads, exports every symbol include/naruto_b200.h declares, and its level table is the
oracle's bit for bit.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, 'include', 'naruto_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(nrt_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from naruto_b200 import _lib
    lib = _lib.load()
    names = _header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/naruto_b200.h but not exported'
    assert sorted(_lib.SIGNATURES) == names, 'ctypes signatures and header disagree'
    assert lib.nrt_abi_version() == _lib.NRT_ABI_VERSION == 4


def test_plan_level_table_matches_oracle(spec):
    from naruto_b200.configs import replica_office0
    from naruto_b200.field import FieldPlan
    plan = FieldPlan(replica_office0(), spec.bound)
    lv = plan.levels()
    for a, b in zip(lv, spec.table):
        assert np.float32(a['scale']) == np.float32(b['scale'])
        assert (a['res'], a['size'], a['offset']) == (b['res'], b['size'], b['offset'])
    assert plan.n_grid_floats == 1628176 and plan.S == 43 and plan.uncert_dims == [49, 56, 35]
    assert plan.resolution_sdf == 275


def test_plan_large_table(spec):
    # 2^21-entry levels, uncert grid [204,69,66].  (resolution_sdf is 1015, not the survey's
    # 1014: the reference divides a float32 dim_max -- 20.300001 / 0.02 -- tp/model/scene_rep.py:22-28.)
    from naruto_b200.configs import mp3d_large, MP3D_LARGE_BOUND
    from naruto_b200.field import FieldPlan
    from oracle import naruto_oracle as no
    plan = FieldPlan(mp3d_large(), MP3D_LARGE_BOUND)
    assert plan.resolution_sdf == 1015 and plan.S == 192
    assert plan.uncert_dims == [204, 69, 66]
    o = no.office0_spec(n_samples_d=181, log2_hashmap_size=21, bound=MP3D_LARGE_BOUND)
    for a, b in zip(plan.levels(), o.table):
        assert np.float32(a['scale']) == np.float32(b['scale']) and a['size'] == b['size'] and a['offset'] == b['offset']
    assert plan.n_grid_floats == o.n_grid_entries * 2 and plan.n_grid_floats * 4 > 126e6      # table larger than L2
    assert [lv['size'] for lv in plan.levels()][8:] == [1 << 21] * 8


def test_plan_rejects_unsupported_config():
    from naruto_b200 import _lib
    from naruto_b200.configs import replica_office0, OFFICE0_BOUND
    from naruto_b200.field import FieldPlan
    cfg = replica_office0()
    cfg['decoder']['hidden_dim'] = 64
    with pytest.raises(_lib.NrtError):
        FieldPlan(cfg, OFFICE0_BOUND)
    cfg = replica_office0()
    cfg['grid']['enc'] = 'DenseGrid'
    with pytest.raises(_lib.NrtError):
        FieldPlan(cfg, OFFICE0_BOUND)


def test_no_cpu_fallback():
    import torch
    from naruto_b200 import _lib
    from naruto_b200.configs import replica_office0, OFFICE0_BOUND
    from naruto_b200.scene_rep import JointEncodingNaruto
    m = JointEncodingNaruto(replica_office0(), torch.tensor(OFFICE0_BOUND))
    with pytest.raises(_lib.NrtError):
        m.query_sdf(torch.rand(4, 3), embed=True)      # CPU tensor: must refuse, not fall back


def test_state_dict_keys_match_reference():
    import torch
    from naruto_b200.configs import replica_office0, OFFICE0_BOUND
    from naruto_b200.scene_rep import JointEncodingNaruto
    m = JointEncodingNaruto(replica_office0(), torch.tensor(OFFICE0_BOUND))
    m.uncert_grid = torch.nn.Parameter(torch.ones(49, 56, 35) * 3)     # get_uncert_grid allocates on cuda
    want = [ln.split(' ', 1) for ln in open(os.path.join(ROOT, 'tests', 'golden', 'state_dict_keys.txt')).read().splitlines()]
    got = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert sorted(got) == sorted(k for k, _ in want)
    for k, shp in want:
        assert str(got[k]) == shp, k
