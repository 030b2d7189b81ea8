import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz')))


def golden_params(g, spec):
    """Regenerate the parameters a golden file was produced with and check the checksum."""
    from oracle import naruto_oracle as no
    P = no.init_params(spec, seed=int(g['param_seed']), grid_range=float(g['grid_range']),
                       uncert_jitter=float(g['uncert_jitter']))
    assert abs(P.grid.double().sum().item() - float(g['grid_checksum'])) < 1e-9, 'param generator drifted'
    assert abs(P.w1.double().sum().item() - float(g['w1_checksum'])) < 1e-9
    return P


def t(a, device='cpu'):
    return torch.from_numpy(np.asarray(a)).to(device)


@pytest.fixture(scope='session')
def spec():
    from oracle import naruto_oracle as no
    return no.office0_spec()
