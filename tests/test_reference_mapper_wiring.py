"""CPU: the reference's OWN Mapper class (CoSLAMNaruto, imported unmodified from /root/reference) constructed around the
drop-in scene model -- the one-import swap of INTEGRATION.md section 1 -- and what it does with the model outside the kernels:
create_optimizer, init_uncert_grid_optim, freeze_model, save_ckpt / load_ckpt.  Skipped where the reference tree is absent
(the GPU box); the kernels themselves are covered by the `-m gpu` tests."""
import os

import pytest
import torch

from oracle import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason='needs /root/reference (build container only)')
GOLD_KEYS = os.path.join(os.path.dirname(__file__), 'golden', 'state_dict_keys.txt')


@pytest.fixture(scope='module')
def slam(tmp_path_factory):
    from naruto_b200.scene_rep import JointEncodingNaruto
    s, mod = rh.build_reference_slam(JointEncodingNaruto, tmp_dir=str(tmp_path_factory.mktemp('slam')))
    return s


def test_reference_mapper_builds_around_the_drop_in_model(slam):
    from naruto_b200.scene_rep import JointEncodingNaruto
    assert type(slam.model) is JointEncodingNaruto
    # create_optimizer (src/slam/coslam/coslam.py:409-419): decoder group (4 weights, wd 1e-6), grid group (eps 1e-15)
    g_dec, g_grid = slam.map_optimizer.param_groups
    assert len(g_dec['params']) == 4 and g_dec['weight_decay'] == 1e-6 and g_dec['lr'] == slam.config['mapping']['lr_decoder']
    assert len(g_grid['params']) == 1 and g_grid['eps'] == 1e-15 and g_grid['params'][0] is slam.model.embed_fn.params
    assert slam.map_optimizer.defaults['betas'] == (0.9, 0.99)
    # init_uncert_grid_optim (:240-243): Adam(lr=1) over the uncertainty grid the model hands out
    (g_unc,) = slam.uncert_optim.param_groups
    assert g_unc['lr'] == 1 and g_unc['params'][0] is slam.model.uncert_grid
    assert tuple(slam.model.uncert_grid.shape) == (49, 56, 35) and float(slam.model.uncert_grid.min()) == 3.0
    assert slam.model.cache_uncert.shape == (49, 56, 35)
    slam.freeze_model()                        # sets a misspelt attribute: must not raise
    assert all(p.requires_grad for p in slam.model.decoder.parameters())


def test_state_dict_keys_and_checkpoint_round_trip(slam):
    keys = [ln.split()[0] for ln in open(GOLD_KEYS).read().splitlines() if ln.strip()]
    assert sorted(slam.model.state_dict().keys()) == sorted(keys)
    slam.est_c2w_data[0] = torch.eye(4)
    slam.est_c2w_data_rel[0] = torch.eye(4)
    slam.save_ckpt(7)
    path = os.path.join(slam.main_cfg.dirs.result_dir, 'coslam', 'checkpoint', 'ckpt_0007.pt')
    assert os.path.exists(path)
    before = {k: v.clone() for k, v in slam.model.state_dict().items()}
    with torch.no_grad():
        slam.model.embed_fn.params.add_(1.0)
        slam.model.decoder.sdf_net.model[0].weight.zero_()
    slam.load_ckpt(path)
    for k, v in slam.model.state_dict().items():
        assert torch.equal(v, before[k]), k


def test_forward_on_cpu_fails_loudly(slam):
    """There is no CPU fallback behind the drop-in class: the Mapper's model.forward on host tensors raises."""
    from naruto_b200._lib import NrtError
    slam.model.train()
    with pytest.raises(NrtError):
        slam.model.forward(torch.zeros(4, 3), torch.ones(4, 3), torch.zeros(4, 3), torch.ones(4, 1))
    with pytest.raises(NrtError):
        slam.model.query_sdf(torch.zeros(4, 3))
