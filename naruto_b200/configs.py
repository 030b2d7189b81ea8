"""Scene configurations in the reference's nested-dict form (what load_config returns for
configs/Replica/office0/coslam.yaml -> configs/Replica/replica_coslam.yaml), hard-wired so that tests and the
benchmark run where the reference tree is absent.  Synthetic variants follow SURVEY.md section 8(d)."""
import copy

_REPLICA = {
    'dataset': 'replica',
    'data': {'downsample': 1, 'sc_factor': 1, 'translation': 0, 'num_workers': 4},
    'mapping': {'sample': 2048, 'first_mesh': True, 'iters': 10, 'lr_embed': 0.01, 'lr_decoder': 0.01, 'lr_rot': 0.001,
                'lr_trans': 0.001, 'keyframe_every': 5, 'map_every': 5, 'n_pixels': 0.05, 'first_iters': 200,
                'optim_cur': True, 'min_pixels_cur': 100, 'map_accum_step': 1, 'pose_accum_step': 5, 'map_wait_step': 0,
                'filter_depth': True, 'active_ray': False},
    'tracking': {'disable': True},
    'grid': {'enc': 'HashGrid', 'tcnn_encoding': True, 'hash_size': 16, 'voxel_color': 0.08, 'voxel_sdf': 0.02, 'oneGrid': True},
    'pos': {'enc': 'OneBlob', 'n_bins': 16},
    'decoder': {'geo_feat_dim': 15, 'hidden_dim': 32, 'num_layers': 2, 'num_layers_color': 2, 'hidden_dim_color': 32,
                'tcnn_network': False, 'pred_uncert': False, 'uncert_grid': True},
    'cam': {'H': 680, 'W': 1200, 'fx': 600.0, 'fy': 600.0, 'cx': 599.5, 'cy': 339.5, 'png_depth_scale': 6553.5,
            'crop_edge': 0, 'near': 0, 'far': 5, 'depth_trunc': 100.0},
    'training': {'rgb_weight': 5.0, 'depth_weight': 0.1, 'sdf_weight': 1000, 'fs_weight': 10, 'uncert_weight': 0.005,
                 'eikonal_weight': 0, 'smooth_weight': 0.000001, 'smooth_pts': 32, 'smooth_vox': 0.1, 'smooth_margin': 0.05,
                 'n_samples_d': 32, 'range_d': 0.1, 'n_range_d': 11, 'n_importance': 0, 'perturb': 1, 'white_bkgd': False,
                 'trunc': 0.1, 'rot_rep': 'axis_angle', 'rgb_missing': 0.05},
    'mesh': {'resolution': 512, 'render_color': False, 'vis': 500, 'voxel_eval': 0.05, 'voxel_final': 0.02},
}

OFFICE0_BOUND = [[-2.2, 2.6], [-3.4, 2.1], [-1.4, 2.0]]                 # configs/Replica/office0/coslam.yaml:3
MP3D_LARGE_BOUND = [[-16.2, 4.1], [-5.5, 1.3], [-0.5, 6.0]]             # configs/MP3D/YmJkqBEsHnH/coslam.yaml:3


def replica_office0(n_samples_d=32, hash_size=16, bound=None, perturb=1):
    """office0 as shipped (32+11 samples).  n_samples_d=117 gives the 128-samples/ray benchmark shape."""
    cfg = copy.deepcopy(_REPLICA)
    cfg['mapping']['bound'] = copy.deepcopy(OFFICE0_BOUND if bound is None else bound)
    cfg['mapping']['marching_cubes_bound'] = copy.deepcopy(cfg['mapping']['bound'])
    cfg['training']['n_samples_d'] = n_samples_d
    cfg['training']['perturb'] = perturb
    cfg['grid']['hash_size'] = hash_size
    return cfg


def mp3d_large(n_samples_d=181, hash_size=21):
    """largest shipped MP3D bound, 2^21-entry levels (153.8 MB table), 192 samples/ray."""
    return replica_office0(n_samples_d=n_samples_d, hash_size=hash_size, bound=MP3D_LARGE_BOUND)
