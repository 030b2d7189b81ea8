#!/usr/bin/env python
"""Per-iteration ray sampling: the device-resident path (naruto_b200/ray_sampler.py) against the
reference's host procedure restated by the oracle (Python random.sample, CPU boolean index of the 816 000-pixel frame,
numpy uncertainty lookup), at the reference's sizes: 680x1200 frame, 40 800 rays per key frame, 100 key frames,
8 192 global + 400 current rays, K = 500.  Prints one JSON line."""
import json
import os
import random
import statistics
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from naruto_b200 import ray_sampler as rs   # noqa: E402
from oracle import sampler_oracle as so   # noqa: E402


def main():
    dev = 'cuda:0'
    H, W, P, n_kf = 680, 1200, 40800, 100
    cfg = {'cam': {'depth_trunc': 100.0}, 'mapping': {'sample': 2048, 'min_pixels_cur': 100, 'keyframe_every': 5, 'filter_depth': True}}
    bbox = [[-2.2, 2.6], [-3.4, 2.1], [-1.4, 2.0]]
    g = torch.Generator().manual_seed(0)
    direction = rs.camera_rays(H, W, 600.0, 600.0, 599.0, 339.0, dev)
    depth = torch.rand(1, H, W, generator=g) * 4 + 0.3
    depth[torch.rand(1, H, W, generator=g) < 0.02] = 0.0
    rgb = torch.rand(1, H, W, 3, generator=g)
    kf = rs.DeviceKeyFrameDatabase(cfg, H, W, n_kf, P, dev)
    for f in range(n_kf):
        kf.add_keyframe({'direction': direction[None], 'rgb': rgb.to(dev), 'depth': depth.to(dev), 'frame_id': 5 * f}, filter_depth=True)
    cur = rs.pack_frame(direction[None], rgb.to(dev), depth.to(dev))
    poses = torch.eye(4).repeat(n_kf + 1, 1, 1).to(dev)
    sampler = rs.DeviceActiveRaySampler(cfg, 500, 4)
    vol = torch.rand(49, 56, 35, generator=g)
    vol_d = vol.to(dev)
    ts = []
    for i in range(25):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        o, d, s, t = rs.sample_mapping_batch(kf, cur, poses, cfg, vol_d, bbox, sampler=sampler, seed=i)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    gpu_ms = statistics.median(ts[5:])
    w0 = time.perf_counter()
    for i in range(20):
        rs.sample_mapping_batch(kf, cur, poses, cfg, vol_d, bbox, sampler=sampler, seed=i)
    torch.cuda.synchronize()
    gpu_wall_ms = (time.perf_counter() - w0) / 20 * 1e3
    # the reference's host procedure (oracle restatement, same sizes)
    kf_cpu, ids_cpu = kf.rays.cpu(), kf.frame_ids.cpu()
    cur_cpu, poses_cpu, vol_np = cur.cpu(), poses.cpu(), vol.numpy()
    cs = []
    for i in range(6):
        t0 = time.perf_counter()
        idxs = random.sample(range(n_kf * P), 8192)
        rays_g, ids_g = so.sample_global(kf_cpu, ids_cpu, idxs, P)
        valid = so.valid_depth_mask(cur_cpu, 100.0)
        cur_valid = cur_cpu[valid, :]
        idx_cur = random.sample(range(len(cur_valid)), max(8192 // n_kf, 400))
        ro, rd, s_, t_ = so.assemble(rays_g, ids_g, cur_cpu, idx_cur, 5, poses_cpu)
        so.active_select(ro, rd, s_, t_, len(idx_cur), vol_np, bbox, 2048, 500, 4)
        cs.append((time.perf_counter() - t0) * 1e3)
    cpu_ms = statistics.median(cs[1:])
    print(json.dumps({'what': 'ray sampling of one mapping iteration (8192 global + 400 current rays -> 2148-ray batch)',
                      'device_ms_cuda_events': round(gpu_ms, 4), 'device_ms_wall_incl_python': round(gpu_wall_ms, 4),
                      'host_reference_procedure_ms': round(cpu_ms, 3), 'host_cores': os.cpu_count(),
                      'batch_rows': int(o.shape[0])}))


if __name__ == '__main__':
    main()
