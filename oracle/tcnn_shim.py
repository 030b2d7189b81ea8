"""TEST INFRASTRUCTURE -- torch-native stand-in for the `tinycudann` module.

PARITY UNPINNED: tiny-cuda-nn (NVlabs/tiny-cuda-nn, installed unpinned from git HEAD by
the reference: third_parties/coslam/requirements.txt:29; fallback commit 91ee479d in
third_parties/coslam/README.md:68) is neither vendored under /root/reference nor installed in
this image.  This file restates its two encodings that the reference reaches
(third_parties/coslam/model/encodings.py:31-46 HashGrid, :61-71 OneBlob) from the
published algorithm (tiny-cuda-nn include/tiny-cuda-nn/encodings/grid.h, oneblob.h,
common_device.h) as summarised in SURVEY.md Appendix A.  It *defines* parity for this repo.

Only tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs and
oracle/make_golden.py may import this file.  The product (naruto_b200/) never does.

Usage (see oracle/ref_harness.py):  sys.modules['tinycudann'] = oracle.tcnn_shim
"""
import math

import numpy as np
import torch
import torch.nn as nn

# coherent_prime_hash primes (tiny-cuda-nn common_device.h); dim0 uses 1 so that x stays linear
PRIMES = (1, 2654435761, 805459861)


def grid_scale(level, per_level_scale, base_resolution):
    """tcnn grid_scale(): exp2(level*log2(s))*base - 1.  Evaluated in float64 and rounded once to
    float32 so that every implementation in this repo (oracle, C library) gets the same bits."""
    return float(np.float32(np.exp2(level * np.log2(np.float64(per_level_scale))) * base_resolution - 1.0))


def grid_resolution(scale):
    """tcnn grid_resolution(): ceil(scale) + 1 grid vertices per axis."""
    return int(math.ceil(scale)) + 1


def level_table(n_levels, base_resolution, per_level_scale, log2_hashmap_size):
    """Per-level (scale, resolution, n_entries, entry_offset) following the GridEncoding ctor:
    entries = min(next_multiple(res^3, 8), 2^T)."""
    table, offset = [], 0
    for lvl in range(n_levels):
        scale = grid_scale(lvl, per_level_scale, base_resolution)
        res = grid_resolution(scale)
        dense = min(res ** 3, (2 ** 32 - 1) // 2)
        dense = (dense + 7) // 8 * 8
        size = min(dense, 1 << log2_hashmap_size)
        table.append(dict(scale=scale, res=res, size=size, offset=offset))
        offset += size
    return table, offset


def hash_encode(x, params, table, n_features):
    """Multiresolution hash-grid forward (tcnn kernel_grid, Linear interpolation).

    x: [N,3] (any float dtype; arithmetic runs in x.dtype except `pos`, which emulates the
    single-rounding fmaf(scale, x, 0.5f) of tcnn pos_fract via float64).  params: flat
    [(sum sizes) * F] laid out level-major, entry-major, feature-minor.
    """
    N = x.shape[0]
    dt = x.dtype
    outs = []
    p2 = params.view(-1, n_features)
    for lv in table:
        scale, res, size, off = lv['scale'], lv['res'], lv['size'], lv['offset']
        if dt == torch.float32:
            pos = (x.double() * scale + 0.5).to(dt)          # == fmaf in fp32 (exact product in fp64)
        else:
            pos = x * scale + 0.5
        g = torch.floor(pos)
        frac = pos - g
        # uint32 wrap-around of (uint32_t)(int)floor(pos): keep int64 and mask to 32 bits
        gi = g.detach().to(torch.int64) & 0xFFFFFFFF
        acc = torch.zeros(N, n_features, dtype=dt, device=x.device)
        for corner in range(8):
            w = torch.ones(N, dtype=dt, device=x.device)
            c = []
            for d in range(3):
                if (corner >> d) & 1:
                    w = w * frac[:, d]
                    c.append((gi[:, d] + 1) & 0xFFFFFFFF)
                else:
                    w = w * (1 - frac[:, d])
                    c.append(gi[:, d])
            # grid_index(): dense stride walk while stride <= size, hashed if the level overflows
            stride, idx, d = 1, torch.zeros(N, dtype=torch.int64, device=x.device), 0
            while d < 3 and stride <= size:
                idx = (idx + c[d] * stride) & 0xFFFFFFFF
                stride *= res
                d += 1
            if size < stride:
                idx = torch.zeros(N, dtype=torch.int64, device=x.device)
                for d in range(3):
                    idx = idx ^ ((c[d] * PRIMES[d]) & 0xFFFFFFFF)
            idx = idx % size
            acc = acc + w[:, None] * p2[off + idx]
        outs.append(acc)
    return torch.cat(outs, dim=1)


def _quartic_cdf(t, n_bins):
    u = t * n_bins
    u2 = u * u
    u4 = u2 * u2
    return torch.clamp((15.0 / 16.0) * u * (1 - (2.0 / 3.0) * u2 + (1.0 / 5.0) * u4) + 0.5, 0.0, 1.0)


def oneblob_encode(x, n_bins):
    """OneBlob forward (tcnn kernel_one_blob): per dim, bin b = C((b+1)/n - x) - C(b/n - x) with the
    wrapped boundary CDF C(t) = Q(t) + Q(t-1) + Q(t+1); the last bin's right boundary is C(0 - x) + 1."""
    N, D = x.shape
    b = torch.arange(n_bins, dtype=x.dtype, device=x.device) / n_bins           # left boundaries
    t = b[None, None, :] - x[:, :, None]                                        # [N,D,n_bins]
    left = _quartic_cdf(t, n_bins) + _quartic_cdf(t - 1.0, n_bins) + _quartic_cdf(t + 1.0, n_bins)
    right = torch.cat([left[..., 1:], left[..., :1] + 1.0], dim=-1)
    return (right - left).reshape(N, D * n_bins)


class Encoding(nn.Module):
    """tcnn.Encoding(n_input_dims, encoding_config, dtype) look-alike (bindings/torch/tinycudann/modules.py):
    one flat fp32 `.params`, `.n_output_dims`, __call__([N,3]) -> [N,C]."""

    def __init__(self, n_input_dims, encoding_config, dtype=torch.float, seed=1337):
        super().__init__()
        self.n_input_dims = n_input_dims
        self.encoding_config = dict(encoding_config)
        otype = encoding_config['otype']
        if otype in ('HashGrid', 'Grid'):
            assert n_input_dims == 3
            self.kind = 'hash'
            self.n_levels = int(encoding_config.get('n_levels', 16))
            self.n_features = int(encoding_config.get('n_features_per_level', 2))
            self.table, total = level_table(
                self.n_levels, int(encoding_config.get('base_resolution', 16)),
                float(encoding_config.get('per_level_scale', 2.0)),
                int(encoding_config.get('log2_hashmap_size', 19)))
            g = torch.Generator().manual_seed(seed)
            init = (torch.rand(total * self.n_features, generator=g) * 2 - 1) * 1e-4   # tcnn: U(-1e-4, 1e-4)
            self.params = nn.Parameter(init.to(torch.float32))
            self.n_output_dims = self.n_levels * self.n_features
        elif otype == 'OneBlob':
            self.kind = 'oneblob'
            self.n_bins = int(encoding_config.get('n_bins', 16))
            self.params = nn.Parameter(torch.zeros(0, dtype=torch.float32))
            self.n_output_dims = n_input_dims * self.n_bins
        else:
            raise NotImplementedError(f'oracle shim: encoding {otype} is not on the hot path')

    def forward(self, x):
        x = x.to(self.params.dtype if self.params.numel() else torch.float32).contiguous()
        if self.kind == 'hash':
            return hash_encode(x, self.params, self.table, self.n_features)
        return oneblob_encode(x, self.n_bins)


class Network(nn.Module):  # decoder.tcnn_network is False at every shipped config
    def __init__(self, *a, **k):
        raise NotImplementedError('oracle shim: tcnn.Network (FullyFusedMLP) is not used by NARUTO configs')
