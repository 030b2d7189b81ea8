"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the reference's OWN Python.

Run in the build container only:   python -m oracle.make_golden
It imports /root/reference (read-only) through oracle/ref_harness.py (tinycudann -> oracle/tcnn_shim.py),
drives src/slam/coslam/model/scene_rep.py:JointEncodingNaruto.{query_sdf,query_color_sdf,render_rays,forward}
on seeded synthetic inputs, and stores inputs + outputs (+ gradients).  Parameters are NOT stored (6.5 MB):
they are regenerated from the seed by oracle.naruto_oracle.init_params and guarded by a checksum.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import naruto_oracle as no            # noqa: E402
from oracle import ref_harness as rh              # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
GRAD_PROBE = 4096


def load_params_into(model, P):
    with torch.no_grad():
        model.embed_fn.params.copy_(P.grid)
        model.decoder.sdf_net.model[0].weight.copy_(P.w1)
        model.decoder.sdf_net.model[2].weight.copy_(P.w2)
        model.decoder.color_net.model[0].weight.copy_(P.w3)
        model.decoder.color_net.model[2].weight.copy_(P.w4)
        model.uncert_grid.copy_(P.uncert_grid)


def synth_rays(spec, B, seed, invalid_every=9):
    """Camera at a jittered bbox centre looking in random directions; depth = distance (along the
    un-normalised ray parameter) to the bbox walls shrunk by 10%, so surfaces lie inside the bound."""
    g = torch.Generator().manual_seed(seed)
    b = torch.tensor(spec.bound)
    centre = b.mean(1) + (torch.rand(3, generator=g) - 0.5) * 0.5
    d = torch.randn(B, 3, generator=g)
    d = d / d.norm(dim=1, keepdim=True) * (1 + 0.5 * torch.rand(B, 1, generator=g))
    o = centre[None].repeat(B, 1)
    lo, hi = b[:, 0] * 0.9, b[:, 1] * 0.9
    t = torch.where(d > 0, (hi - o) / d, (lo - o) / d).min(dim=1).values
    depth = t[:, None].clone()
    depth[::invalid_every] = 0.0                      # invalid-depth rays (target_d <= 0 branch)
    rgb = torch.rand(B, 3, generator=g)
    return o, d, rgb, depth


def grad_probe_idx(n, seed=123):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, n, (GRAD_PROBE,), generator=g)


def summarise_grid_grad(gg):
    idx = grad_probe_idx(gg.numel())
    return dict(grid_grad_probe=gg[idx].numpy(), grid_grad_sum=np.float64(gg.double().sum().item()),
                grid_grad_l2=np.float64(gg.double().norm().item()),
                grid_grad_nnz=np.int64((gg != 0).sum().item()))


def main():
    assert rh.reference_available(), 'needs /root/reference'
    os.makedirs(OUT, exist_ok=True)
    cfg = rh.load_reference_config('configs/Replica/office0/coslam.yaml')
    spec = no.spec_from_config(cfg)
    model = rh.build_reference_model(cfg)

    # anchors derivable from the reference text
    assert model.resolution_sdf == 275 and list(model.uncert_grid.shape) == [49, 56, 35]
    assert model.embed_fn.params.numel() == 1628176

    for tag, grid_range, jitter, pseed in (('small', 1e-4, 0.0, 11), ('wide', 0.5, 1.0, 12)):
        P = no.init_params(spec, seed=pseed, grid_range=grid_range, uncert_jitter=jitter)
        load_params_into(model, P)
        meta = dict(param_seed=pseed, grid_range=grid_range, uncert_jitter=jitter,
                    grid_checksum=np.float64(P.grid.double().sum().item()),
                    w1_checksum=np.float64(P.w1.double().sum().item()),
                    per_level_scale=np.float64(spec.per_level_scale))

        # ---- point queries (a8-a12, a19) -------------------------------------------------------
        g = torch.Generator().manual_seed(100 + pseed)
        x = torch.rand(509, 3, generator=g) * 1.2 - 0.1        # includes points outside [0,1]
        x[0] = 0.0
        x[1] = 1.0
        x[2] = torch.tensor([0.5, 0.25, 0.75])
        model.eval()
        with torch.no_grad():
            enc_hash = model.embed_fn(x)
            enc_blob = model.embedpos_fn(x)
            enc_full = model.calc_embedding(x)
            raw = model.query_color_sdf(x)
            sdf_u, geo = model.query_sdf(x[None], return_geo=True, return_uncert=True)
            sdf_only = model.query_sdf(x.reshape(1, 509, 3))
            emb = model.query_sdf(x.reshape(1, 509, 3), embed=True)
            col = model.query_color(x)
        np.savez_compressed(os.path.join(OUT, f'points_{tag}.npz'), x=x.numpy(), hash=enc_hash.numpy(),
                            oneblob=enc_blob.numpy(), uncert=enc_full[:, 0].numpy(), raw=raw.numpy(),
                            sdf_uncert=sdf_u[0].numpy(), geo=geo[0].numpy(), sdf=sdf_only[0].numpy(),
                            embed=emb[0].numpy(), color=col.numpy(), **meta)

        # ---- render_rays, eval mode (a6-a14) ---------------------------------------------------
        B = 160
        o, d, rgb, depth = synth_rays(spec, B, seed=200 + pseed)
        for perturb in (0, 1):
            model.config['training']['perturb'] = perturb
            torch.manual_seed(300 + pseed)
            with torch.no_grad():
                r = model.render_rays(o, d, depth)
            torch.manual_seed(300 + pseed)
            u = torch.rand(B, spec.n_samples) if perturb else torch.zeros(0)
            np.savez_compressed(os.path.join(OUT, f'render_{tag}_p{perturb}.npz'), rays_o=o.numpy(),
                                rays_d=d.numpy(), target_d=depth.numpy(), u=u.numpy(),
                                **{k: v.numpy() for k, v in r.items()}, **meta)
        model.config['training']['perturb'] = 1

        # ---- forward, train mode + gradients (a15) --------------------------------------------
        model.train()
        model.zero_grad()
        torch.manual_seed(400 + pseed)
        ret = model.forward(o, d, rgb, depth)
        torch.manual_seed(400 + pseed)
        u = torch.rand(B, spec.n_samples)
        names = ['rgb_loss', 'depth_loss', 'sdf_loss', 'fs_loss', 'uncert_loss']
        wts = [spec.rgb_weight, spec.depth_weight, spec.sdf_weight, spec.fs_weight, spec.uncert_weight]
        loss = sum(w * ret[n] for w, n in zip(wts, names))
        loss.backward()
        out = {k: v.detach().numpy() for k, v in ret.items()}
        out.update(summarise_grid_grad(model.embed_fn.params.grad))
        out.update(w1_grad=model.decoder.sdf_net.model[0].weight.grad.numpy(),
                   w2_grad=model.decoder.sdf_net.model[2].weight.grad.numpy(),
                   w3_grad=model.decoder.color_net.model[0].weight.grad.numpy(),
                   w4_grad=model.decoder.color_net.model[2].weight.grad.numpy(),
                   uncert_grid_grad=model.uncert_grid.grad.numpy(), loss=loss.detach().numpy())
        np.savez_compressed(os.path.join(OUT, f'train_{tag}.npz'), rays_o=o.numpy(), rays_d=d.numpy(),
                            target_rgb=rgb.numpy(), target_d=depth.numpy(), u=u.numpy(), **out, **meta)
        print(tag, {n: float(ret[n]) for n in names}, 'loss', float(loss))

    # state_dict key contract
    with open(os.path.join(OUT, 'state_dict_keys.txt'), 'w') as f:
        for k, v in model.state_dict().items():
            f.write(f'{k} {list(v.shape)}\n')


if __name__ == '__main__':
    main()
