"""TEST INFRASTRUCTURE -- CPU restatement of ERPDepth2Dist.forward (src/layers/erp_conversions.py:288-354) in torch ops.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this.  Pinned to the reference's own
ERPDepth2Dist by tests/golden/erp_small.npz (oracle/make_golden_erp.py imports the class from /root/reference)."""
import torch
import torch.nn.functional as F


def erp_depth2dist(erp_depth, c2e_grid, face_coor, face_rays, skybox_size):
    """erp_depth [H,W]; c2e_grid [H,W,3]; face_coor [6,s,s,2]; face_rays [3,s*s] -> erp_dist [H,W] (fp32, CPU)."""
    s = int(skybox_size)
    H, W = erp_depth.shape
    img = erp_depth.reshape(1, 1, H, W).float()
    faces = []
    for f in range(6):
        # E2P.forward (src/layers/erp_conversions.py:70-82)
        pers = F.grid_sample(img, face_coor[f].reshape(1, s, s, 2), align_corners=True, padding_mode='border', mode='bilinear')
        # depth2dist (src/layers/erp_conversions.py:268-286): points = depth * (K^-1 pix); norm over xyz
        pts = pers.reshape(1, 1, -1) * face_rays.reshape(1, 3, -1)
        faces.append(torch.norm(pts, dim=1).reshape(s, s))
    cube = torch.stack(faces).reshape(1, 1, 6, s, s)
    # C2E.forward (src/layers/c2e.py:132-137): 3-D nearest, zeros padding
    pano = F.grid_sample(cube, c2e_grid.reshape(1, 1, H, W, 3), align_corners=True, mode='nearest')
    return pano[0, 0, 0]
