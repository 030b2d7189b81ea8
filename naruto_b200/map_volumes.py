"""Dense uncertainty / SDF sweep for the planner: the reference's `get_map_volumes(query_fn, bounding_box, voxel_size)`
(src/slam/coslam/coslam_utils.py:58-97, called at src/slam/coslam/coslam.py:583,614) as ONE kernel launch -- lattice
generation, encoding, SDF net and the on-surface mask fused (`nrt_map_volumes`), instead of two `query_sdf` passes (one of
them discarded) plus a dozen elementwise ops.  Pass the scene model instead of its `query_sdf` bound method."""
import numpy as np
import torch

from . import _lib as L


def get_map_volumes(model, bounding_box=None, voxel_size=0.1, to_numpy=True):
    """-> [uncert_vol, sdf_vol], numpy arrays like the reference (to_numpy=True) or device tensors (hand-over to a
    device-side planner without the GPU -> numpy -> GPU round trip of src/planner/naruto_planner.py:633-634)."""
    if bounding_box is not None:
        bb = torch.as_tensor(bounding_box, dtype=torch.float32).cpu()
        if not torch.equal(bb, model.plan.bound):
            raise L.NrtError('get_map_volumes: bounding_box differs from the bound the scene model was built with')
    with torch.no_grad():
        unc, sdf = model.plan.map_volumes(model._tensors(), float(voxel_size))
    if to_numpy:
        return [unc.cpu().numpy().copy(), sdf.cpu().numpy().copy()]
    return [unc, sdf]
