"""utils/canonical_trafo.py mirror: canonical_trafo / flip_right_hand on device (h3d_canonical_trafo, h3d_flip_right_hand)."""
from __future__ import annotations

import torch

from .. import runtime


def canonical_trafo(coords_xyz):
    """utils/canonical_trafo.py:97-136: [B,21,3] (or [21,3]) -> (coords in the canonical frame [B,21,3], total rotation [B,3,3])."""
    can, rot, _ = runtime.default_context().canonical_trafo(coords_xyz.reshape(-1, 21, 3).to(torch.float32))
    return can, rot


def flip_right_hand(coords_xyz_canonical, cond_right):
    """utils/canonical_trafo.py:139-162."""
    c = coords_xyz_canonical.reshape(-1, 21, 3)
    cond = torch.as_tensor(cond_right, device=c.device).reshape(-1)
    if cond.numel() == 1 and c.shape[0] > 1:
        cond = cond.expand(c.shape[0])
    return runtime.default_context().flip_right_hand(c.contiguous(), cond)
