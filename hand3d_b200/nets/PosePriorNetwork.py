"""PosePriorNetwork -- lifting 2D score maps to 3D behind the reference's API
(nets/PosePriorNetwork.py:29-159).  All five variants ('direct', 'bottleneck', 'local',
'local_w_xyz_loss', 'proposed') run on the same sm_100a kernels as ColorHandPose3DNetwork; the 'local*' variants add
the forward-kinematics kernel that replaces bone_rel_trafo_inv (utils/relative_trafo.py:243-295).
"""
from __future__ import annotations

import os

import torch

from .. import runtime, weights as _weights


class PosePriorNetwork(object):
    """ Network containing different variants for lifting 2D predictions into 3D. """
    def __init__(self, variant):
        self.num_kp = 21
        self.variant = variant

    def init(self, session=None, weight_files=None, exclude_var_list=None, weights=None):
        """ Initializes weights from pickled python dictionaries (reference :36-57). """
        if exclude_var_list is None:
            exclude_var_list = list()
        ctx = runtime.default_context()
        if weights is not None:
            wd = {k: v for k, v in weights.items() if not any([x in k for x in exclude_var_list])}
            ctx.load_weights(wd)
            print('Loaded %d variables from %s' % (len(wd), 'memory'))
            return
        for file_name in weight_files:
            assert os.path.exists(file_name), "File not found."
            wd = _weights.load_weight_files([file_name], exclude_var_list, verbose=False)
            if len(wd) > 0:
                ctx.load_weights(wd)
                print('Loaded %d variables from %s' % (len(wd), file_name))

    def inference(self, scoremap, hand_side, evaluation=True):
        """ Infere 3D coordinates from 2D scoremaps (reference :59-95).

            scoremap [B,256,256,21] -> avg_pool 8x8 -> variant.  Returns (coord_xyz_rel_normed, coord3d, R).
        """
        ev = bool(evaluation.item()) if torch.is_tensor(evaluation) else bool(evaluation)
        if not ev:
            raise NotImplementedError("forward pass only: evaluation must be True")
        ctx = runtime.default_context()
        scoremap_pooled = ctx.avg_pool8(scoremap)                       # :61
        if self.variant in ('direct', 'bottleneck'):
            c, _, _ = ctx.lifting(scoremap_pooled, hand_side, self.variant)
            return c, c, None
        elif self.variant in ('local', 'local_w_xyz_loss'):
            # :70-75 -- the net predicts bone-relative coords; bone_rel_trafo_inv (utils/relative_trafo.py:243) assembles xyz
            normed, rel, _ = ctx.lifting(scoremap_pooled, hand_side, 'local')
            return normed, rel, None
        elif self.variant == 'proposed':
            out, can, R = ctx.lifting(scoremap_pooled, hand_side, 'proposed')
            return out, can, R
        else:
            assert 0, "Unknown variant."
