"""BinaryDbReader / BinaryDbReaderSTB -- B200-native mirrors of the reference's dataset readers for the EVALUATION drivers
(data/BinaryDbReader.py:21-412, data/BinaryDbReaderSTB.py:21-330): same constructor arguments, `num_samples`, and `get()`
returning the same dictionary keys, but eager: every get() call uploads the next `batch_size` fixed-length records as bytes and
produces the raw and derived items on the GPU (h3d_decode_records, h3d_rhd_reader_items / h3d_stb_reader_items,
h3d_crop_image_from_xy, h3d_gaussian_scoremap, h3d_canonical_trafo).  Training-time augmentation (hue, coordinate / crop noise,
score-map dropout, random crops, shuffling) is out of scope and refused.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .. import runtime
from .records import RHD_RECORD_BYTES, STB_RECORD_BYTES

_AUG = ("random_crop_to_size", "hue_aug", "coord_uv_noise", "crop_center_noise", "crop_scale_noise", "crop_offset_noise", "scoremap_dropout")


class _RecordFile:
    def __init__(self, path, record_bytes, num_samples):
        assert os.path.exists(path), "Could not find the binary data file!"
        self.mm = np.memmap(path, dtype=np.uint8, mode="r")
        self.record_bytes = record_bytes
        self.available = self.mm.size // record_bytes
        self.num_samples = min(num_samples, self.available) if self.available else num_samples
        self.pos = 0

    def next_batch(self, n):
        idx = [(self.pos + i) % max(1, self.available) for i in range(n)]       # the TF queue cycles through the file
        self.pos += n
        rec = np.stack([np.asarray(self.mm[i * self.record_bytes:(i + 1) * self.record_bytes]) for i in idx])
        dev = runtime.default_context().device
        return torch.from_numpy(rec).pin_memory().to(dev, non_blocking=True)


class BinaryDbReader(object):
    """ Reads data from a binary dataset created by create_binary_db.py (RHD). """
    def __init__(self, mode=None, batch_size=1, shuffle=True, use_wrist_coord=True, sigma=25.0, hand_crop=False, random_crop_to_size=False,
                 scale_to_size=False, hue_aug=False, coord_uv_noise=False, crop_center_noise=False, crop_scale_noise=False,
                 crop_offset_noise=False, scoremap_dropout=False, path_to_db=None):
        if mode == 'training':
            path, n = './data/bin/rhd_training.bin', 41258
        elif mode == 'evaluation':
            path, n = './data/bin/rhd_evaluation.bin', 2728
        else:
            assert 0, "Unknown dataset mode."
        for name in _AUG:
            if locals()[name]:
                raise NotImplementedError("hand3d_b200 readers serve the evaluation drivers: %s is training-time augmentation" % name)
        if shuffle:
            raise NotImplementedError("shuffle=True (training) is out of scope; the evaluation drivers pass shuffle=False")
        self._file = _RecordFile(path_to_db or path, RHD_RECORD_BYTES, n)
        self.path_to_db = path_to_db or path
        self.num_samples = self._file.num_samples
        self.batch_size, self.sigma, self.shuffle, self.use_wrist_coord = batch_size, sigma, shuffle, use_wrist_coord
        self.scale_to_size, self.scale_target_size, self.hand_crop = scale_to_size, (240, 320), hand_crop
        self.image_size, self.crop_size, self.num_kp = (320, 320), 256, 42

    def get(self):
        """ Next batch as a dict of CUDA tensors with the reference's keys (data/BinaryDbReader.py:100-411). """
        ctx = runtime.default_context()
        B = self.batch_size
        raw = ctx.decode_records(self._file.next_batch(B), "rhd", 1)
        h = raw["header"]
        it = ctx.rhd_reader_items(h, raw["mask"], raw["visibility"], self.use_wrist_coord, self.hand_crop, self.crop_size)
        xyz = h[:, :126].reshape(B, 42, 3)
        uv = h[:, 126:210].reshape(B, 42, 2).to(torch.int32).to(torch.float32)
        vis = raw["visibility"].to(torch.bool)
        if not self.use_wrist_coord:       # the 42-key-point views with the palm substituted (:139-162,195-200)
            xyz = torch.cat([0.5 * (xyz[:, 0:1] + xyz[:, 12:13]), xyz[:, 1:21], 0.5 * (xyz[:, 21:22] + xyz[:, 33:34]), xyz[:, 22:]], 1)
            uv = torch.cat([0.5 * (uv[:, 0:1] + uv[:, 12:13]), uv[:, 1:21], 0.5 * (uv[:, 21:22] + uv[:, 33:34]), uv[:, 22:]], 1)
            vis = torch.cat([vis[:, 0:1] | vis[:, 12:13], vis[:, 1:21], vis[:, 21:22] | vis[:, 33:34], vis[:, 22:]], 1)
        parts = raw["mask"].to(torch.int32)
        hand = parts > 1
        d = {"keypoint_xyz": xyz, "keypoint_uv": uv, "cam_mat": it["cam_mat"], "image": raw["image"], "hand_parts": parts,
             "hand_mask": torch.stack([~hand, hand], 3).to(torch.int32), "keypoint_vis": vis, "hand_side": it["hand_side"],
             "keypoint_xyz21": it["keypoint_xyz21"], "keypoint_scale": it["keypoint_scale"], "keypoint_xyz21_normed": it["keypoint_xyz21_normed"],
             "keypoint_vis21": it["keypoint_vis21"].to(torch.bool), "keypoint_uv21": it["keypoint_uv21"]}
        right = it["hand_side"][:, 1] > 0.5
        can, _, rot_inv = ctx.canonical_trafo(it["keypoint_xyz21_normed"], right)
        d["keypoint_xyz21_can"], d["rot_mat"] = can, rot_inv
        size = self.image_size
        if self.hand_crop:
            d["crop_scale"] = it["crop_scale"]
            d["image_crop"] = ctx.crop_image_from_xy(raw["image"], it["crop_center"], self.crop_size, it["crop_scale"])
            size = (self.crop_size, self.crop_size)
        hw21 = torch.stack([it["keypoint_uv21"][..., 1], it["keypoint_uv21"][..., 0]], -1).contiguous()
        d["scoremap"] = ctx.gaussian_scoremap(hw21, size, self.sigma, it["keypoint_vis21"])
        if self.scale_to_size:             # :368-381: everything else is dropped
            s = self.image_size
            image = ctx.resize_bilinear(raw["image"], *self.scale_target_size)
            sc = (self.scale_target_size[0] / float(s[0]), self.scale_target_size[1] / float(s[1]))
            uv21 = torch.stack([d["keypoint_uv21"][..., 0] * sc[1], d["keypoint_uv21"][..., 1] * sc[0]], -1)
            d = {"image": image, "keypoint_uv21": uv21, "keypoint_vis21": d["keypoint_vis21"]}
        return d


class BinaryDbReaderSTB(object):
    """ Reads data from the STB binary dataset (data/BinaryDbReaderSTB.py). """
    def __init__(self, mode=None, batch_size=1, shuffle=True, use_wrist_coord=True, sigma=25.0, hand_crop=False, random_crop_to_size=False,
                 hue_aug=False, coord_uv_noise=False, crop_center_noise=False, crop_scale_noise=False, crop_offset_noise=False,
                 scoremap_dropout=False, path_to_db=None, with_scoremap=False):
        if mode == 'training':
            path, n = './data/stb/stb_train_shuffled.bin', 30000
        elif mode == 'evaluation':
            path, n = './data/stb/stb_eval.bin', 6000
        else:
            assert 0, "Unknown dataset mode."
        for name in _AUG:
            if locals().get(name):
                raise NotImplementedError("hand3d_b200 readers serve the evaluation drivers: %s is training-time augmentation" % name)
        if shuffle or hand_crop:
            raise NotImplementedError("shuffle / hand_crop on STB are training-time options; the evaluation driver (eval_full.py:45) uses neither")
        self._file = _RecordFile(path_to_db or path, STB_RECORD_BYTES, n)
        self.path_to_db = path_to_db or path
        self.num_samples = self._file.num_samples
        self.batch_size, self.sigma, self.use_wrist_coord, self.with_scoremap = batch_size, sigma, use_wrist_coord, with_scoremap
        self.image_size, self.crop_size, self.num_kp = (480, 640), 256, 21

    def get(self):
        ctx = runtime.default_context()
        B = self.batch_size
        raw = ctx.decode_records(self._file.next_batch(B), "stb", 1, want_aux=True)
        it = ctx.stb_reader_items(raw["header"], self.use_wrist_coord)
        dev = raw["image"].device
        d = {"keypoint_xyz21": it["keypoint_xyz21"], "keypoint_vis21": it["keypoint_vis21"].to(torch.bool), "keypoint_uv21": it["keypoint_uv21"],
             "image": raw["image"], "keypoint_scale": it["keypoint_scale"], "keypoint_xyz21_normed": it["keypoint_xyz21_normed"],
             "cam_mat": torch.tensor([[822.79041, 0.0, 318.47345], [0.0, 822.79041, 250.31296], [0.0, 0.0, 1.0]], device=dev).expand(B, 3, 3),
             "hand_side": torch.tensor([1.0, 0.0], device=dev).expand(B, 2).contiguous()}
        can, _, rot_inv = ctx.canonical_trafo(it["keypoint_xyz21_normed"], None)
        d["keypoint_xyz21_can"], d["rot_mat"] = can, rot_inv
        if self.with_scoremap:             # 480 x 640 x 21 targets (26 MB per sample): training only, off by default
            hw21 = torch.stack([it["keypoint_uv21"][..., 1], it["keypoint_uv21"][..., 0]], -1).contiguous()
            d["scoremap"] = ctx.gaussian_scoremap(hw21, self.image_size, self.sigma, it["keypoint_vis21"])
        return d
