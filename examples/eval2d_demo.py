#!/usr/bin/env python
"""Evaluation loop in the shape of the reference's eval2d.py (eval2d.py:44-115) on the GPU at batch rate:
binary RHD records -> on-device decode -> inference2d -> detect_keypoints -> trafo_coords -> EvalUtil.

The RHD evaluation set is not available offline; with no arguments the script fabricates records with the exact
on-disk layout (data/BinaryDbReader.py:103-208) from random images / key-points, so the numbers only demonstrate the
plumbing.  Pass --db path/to/rhd_evaluation.bin for the real evaluation.
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from nets.ColorHandPose3DNetwork import ColorHandPose3DNetwork
from utils.general import EvalUtil, detect_keypoints, trafo_coords
from hand3d_b200.data.records import RHD_RECORD_BYTES, decode_rhd_records


def fake_records(n, seed=0):
    rng = np.random.default_rng(seed)
    out = bytearray()
    for _ in range(n):
        out += rng.normal(size=(42, 3)).astype(np.float32).tobytes()
        out += rng.uniform(40, 280, size=(42, 2)).astype(np.float32).tobytes()
        out += np.eye(3, dtype=np.float32).tobytes() + b"\xff\xff"
        out += rng.integers(0, 256, size=(320, 320, 3), dtype=np.uint8).tobytes()
        out += rng.integers(0, 34, size=(320, 320), dtype=np.uint8).tobytes()
        out += rng.integers(0, 2, size=42, dtype=np.uint8).tobytes()
    return bytes(out)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument("--db", default=None)
    ap.add_argument("--weights", nargs="*", default=None)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--samples", type=int, default=64)
    args = ap.parse_args()

    net = ColorHandPose3DNetwork()
    if args.weights:
        net.init(None, weight_files=args.weights, exclude_var_list=['PosePrior', 'ViewpointNet'])      # eval2d.py:78-79
    else:
        from hand3d_b200.weights import synthetic_weights
        net.init(None, weights=synthetic_weights(0), exclude_var_list=['PosePrior', 'ViewpointNet'])

    blob = open(args.db, "rb").read() if args.db else fake_records(args.samples)
    n = len(blob) // RHD_RECORD_BYTES
    util = EvalUtil()
    for lo in range(0, n, args.batch):
        data = decode_rhd_records(blob[lo * RHD_RECORD_BYTES:(lo + args.batch) * RHD_RECORD_BYTES])
        image = data['image'][:, :240].contiguous()                       # evaluation frames are 240x320 (eval2d.py:53)
        keypoints_scoremap, image_crop, scale_crop, center = net.inference2d(image)              # eval2d.py:58
        coord_hw_crop = detect_keypoints(keypoints_scoremap)              # [B,21,2] (row, col) on device
        coord_hw = trafo_coords(coord_hw_crop, center, scale_crop, 256)   # eval2d.py:94
        coord_uv = torch.stack([coord_hw[:, :, 1], coord_hw[:, :, 0]], -1).to(torch.float32)      # eval2d.py:95
        # left hand = first 21 key-points of the record (the reader's dominant-hand selection is out of scope here)
        util.feed(data['keypoint_uv'][:, :21], data['keypoint_vis'][:, :21], coord_uv)            # eval2d.py:106
    mean, median, auc, _, _ = util.get_measures(0.0, 30.0, 20)            # eval2d.py:112
    print('Evaluation results (%d samples):' % n)
    print('Average mean EPE: %.3f pixels' % mean)
    print('Average median EPE: %.3f pixels' % median)
    print('Area under curve: %.3f' % auc)
