"""Generates tests/golden/golden_small.npz from the oracle (the reference itself cannot run here: no TensorFlow 1.3,
no released weights -- see DESIGN.md section 2, "parity unpinned").  The fixtures freeze the oracle's outputs on small
seeded inputs so that (a) drift of the oracle across machines / library versions is detected on CPU and (b) the CUDA
path is checked against committed numbers, not only against an oracle evaluated in the same process.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from hand3d_b200 import weights as Wt  # noqa: E402
from oracle import hand3d_oracle as O  # noqa: E402
from oracle import tf1_ops as T  # noqa: E402


def golden_inputs():
    """Seeded inputs (regenerated, not stored): everything numpy's PCG64 / the synthetic generators give bit-identically."""
    rng = np.random.default_rng(43)
    sm = rng.normal(size=(2, 32, 32, 21)).astype(np.float32)
    low = (rng.normal(size=(2, 40, 40, 2)) * 3).astype(np.float32)
    low[..., 1] -= 1.5
    return {
        "seg_image": Wt.synthetic_images(1, 64, 64, seed=41),
        "pose_crop": Wt.synthetic_images(1, 64, 64, seed=42),
        "lift_scoremap": sm,
        "lift_hand_side": np.array([[1, 0], [0, 1]], np.float32),
        "pp_low": low,
        "crop_image": Wt.synthetic_images(2, 320, 320, seed=44),
    }


def build():
    wd = Wt.synthetic_weights(0)
    i = golden_inputs()
    g = {}
    g["seg_logits"] = O.inference_detection(i["seg_image"], wd)[-1]            # [1,64,64,2]
    s = O.inference_pose2d(i["pose_crop"], wd)
    g["pose_s0"], g["pose_s2"] = s[0], s[2]                                   # [1,8,8,21]
    out, can, R = O.inference_pose3d(i["lift_scoremap"], i["lift_hand_side"], wd)
    g["lift_coord3d"], g["lift_can"], g["lift_R"] = out, can, R
    # post-processing on a synthetic smooth logit field (exact integer / bit-exact outputs)
    logits = T.resize_bilinear_tf1(i["pp_low"], 320, 320)
    mask = O.single_obj_scoremap(logits, literal=True)
    center, _, size = O.calc_center_bb(mask)
    g["pp_mask_rows"] = mask[..., 0].sum(axis=2).astype(np.int32)              # per-row / per-column pixel counts (compact)
    g["pp_mask_cols"] = mask[..., 0].sum(axis=1).astype(np.int32)
    g["pp_center"], g["pp_size"], g["pp_scale"] = center, size, O.crop_scale(size)
    fg, _ = O.seg_fg_det(logits)
    g["pp_max_loc"] = O.find_max_location(fg)
    cropped = O.crop_image_from_xy(i["crop_image"], center, 256, g["pp_scale"])
    g["crop_checksum"] = np.array([cropped.astype(np.float64).sum(), np.abs(cropped).astype(np.float64).sum()])
    g["crop_samples"] = cropped[:, ::37, ::41, :]
    kp = T.resize_bilinear_tf1(s[2], 64, 64)
    g["kp_uv"] = O.detect_keypoints(kp[0]).astype(np.int32)
    return g


if __name__ == "__main__":
    g = build()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_small.npz")
    np.savez_compressed(path, **g)
    print(path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in g.items()})
