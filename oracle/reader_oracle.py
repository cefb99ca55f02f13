"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, one sample at a time) of the forward generators the reference's
dataset readers compute next to the raw record decode (evaluation mode, i.e. no augmentation noise):

* data/BinaryDbReader.py:139-162 palm-coordinate substitution, :210-250 dominant-hand selection / 21-key-point subsets /
  root-relative normalisation, :269-346 ground-truth hand crop (centre, size, scale, crop, key-points and intrinsics in crop
  space), :348-366 score-map targets, :368-381 scale_to_size, :413-459 create_multiple_gaussian_map;
* data/BinaryDbReaderSTB.py:123-196 (mm -> m, convert_kp :397-410, wrist extrapolation, root-relative normalisation);
* utils/canonical_trafo.py:20-136 (atan2, canonical_trafo) and :139-162 (flip_right_hand).

Pinned to the reference SOURCE by tests/golden/golden_reference_reader.npz: the unmodified reader classes executed over the eager
TF stand-in (oracle/tf1_eager.py) on seeded synthetic records (tests/golden/make_golden_reference_reader.py).  Only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import this package; the product never does.
"""
from __future__ import annotations

import numpy as np

from . import hand3d_oracle as O
from . import tf1_ops as T

f32 = np.float32
PI = f32(3.141592653589793)


# ------------------------------------------------------------------------------------------ score-map targets
def create_multiple_gaussian_map(coords_hw, output_size, sigma, valid_vec=None):
    """data/BinaryDbReader.py:413-459.  coords_hw [N,2] (row, col) float; -> [H,W,N] float32.
    The coordinates are truncated to int32 first (:421); a key-point contributes only if it is valid and strictly inside
    (0, size-1) in both axes (:430-433); map = exp(-((X-r)^2 + (Y-c)^2) / sigma^2)."""
    sigma = f32(sigma)
    c = np.asarray(coords_hw, f32).astype(np.int32)
    n = c.shape[0]
    val = np.ones(n, bool) if valid_vec is None else (np.asarray(valid_vec).astype(f32).reshape(-1) > f32(0.5))
    cond = val & (c[:, 0] < output_size[0] - 1) & (c[:, 0] > 0) & (c[:, 1] < output_size[1] - 1) & (c[:, 1] > 0)
    cf = c.astype(f32)
    X = np.arange(output_size[0], dtype=f32).reshape(-1, 1, 1) - cf[:, 0].reshape(1, 1, n)
    Y = np.arange(output_size[1], dtype=f32).reshape(1, -1, 1) - cf[:, 1].reshape(1, 1, n)
    dist = np.square(X) + np.square(Y)
    return (np.exp(-dist / np.square(sigma)) * cond.astype(f32)).astype(f32)


# ------------------------------------------------------------------------------------------ canonical frame
def atan2_ref(y, x):
    """utils/canonical_trafo.py:20-38 (the reference's own atan2 built from atan)."""
    y, x = np.asarray(y, f32), np.asarray(x, f32)
    tan = np.arctan(y / (x + f32(1e-8))).astype(f32)
    tan_c = tan + np.where(x + f32(1e-8) < 0, PI, f32(0)).astype(f32)
    t2 = tan_c + np.where(tan_c < 0, f32(2) * PI, f32(0)).astype(f32)
    return (t2 + np.where(t2 > PI, f32(-2) * PI, f32(0)).astype(f32)).astype(f32)


def _rot_x(a):
    c, s = np.cos(a).astype(f32), np.sin(a).astype(f32)
    return np.array([[1, 0, 0], [0, c, s], [0, -s, c]], f32)          # :67-74


def _rot_y(a):
    c, s = np.cos(a).astype(f32), np.sin(a).astype(f32)
    return np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]], f32)          # :77-84


def _rot_z(a):
    c, s = np.cos(a).astype(f32), np.sin(a).astype(f32)
    return np.array([[c, s, 0], [-s, c, 0], [0, 0, 1]], f32)          # :87-94


def canonical_trafo(coords_xyz):
    """utils/canonical_trafo.py:97-136 for ONE sample [21,3] -> (coords in the canonical frame [21,3], total rotation [3,3])."""
    x = np.asarray(coords_xyz, f32).reshape(21, 3)
    t = x - x[0:1]
    p = t[12]
    r1 = _rot_z(atan2_ref(p[0], p[1]))
    t1 = (t @ r1).astype(f32)
    total = r1
    p = t1[12]
    beta = -atan2_ref(p[2], p[1])
    r2 = _rot_x(f32(beta + PI))
    t2 = (t1 @ r2).astype(f32)
    total = (total @ r2).astype(f32)
    p = t2[20]
    r3 = _rot_y(atan2_ref(p[2], p[0]))
    return (t2 @ r3).astype(f32), (total @ r3).astype(f32)


def flip_right_hand(coords, cond_right):
    """utils/canonical_trafo.py:139-162: z -> -z where cond_right."""
    c = np.array(coords, f32, copy=True)
    if bool(np.asarray(cond_right).reshape(-1)[0]):
        c[..., 2] = -c[..., 2]
    return c


# ------------------------------------------------------------------------------------------ RHD reader
def rhd_items(record, use_wrist_coord=True, hand_crop=False, scale_to_size=False, sigma=25.0, crop_size=256):
    """BinaryDbReader(mode, shuffle=False, use_wrist_coord, hand_crop, scale_to_size, sigma).get() for ONE record, evaluation mode."""
    raw = O.decode_rhd_record(record)
    xyz, uv, vis = raw["keypoint_xyz"].astype(f32), raw["keypoint_uv"].astype(f32), raw["keypoint_vis"].astype(bool)
    if not use_wrist_coord:                                            # :139-162
        xyz = np.concatenate([(f32(0.5) * (xyz[0] + xyz[12]))[None], xyz[1:21], (f32(0.5) * (xyz[21] + xyz[33]))[None], xyz[-20:]], 0)
        uv = np.concatenate([(f32(0.5) * (uv[0] + uv[12]))[None], uv[1:21], (f32(0.5) * (uv[21] + uv[33]))[None], uv[-20:]], 0)
        vis = np.concatenate([[vis[0] | vis[12]], vis[1:21], [vis[21] | vis[33]], vis[-20:]], 0)
    d = {"keypoint_xyz": xyz, "keypoint_uv": uv, "cam_mat": raw["cam_mat"], "image": raw["image"], "hand_parts": raw["hand_parts"],
         "hand_mask": raw["hand_mask"], "keypoint_vis": vis}
    parts = raw["hand_parts"]
    n_left = int(((parts > 1) & (parts < 18)).sum())                   # :212-219
    n_right = int((parts > 17).sum())
    left = n_left > n_right
    xyz21 = xyz[:21] if left else xyz[-21:]
    d["hand_side"] = np.array([1.0, 0.0] if left else [0.0, 1.0], f32)
    d["keypoint_xyz21"] = xyz21
    rel = xyz21 - xyz21[0]
    scale_len = np.sqrt(np.sum(np.square(rel[12] - rel[11]))).astype(f32)
    d["keypoint_scale"] = scale_len
    d["keypoint_xyz21_normed"] = (rel / scale_len).astype(f32)
    d["keypoint_xyz21_local"] = O.bone_rel_trafo(d["keypoint_xyz21_normed"][None])[0]
    can, rot = canonical_trafo(d["keypoint_xyz21_normed"])
    d["keypoint_xyz21_can"] = flip_right_hand(can, not left)
    d["rot_mat"] = np.linalg.inv(rot).astype(f32)
    vis21 = vis[:21] if left else vis[-21:]
    uv21 = uv[:21] if left else uv[-21:]
    d["keypoint_vis21"], d["keypoint_uv21"] = vis21, uv21
    size = (320, 320)
    if hand_crop:                                                      # :269-346
        center = uv21[12, ::-1].astype(f32)
        if not np.all(np.isfinite(center)):
            center = np.zeros(2, f32)
        hw = np.stack([uv21[:, 1][vis21], uv21[:, 0][vis21]], 1)
        mn = np.maximum(hw.min(0) if hw.size else np.full(2, np.inf, f32), f32(0.0))
        mx = np.minimum(hw.max(0) if hw.size else np.full(2, -np.inf, f32), np.array(size, f32))
        best = (f32(2) * np.maximum(mx - center, center - mn)).max()
        best = np.minimum(np.maximum(best, f32(50.0)), f32(500.0))
        if not np.isfinite(best):
            best = f32(200.0)
        scale = f32(crop_size) / f32(best)
        scale = f32(np.minimum(np.maximum(scale, f32(1.0)), f32(10.0)))
        d["crop_scale"] = scale
        d["crop_center"] = center
        d["image_crop"] = O.crop_image_from_xy(raw["image"][None], center[None], crop_size, np.array([[scale]], f32))[0]
        u = (uv21[:, 0] - center[1]) * scale + f32(crop_size // 2)
        v = (uv21[:, 1] - center[0]) * scale + f32(crop_size // 2)
        uv21 = np.stack([u, v], 1).astype(f32)
        d["keypoint_uv21"] = uv21
        S = np.array([[scale, 0, 0], [0, scale, 0], [0, 0, 1]], f32)
        t1 = center[0] * scale - f32(crop_size // 2)
        t2 = center[1] * scale - f32(crop_size // 2)
        Tm = np.array([[1, 0, -t2], [0, 1, -t1], [0, 0, 1]], f32)
        d["cam_mat"] = (Tm @ (S @ raw["cam_mat"]).astype(f32)).astype(f32)
        size = (crop_size, crop_size)
    hw21 = np.stack([uv21[:, 1], uv21[:, 0]], -1)
    d["scoremap"] = create_multiple_gaussian_map(hw21, size, sigma, vis21)
    if scale_to_size:                                                  # :368-381 (eval2d.py:43)
        img = T.resize_bilinear_tf1(raw["image"][None], 240, 320)[0]
        sc = (240 / float(320), 320 / float(320))
        d = {"image": img, "keypoint_uv21": np.stack([d["keypoint_uv21"][:, 0] * f32(sc[1]), d["keypoint_uv21"][:, 1] * f32(sc[0])], 1).astype(f32),
             "keypoint_vis21": d["keypoint_vis21"]}
    return d


# ------------------------------------------------------------------------------------------ STB reader
STB_ORDER = [0, 20, 19, 18, 17, 16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1]    # data/BinaryDbReaderSTB.py:397-410


def stb_items(record, use_wrist_coord=True, sigma=25.0, with_scoremap=True):
    """BinaryDbReaderSTB(mode, shuffle=False, use_wrist_coord, sigma).get() for ONE record (evaluation mode, no hand crop)."""
    raw = np.frombuffer(bytes(record), dtype=np.uint8)
    assert raw.size == 922104, "Doesnt add up."
    f = np.frombuffer(bytes(record[:504]), dtype=f32)
    xyz = (f[:63].reshape(21, 3) / f32(1000.0))[STB_ORDER]
    uvv = f[63:126].reshape(21, 3)[STB_ORDER]
    uv, vis = uvv[:, :2].astype(f32), uvv[:, 2] == f32(1.0)
    if use_wrist_coord:                                                # :131-134,147-154
        xyz = np.concatenate([(xyz[16] + f32(2.0) * (xyz[0] - xyz[16]))[None], xyz[1:]], 0)
        vis = np.concatenate([[vis[16] | vis[0]], vis[1:]], 0)
        uv = np.concatenate([(uv[16] + f32(2.0) * (uv[0] - uv[16]))[None], uv[1:]], 0)
    img = raw[504:].reshape(480, 640, 3).astype(f32) / f32(255.0) - f32(0.5)
    rel = xyz - xyz[0]
    scale_len = np.sqrt(np.sum(np.square(rel[12] - rel[11]))).astype(f32)
    d = {"keypoint_xyz21": xyz.astype(f32), "keypoint_vis21": vis, "keypoint_uv21": uv, "image": img,
         "cam_mat": np.array([[822.79041, 0.0, 318.47345], [0.0, 822.79041, 250.31296], [0.0, 0.0, 1.0]], f32),
         "hand_side": np.array([1.0, 0.0], f32), "keypoint_scale": scale_len, "keypoint_xyz21_normed": (rel / scale_len).astype(f32)}
    d["keypoint_xyz21_local"] = O.bone_rel_trafo(d["keypoint_xyz21_normed"][None])[0]       # :199-202
    can, rot = canonical_trafo(d["keypoint_xyz21_normed"])                                   # :204-208 (left hands only: no flip)
    d["keypoint_xyz21_can"] = can
    d["rot_mat"] = np.linalg.inv(rot).astype(f32)
    if with_scoremap:                                                                        # :296-313 on the full 480 x 640 image
        d["scoremap"] = create_multiple_gaussian_map(np.stack([uv[:, 1], uv[:, 0]], -1), (480, 640), sigma, vis)
    return d
