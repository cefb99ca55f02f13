"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference forward pass
(lmb-freiburg/hand3d @ 9f00633).  PARITY UNPINNED (see oracle/tf1_ops.py header).

Every function cites the reference file:line it follows.  Inputs / outputs are numpy float32
NHWC arrays; ``weights`` is the reference's pickled ``{variable_name: ndarray}`` dictionary
(HWIO conv kernels, [in,out] FC matrices).  ``dtype=np.float64`` runs
the conv / FC arithmetic in double precision (the "fp64 twin" used to bound fp32 error).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product path never does.
"""
from __future__ import annotations

import numpy as np

from . import tf1_ops as T

CROP_SIZE = 256  # nets/ColorHandPose3DNetwork.py:31
NUM_KP = 21      # nets/ColorHandPose3DNetwork.py:32


def _conv(x, weights, scope, name, stride=1, relu=True, dtype=np.float32):
    w = weights[f"{scope}/{name}/weights"]
    b = weights[f"{scope}/{name}/biases"]
    y = T.conv2d_same(x, w, b, stride, dtype)
    return T.leaky_relu(y) if relu else y


def _fc(x, weights, scope, name, relu, dtype=np.float32):
    y = T.fully_connected(x, weights[f"{scope}/{name}/weights"], weights[f"{scope}/{name}/biases"], dtype)
    return np.maximum(y, y.dtype.type(T.NEG_SLOPE) * y) if relu else y  # utils/general.py:132-136


# --------------------------------------------------------------------------------------------
# HandSegNet  (nets/ColorHandPose3DNetwork.py:131-168)
# --------------------------------------------------------------------------------------------
def inference_detection(image, weights, dtype=np.float32, upsample=True):
    sc = "HandSegNet"
    x = image.astype(dtype)
    for block_id, (n, pool) in enumerate(zip([2, 2, 4, 4], [True, True, True, False]), 1):  # :144-154
        for layer_id in range(n):
            x = _conv(x, weights, sc, "conv%d_%d" % (block_id, layer_id + 1), dtype=dtype)
        if pool:
            x = T.max_pool_2x2(x)
    x = _conv(x, weights, sc, "conv5_1", dtype=dtype)                    # :156
    enc = _conv(x, weights, sc, "conv5_2", dtype=dtype)                  # :157
    x = _conv(enc, weights, sc, "conv6_1", dtype=dtype)                  # :160
    scoremap = _conv(x, weights, sc, "conv6_2", relu=False, dtype=dtype)  # :161
    if not upsample:
        return [scoremap]
    H, W = image.shape[1:3]
    return [T.resize_bilinear_tf1(scoremap, H, W)]                         # :165-166


# --------------------------------------------------------------------------------------------
# PoseNet2D  (nets/ColorHandPose3DNetwork.py:170-219)
# --------------------------------------------------------------------------------------------
def inference_pose2d(image_crop, weights, dtype=np.float32):
    sc = "PoseNet2D"
    x = image_crop.astype(dtype)
    for block_id, (n, pool) in enumerate(zip([2, 2, 4, 2], [True, True, True, False]), 1):  # :183-193
        for layer_id in range(n):
            x = _conv(x, weights, sc, "conv%d_%d" % (block_id, layer_id + 1), dtype=dtype)
        if pool:
            x = T.max_pool_2x2(x)
    for name in ("conv4_3", "conv4_4", "conv4_5", "conv4_6"):             # :195-198
        x = _conv(x, weights, sc, name, dtype=dtype)
    enc = _conv(x, weights, sc, "conv4_7", dtype=dtype)                    # :199
    x = _conv(enc, weights, sc, "conv5_1", dtype=dtype)                    # :202
    scoremaps = [_conv(x, weights, sc, "conv5_2", relu=False, dtype=dtype)]  # :203
    for pass_id in range(2):                                               # :209-215
        x = np.concatenate([scoremaps[-1], enc], axis=3)                   # :210
        for rec_id in range(5):
            x = _conv(x, weights, sc, "conv%d_%d" % (pass_id + 6, rec_id + 1), dtype=dtype)
        x = _conv(x, weights, sc, "conv%d_6" % (pass_id + 6), dtype=dtype)
        scoremaps.append(_conv(x, weights, sc, "conv%d_7" % (pass_id + 6), relu=False, dtype=dtype))
    return scoremaps


# --------------------------------------------------------------------------------------------
# Mask post-processing (utils/general.py:199-328)
# --------------------------------------------------------------------------------------------
def find_max_location(scoremap):
    """utils/general.py:199-230 -- first-occurrence arg-max per image -> (row, col) int32."""
    if scoremap.ndim == 4:
        scoremap = scoremap[..., 0]
    if scoremap.ndim == 2:
        scoremap = scoremap[None]
    B, H, W = scoremap.shape
    idx = np.argmax(scoremap.reshape(B, -1), axis=1)
    return np.stack([idx // W, idx % W], axis=1).astype(np.int32)


def seg_fg_det(scoremap):
    """utils/general.py:240-242: softmax -> fg prob -> round."""
    sm = T.softmax_last(scoremap.astype(np.float32))
    fg = sm[..., 1:].max(axis=3)
    return fg, T.round_half_even(fg)


def single_obj_scoremap(scoremap, literal=True):
    """utils/general.py:233-268.  Returns [B,H,W,1] float32 in {0,1}.

    literal=True: the reference's exact op sequence (21x21 dilation2d + multiply + round,
    max(H,W)//10 passes).  literal=False: equivalent boolean geodesic dilation (faster)."""
    B, H, W, _ = scoremap.shape
    fg, det = seg_fg_det(scoremap)
    max_loc = find_max_location(fg)                                        # :245
    num_passes = max(H, W) // (21 // 2)                                    # :256
    out = np.zeros((B, H, W, 1), np.float32)
    for i in range(B):
        obj = np.zeros((H, W), np.float32)
        obj[max_loc[i, 0], max_loc[i, 1]] = 1.0                            # :252-253
        if literal:
            for _ in range(num_passes):
                dil = T.dilation2d_21(obj)                                 # :259
                obj = T.round_half_even(det[i] * dil)                      # :261
        else:
            d = det[i] > 0.5
            o = obj > 0.5
            for _ in range(num_passes):
                o = d & _dilate_bool_21(o)
            obj = o.astype(np.float32)
        out[i, :, :, 0] = obj
    return out


def _dilate_bool_21(o):
    H, W = o.shape
    c = np.cumsum(np.pad(o.astype(np.int32), ((0, 0), (11, 10))), axis=1)
    hor = (c[:, 21:] - c[:, :-21]) > 0
    c = np.cumsum(np.pad(hor.astype(np.int32), ((11, 10), (0, 0))), axis=0)
    return (c[21:, :] - c[:-21, :]) > 0


def calc_center_bb(mask):
    """utils/general.py:271-328.  mask [B,H,W,1] (or [B,H,W]) -> center [B,2] (row,col),
    bb [B,2,2], crop_size [B,1].  Empty mask -> the reference's written fallbacks."""
    if mask.ndim == 4:
        mask = mask[..., 0]
    m = mask.astype(np.int32) == 1
    B = m.shape[0]
    center = np.zeros((B, 2), np.float32)
    bb = np.zeros((B, 2, 2), np.float32)
    size = np.zeros((B, 1), np.float32)
    f = np.float32
    for i in range(B):
        rows, cols = np.nonzero(m[i])
        if rows.size == 0:
            center[i] = (160.0, 160.0)                                     # :311-312
            size[i] = 100.0                                                # :319-320
            bb[i] = np.array([[np.inf, -np.inf], [np.inf, -np.inf]], np.float32)
            continue
        x_min, x_max = f(rows.min()), f(rows.max())                        # :297-300 (X = rows)
        y_min, y_max = f(cols.min()), f(cols.max())
        bb[i] = np.array([[x_min, x_max], [y_min, y_max]], np.float32)     # :302-304
        center[i] = (f(0.5) * (x_max + x_min), f(0.5) * (y_max + y_min))   # :307-309
        size[i] = max(x_max - x_min, y_max - y_min)                        # :316-318
    return center, bb, size


def crop_scale(crop_size_best):
    """nets/ColorHandPose3DNetwork.py:84-85."""
    f = np.float32
    cs = crop_size_best.astype(f) * f(1.25)
    with np.errstate(divide="ignore"):
        return np.minimum(np.maximum(f(CROP_SIZE) / cs, f(0.25)), f(5.0)).astype(f)


def crop_boxes(center, crop_size, scale, H, W):
    """utils/general.py:182-191 -- normalised (y1,x1,y2,x2); note the /H, /W (not H-1)."""
    f = np.float32
    scale = scale.reshape(-1).astype(f)
    loc = center.astype(f).reshape(-1, 2)
    css = f(crop_size) / scale
    half = np.floor(css / f(2.0)).astype(f)                                # float '//' 
    y1 = loc[:, 0] - half
    y2 = y1 + css
    x1 = loc[:, 1] - half
    x2 = x1 + css
    return np.stack([y1 / f(H), x1 / f(W), y2 / f(H), x2 / f(W)], axis=-1).astype(f)


def crop_image_from_xy(image, crop_location, crop_size, scale=1.0):
    """utils/general.py:163-196."""
    B, H, W, _ = image.shape
    scale = np.broadcast_to(np.asarray(scale, np.float32).reshape(-1), (B,))
    boxes = crop_boxes(np.asarray(crop_location), crop_size, scale, H, W)
    return T.crop_and_resize(image.astype(np.float32), boxes, crop_size, crop_size)


# --------------------------------------------------------------------------------------------
# Lifting  (nets/ColorHandPose3DNetwork.py:221-384, nets/PosePriorNetwork.py:59-159)
# --------------------------------------------------------------------------------------------
def inference_pose3d_can(scoremap32, hand_side, weights, dtype=np.float32, bottleneck=False):
    """nets/ColorHandPose3DNetwork.py:249-272 (bottleneck: nets/PosePriorNetwork.py:113-116)."""
    sc = "PosePrior"
    x = scoremap32.astype(dtype)
    B = x.shape[0]
    for i in range(3):
        x = _conv(x, weights, sc, "conv_pose_%d_1" % i, 1, dtype=dtype)
        x = _conv(x, weights, sc, "conv_pose_%d_2" % i, 2, dtype=dtype)
    x = np.concatenate([x.reshape(B, -1), hand_side.astype(dtype)], axis=1)  # :262-263
    x = _fc(x, weights, sc, "fc_rel0", True, dtype)
    x = _fc(x, weights, sc, "fc_rel1", True, dtype)
    if bottleneck:
        x = _fc(x, weights, sc, "fc_bottleneck", False, dtype)
    x = _fc(x, weights, sc, "fc_xyz", False, dtype)
    return x.reshape(B, NUM_KP, 3)


def rotation_estimation(scoremap32, hand_side, weights, dtype=np.float32):
    """nets/ColorHandPose3DNetwork.py:285-309."""
    sc = "ViewpointNet"
    x = scoremap32.astype(dtype)
    B = x.shape[0]
    for i in range(3):
        x = _conv(x, weights, sc, "conv_vp_%d_1" % i, 1, dtype=dtype)
        x = _conv(x, weights, sc, "conv_vp_%d_2" % i, 2, dtype=dtype)
    x = np.concatenate([x.reshape(B, -1), hand_side.astype(dtype)], axis=1)  # :297-298
    x = _fc(x, weights, sc, "fc_vp0", True, dtype)
    x = _fc(x, weights, sc, "fc_vp1", True, dtype)
    ux = _fc(x, weights, sc, "fc_vp_ux", False, dtype)
    uy = _fc(x, weights, sc, "fc_vp_uy", False, dtype)
    uz = _fc(x, weights, sc, "fc_vp_uz", False, dtype)
    return ux, uy, uz


def get_rot_mat(ux_b, uy_b, uz_b):
    """nets/ColorHandPose3DNetwork.py:311-334 + _stitch_mat_from_vecs :363-384 -> [B,3,3]."""
    ft = ux_b.dtype.type
    u_norm = np.sqrt(np.square(ux_b) + np.square(uy_b) + np.square(uz_b) + ft(1e-8))
    theta = u_norm
    st, ct = np.sin(theta)[:, 0], np.cos(theta)[:, 0]
    one_ct = (ft(1.0) - np.cos(theta))[:, 0]
    norm_fac = ft(1.0) / u_norm[:, 0]
    ux, uy, uz = ux_b[:, 0] * norm_fac, uy_b[:, 0] * norm_fac, uz_b[:, 0] * norm_fac
    vecs = [ct + ux * ux * one_ct, ux * uy * one_ct - uz * st, ux * uz * one_ct + uy * st,
            uy * ux * one_ct + uz * st, ct + uy * uy * one_ct, uy * uz * one_ct - ux * st,
            uz * ux * one_ct - uy * st, uz * uy * one_ct + ux * st, ct + uz * uz * one_ct]
    return np.stack(vecs, axis=0).reshape(3, 3, -1).transpose(2, 0, 1).astype(ux_b.dtype)


def flip_right_hand(coords, hand_side):
    """nets/ColorHandPose3DNetwork.py:239-242,336-361: mirror z when argmax(hand_side)==1."""
    right = np.argmax(hand_side, axis=1) == 1
    out = coords.copy()
    out[right, :, 2] = -out[right, :, 2]
    return out


def inference_pose3d(scoremap32, hand_side, weights, dtype=np.float32):
    """nets/ColorHandPose3DNetwork.py:221-247 -> (coord_xyz_rel_normed, coord_can, rot_mat)."""
    coord_can = inference_pose3d_can(scoremap32, hand_side, weights, dtype)
    ux, uy, uz = rotation_estimation(scoremap32, hand_side, weights, dtype)
    R = get_rot_mat(ux, uy, uz)
    flip = flip_right_hand(coord_can, hand_side)
    return np.matmul(flip, R).astype(coord_can.dtype), coord_can, R        # :245


# --------------------------------------------------------------------------------------------
# Orchestration
# --------------------------------------------------------------------------------------------
def inference(image, hand_side, weights, dtype=np.float32, literal_mask=True, forced_crop=None):
    """nets/ColorHandPose3DNetwork.py:61-99.  Returns the reference's 6-tuple
    (hand_scoremap, image_crop, scale_crop, center, keypoints_scoremap, keypoint_coord3d).

    forced_crop=(center, scale_crop): teacher-force the crop parameters (stage-wise parity)."""
    image = image.astype(np.float32)
    hand_scoremap = inference_detection(image, weights, dtype)[-1].astype(np.float32)     # :78-79
    if forced_crop is None:
        hand_mask = single_obj_scoremap(hand_scoremap, literal=literal_mask)               # :82
        center, _, crop_size_best = calc_center_bb(hand_mask)                              # :83
        scale_crop = crop_scale(crop_size_best)                                            # :84-85
    else:
        center, scale_crop = forced_crop
    image_crop = crop_image_from_xy(image, center, CROP_SIZE, scale_crop)                  # :86
    s32 = inference_pose2d(image_crop, weights, dtype)[-1]                                 # :89-90
    coord3d = inference_pose3d(s32, hand_side, weights, dtype)[0]                          # :93
    s32f = s32.astype(np.float32)
    kp_scoremap = T.resize_bilinear_tf1(s32f, CROP_SIZE, CROP_SIZE)                        # :96-97
    return hand_scoremap, image_crop, scale_crop, center, kp_scoremap, coord3d.astype(np.float32)


def inference2d(image, weights, dtype=np.float32, literal_mask=True):
    """nets/ColorHandPose3DNetwork.py:101-129 -> (keypoints_scoremap, image_crop, scale_crop, center)."""
    image = image.astype(np.float32)
    hand_scoremap = inference_detection(image, weights, dtype)[-1].astype(np.float32)
    hand_mask = single_obj_scoremap(hand_scoremap, literal=literal_mask)
    center, _, crop_size_best = calc_center_bb(hand_mask)
    scale_crop = crop_scale(crop_size_best)
    image_crop = crop_image_from_xy(image, center, CROP_SIZE, scale_crop)
    s32 = inference_pose2d(image_crop, weights, dtype)[-1].astype(np.float32)
    return T.resize_bilinear_tf1(s32, CROP_SIZE, CROP_SIZE), image_crop, scale_crop, center


# --------------------------------------------------------------------------------------------
# Kinematic-chain transform (utils/relative_trafo.py) -- needed by the 'local' lifting variants
# --------------------------------------------------------------------------------------------
KINEMATIC_CHAIN_PARENT = {0: None, 4: None, 3: 4, 2: 3, 1: 2, 8: None, 7: 8, 6: 7, 5: 6, 12: None, 11: 12, 10: 11, 9: 10,
                          16: None, 15: 16, 14: 15, 13: 14, 20: None, 19: 20, 18: 19, 17: 18}      # utils/relative_trafo.py:148-171
KINEMATIC_CHAIN_LIST = [0, 4, 3, 2, 1, 8, 7, 6, 5, 12, 11, 10, 9, 16, 15, 14, 13, 20, 19, 18, 17]   # :174-179


def _rot_x_hom(a):   # utils/relative_trafo.py:49-57
    c, s_ = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0, 0], [0, c, -s_, 0], [0, s_, c, 0], [0, 0, 0, 1]], a.dtype)


def _rot_y_hom(a):   # :60-68
    c, s_ = np.cos(a), np.sin(a)
    return np.array([[c, 0, s_, 0], [0, 1, 0, 0], [-s_, 0, c, 0], [0, 0, 0, 1]], a.dtype)


def _trans_hom(t):   # :82-90 -- translation along z only
    m = np.eye(4, dtype=t.dtype)
    m[2, 3] = t
    return m


def _atan2_ref(y, x):
    """utils/relative_trafo.py:27-46 (the reference's own atan2 built from atan)."""
    ft = type(y)
    pi = ft(3.141592653589793)
    tan = np.arctan(y / (x + ft(1e-8)))
    tan_c = tan + (pi if (x + ft(1e-8)) < 0 else ft(0))
    tan_02pi = tan_c + (ft(2) * pi if tan_c < 0 else ft(0))
    return ft(tan_02pi + (ft(-2) * pi if tan_02pi > pi else ft(0)))


def bone_rel_trafo_inv(coords_rel):
    """utils/relative_trafo.py:243-295: [B,21,3] (length, angle_x, angle_y) per bone -> xyz [B,21,3]."""
    coords_rel = np.asarray(coords_rel)
    if coords_rel.ndim == 2:
        coords_rel = coords_rel[None]
    B = coords_rel.shape[0]
    out = np.zeros((B, 21, 3), coords_rel.dtype)
    x0 = np.array([0, 0, 0, 1], coords_rel.dtype)
    for b in range(B):
        trafo = {}
        for bone in KINEMATIC_CHAIN_LIST:
            parent = KINEMATIC_CHAIN_PARENT[bone]
            T = np.eye(4, dtype=coords_rel.dtype) if parent is None else trafo[parent]
            length, ax, ay = coords_rel[b, bone]
            T_this = _trans_hom(-length) @ (_rot_x_hom(-ax) @ _rot_y_hom(-ay))      # _forward :102-114
            T = T_this @ T
            out[b, bone] = (np.linalg.inv(T) @ x0)[:3]
            trafo[bone] = T
    return out


def bone_rel_trafo(coords_xyz):
    """utils/relative_trafo.py:182-240 (forward direction; only used here for round-trip property tests)."""
    coords_xyz = np.asarray(coords_xyz).reshape(-1, 21, 3)
    B = coords_xyz.shape[0]
    dt = coords_xyz.dtype
    out = np.zeros((B, 21, 3), dt)
    for b in range(B):
        trafo = {}
        for bone in KINEMATIC_CHAIN_LIST:
            parent = KINEMATIC_CHAIN_PARENT[bone]
            if parent is None:
                T = np.eye(4, dtype=dt)
                delta = np.append(coords_xyz[b, bone], dt.type(1))
            else:
                T = trafo[parent]
                d3 = (T @ np.append(coords_xyz[b, bone], dt.type(1)) - T @ np.append(coords_xyz[b, parent], dt.type(1)))[:3]
                delta = np.append(d3, dt.type(1))
            length = np.sqrt(delta[0] ** 2 + delta[1] ** 2 + delta[2] ** 2)             # _backward :117-142
            ay = _atan2_ref(delta[0], delta[2])
            tmp = _rot_y_hom(-ay) @ delta
            ax = _atan2_ref(-tmp[1], tmp[2])
            T_this = _trans_hom(-length) @ (_rot_x_hom(-ax) @ _rot_y_hom(-ay))
            trafo[bone] = T_this @ T
            out[b, bone] = (length, ax, ay)
    return out


def pose_prior_inference(scoremap256, hand_side, weights, variant, dtype=np.float32):
    """nets/PosePriorNetwork.py:59-95 for all five variants (direct / bottleneck / local / local_w_xyz_loss / proposed).
    Returns (coord_xyz_rel_normed, coord3d, R)."""
    pooled = T.avg_pool_8x8(scoremap256.astype(np.float32))                # :61
    if variant == "direct":
        c = inference_pose3d_can(pooled, hand_side, weights, dtype)
        return c, c, None
    if variant == "bottleneck":
        c = inference_pose3d_can(pooled, hand_side, weights, dtype, bottleneck=True)
        return c, c, None
    if variant in ("local", "local_w_xyz_loss"):                           # :70-75
        c = inference_pose3d_can(pooled, hand_side, weights, dtype)
        return bone_rel_trafo_inv(c.astype(np.float32)), c, None
    if variant == "proposed":
        out, can, R = inference_pose3d(pooled, hand_side, weights, dtype)
        return out, can, R
    raise AssertionError("Unknown variant.")                                # :93


def detect_keypoints(scoremaps):
    """utils/general.py:331-344 -- per-channel first-occurrence arg-max -> float64 [C,2] (v,u)."""
    if scoremaps.ndim == 4:
        scoremaps = np.squeeze(scoremaps)
    s = scoremaps.shape
    assert len(s) == 3, "This function was only designed for 3D Scoremaps."
    assert (s[2] < s[1]) and (s[2] < s[0]), "Probably the input is not correct, because [H, W, C] is expected."
    out = np.zeros((s[2], 2))
    for i in range(s[2]):
        v, u = np.unravel_index(np.argmax(scoremaps[:, :, i]), (s[0], s[1]))
        out[i] = (v, u)
    return out


def trafo_coords(keypoints_crop_coords, centers, scale, crop_size):
    """utils/general.py:347-357."""
    k = np.copy(keypoints_crop_coords)
    k -= crop_size // 2
    k /= scale
    k += centers
    return k


# --------------------------------------------------------------------------------------------
# Record formats and evaluation bookkeeping
# --------------------------------------------------------------------------------------------
def decode_rhd_record(record):
    """data/BinaryDbReader.py:103-208 (raw items only) for ONE 410 520-byte record (bytes / uint8 array)."""
    raw = np.frombuffer(bytes(record), dtype=np.uint8)
    assert raw.size == 410520, "Doesnt add up."                                          # :210
    f = np.frombuffer(bytes(record[:876]), dtype=np.float32)
    u8 = raw[878:]
    img = u8[:307200].reshape(320, 320, 3).astype(np.float32) / np.float32(255.0) - np.float32(0.5)   # :176-182
    parts = u8[307200:409600].reshape(320, 320).astype(np.int32)
    hand = parts > 1
    return {"keypoint_xyz": f[:126].reshape(42, 3), "keypoint_uv": f[126:210].reshape(42, 2).astype(np.int32).astype(np.float32),
            "cam_mat": f[210:219].reshape(3, 3), "image": img, "hand_parts": parts,
            "hand_mask": np.stack([~hand, hand], 2).astype(np.int32), "keypoint_vis": u8[409600:409642].astype(bool)}


def decode_stb_record(record, subsample=2):
    """data/BinaryDbReaderSTB.py:99-185 (raw items) + eval_full.py:50 (legacy bilinear 480x640 -> 240x320 = every 2nd pixel)."""
    raw = np.frombuffer(bytes(record), dtype=np.uint8)
    assert raw.size == 922104
    f = np.frombuffer(bytes(record[:504]), dtype=np.float32)
    img = raw[504:].reshape(480, 640, 3).astype(np.float32) / np.float32(255.0) - np.float32(0.5)
    if subsample > 1:
        img = T.resize_bilinear_tf1(img[None], 480 // subsample, 640 // subsample)[0]
    uvv = f[63:126].reshape(21, 3)
    return {"keypoint_xyz": f[:63].reshape(21, 3), "keypoint_uv": uvv[:, :2], "keypoint_vis": uvv[:, 2] > 0.5, "image": img}


class EvalUtil:
    """utils/general.py:522-611, restated per sample."""
    def __init__(self, num_kp=21):
        self.num_kp = num_kp
        self.data = [[] for _ in range(num_kp)]

    def feed(self, keypoint_gt, keypoint_vis, keypoint_pred):
        gt, pr = np.squeeze(keypoint_gt), np.squeeze(keypoint_pred)
        vis = np.squeeze(keypoint_vis).astype(bool)
        d = np.sqrt(np.sum(np.square(gt - pr), axis=1))
        for i in range(gt.shape[0]):
            if vis[i]:
                self.data[i].append(d[i])

    def get_measures(self, val_min, val_max, steps):
        trapz = getattr(np, "trapezoid", None) or np.trapz
        th = np.linspace(val_min, val_max, steps)
        norm = trapz(np.ones_like(th), th)
        mean, med, auc, curves = [], [], [], []
        for k in range(self.num_kp):
            if not self.data[k]:
                continue
            d = np.array(self.data[k])
            mean.append(d.mean()); med.append(np.median(d))
            c = np.array([np.mean((d <= t).astype(float)) for t in th])
            curves.append(c); auc.append(trapz(c, th) / norm)
        return np.mean(mean), np.mean(med), np.mean(auc), np.mean(np.array(curves), 0), th
