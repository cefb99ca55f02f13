"""Shared by tests/test_gpu_configs.py and scripts/mismatch_report.py: parity statistics of the CUDA pipeline against the CPU
oracle at the BASELINE.json batch sizes (free-running and teacher-forced), per precision mode.  Test infrastructure only."""
import numpy as np
import torch

from hand3d_b200 import weights as Wt
from oracle import hand3d_oracle as O
from oracle import tf1_ops as T

f32 = np.float32


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def mixed_images(n, H=320, W=320, seed=1):
    """half seeded noise images, half blob images (varied masks / crops that leave the image)"""
    a = Wt.synthetic_images(n - n // 2, H, W, seed=seed)
    b = Wt.synthetic_blob_images(n // 2, H, W, seed=seed + 4)
    return np.concatenate([a, b], 0)


def keypoint_stats(dev_uv, dev_map, ref_map):
    """dev_uv [B,21,2] int, maps [B,H,W,21]: identical-index rate vs the oracle's arg-max, and for every differing key-point the
    oracle-map margin between its maximum and the value at the device's location (a near-tie explains the difference)."""
    B = dev_uv.shape[0]
    same, margins = 0, []
    for b in range(B):
        kp_ref = O.detect_keypoints(ref_map[b]).astype(np.int64)
        for c in range(21):
            if np.array_equal(dev_uv[b, c], kp_ref[c]):
                same += 1
            else:
                v, u = dev_uv[b, c]
                margins.append(float(ref_map[b, :, :, c].max() - ref_map[b, v, u, c]))
    return same, 21 * B, margins


def full_pipeline_stats(ctx, img, hs, wd, ref, precision, chunk=32):
    """ref = O.inference(img, hs, wd, literal_mask=False).  Returns a dict of rates and max abs errors for `precision`."""
    ctx.load_weights(wd)
    ctx.set_precision(precision)
    B = img.shape[0]
    out = {"precision": precision, "images": int(B)}
    agree, same_free, tot_free, same_tf, tot_tf = 0, 0, 0, 0, 0
    e_seg = e_map = e_3d = 0.0
    margins_tf, margins_free = [], []
    for lo in range(0, B, chunk):
        sl = slice(lo, min(B, lo + chunk))
        x, h = dev(img[sl]), dev(hs[sl])
        free = ctx.pipeline(x, h, True)
        g = {k: v.cpu().numpy() for k, v in free.items() if v is not None}
        ok = (g["center"] == ref[3][sl]).all(1) & (g["scale_crop"] == ref[2][sl]).all(1)
        agree += int(ok.sum())
        if ok.any():
            s, t, m = keypoint_stats(g["keypoints_uv"][ok], g["keypoints_scoremap"][ok], ref[4][sl][ok])
            same_free += s; tot_free += t; margins_free += m
        forced = ctx.pipeline(x, h, True, force_center=dev(ref[3][sl]), force_scale=dev(ref[2][sl]))
        f = {k: v.cpu().numpy() for k, v in forced.items() if v is not None}
        assert np.array_equal(f["image_crop"], ref[1][sl]), "crop differs from the oracle's under identical crop parameters"
        e_seg = max(e_seg, float(np.abs(f["hand_scoremap"] - ref[0][sl]).max()))
        e_map = max(e_map, float(np.abs(f["keypoints_scoremap"] - ref[4][sl]).max()))
        e_3d = max(e_3d, float(np.abs(f["keypoint_coord3d"] - ref[5][sl]).max()))
        s, t, m = keypoint_stats(f["keypoints_uv"], f["keypoints_scoremap"], ref[4][sl])
        same_tf += s; tot_tf += t; margins_tf += m
        for b in range(f["keypoints_uv"].shape[0]):       # the device's indices are the exact arg-max of the device's own map
            assert np.array_equal(f["keypoints_uv"][b], O.detect_keypoints(f["keypoints_scoremap"][b]).astype(np.int32))
    out.update(crop_params_agree=agree, crop_params_agree_rate=agree / B,
               keypoints_identical_free=same_free, keypoints_compared_free=tot_free,
               keypoints_identical_free_rate=(same_free / tot_free) if tot_free else None,
               keypoints_identical_forced=same_tf, keypoints_compared_forced=tot_tf, keypoints_identical_forced_rate=same_tf / tot_tf,
               max_margin_of_differing_keypoints=max(margins_tf + margins_free) if (margins_tf or margins_free) else 0.0,
               max_abs_err_hand_scoremap=e_seg, max_abs_err_keypoints_scoremap=e_map, max_abs_err_coord3d=e_3d)
    return out


def posenet_stats(ctx, crops, wd, ref_map, precision, chunk=32):
    """ref_map = resize_x8(O.inference_pose2d(crops)[-1]) [B,256,256,21]."""
    ctx.load_weights(wd)
    ctx.set_precision(precision)
    B = crops.shape[0]
    e, same, tot, margins = 0.0, 0, 0, []
    for lo in range(0, B, chunk):
        sl = slice(lo, min(B, lo + chunk))
        r = ctx.pose2d(dev(crops[sl]))
        sm, uv = r["keypoints_scoremap"].cpu().numpy(), r["keypoints_uv"].cpu().numpy()
        e = max(e, float(np.abs(sm - ref_map[sl]).max()))
        s, t, m = keypoint_stats(uv, sm, ref_map[sl])
        same += s; tot += t; margins += m
        for b in range(uv.shape[0]):
            assert np.array_equal(uv[b], O.detect_keypoints(sm[b]).astype(np.int32))
    return {"precision": precision, "images": int(B), "max_abs_err_keypoints_scoremap": e, "keypoints_identical": same,
            "keypoints_compared": tot, "keypoints_identical_rate": same / tot, "max_margin_of_differing_keypoints": max(margins) if margins else 0.0}


def posenet_reference(crops, wd):
    return T.resize_bilinear_tf1(O.inference_pose2d(crops, wd)[-1], crops.shape[1], crops.shape[2])
