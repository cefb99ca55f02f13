"""Known-answer tests that pin the oracle to the TF-1.3 semantics listed in SURVEY.md section 9.
(The reference itself ships no tests / golden vectors: parity is otherwise unpinned.)"""
import numpy as np
import pytest

from oracle import hand3d_oracle as O
from oracle import tf1_ops as T

f32 = np.float32


def test_same_padding_table():
    assert T.same_pad(320, 3, 1) == (1, 1)
    assert T.same_pad(32, 7, 1) == (3, 3)
    assert T.same_pad(32, 1, 1) == (0, 0)
    assert T.same_pad(32, 3, 2) == (0, 1)          # asymmetric: window i covers inputs 2i..2i+2
    assert T.same_pad(5, 3, 2) == (1, 1)


def test_conv_stride2_kat():
    a, b, c, d = 1.0, 2.0, 3.0, 5.0
    w0, w1, w2 = 0.5, -1.0, 2.0
    x = np.zeros((1, 1, 4, 1), f32); x[0, 0, :, 0] = (a, b, c, d)
    w = np.zeros((1, 3, 1, 1), f32); w[0, :, 0, 0] = (w0, w1, w2)
    y = T.conv2d_same(x, w, np.zeros(1, f32), stride=2)
    np.testing.assert_allclose(y[0, 0, :, 0], [w0 * a + w1 * b + w2 * c, w0 * c + w1 * d], rtol=1e-6)


def test_conv_is_cross_correlation_hwio():
    x = np.zeros((1, 3, 3, 2), f32); x[0, 0, 2, 1] = 1.0       # top-right pixel, channel 1
    w = np.arange(3 * 3 * 2 * 4, dtype=f32).reshape(3, 3, 2, 4)
    y = T.conv2d_same(x, w, np.zeros(4, f32))
    # centre output (1,1) sees input (0,2) through tap (kh=0, kw=2)
    np.testing.assert_array_equal(y[0, 1, 1], w[0, 2, 1])


def test_leaky_relu():
    np.testing.assert_allclose(T.leaky_relu(np.array([-2.0, 0.0, 3.0], f32)), [-0.02, 0.0, 3.0], rtol=1e-6)


def test_max_pool_valid():
    x = np.arange(16, dtype=f32).reshape(1, 4, 4, 1)
    np.testing.assert_array_equal(T.max_pool_2x2(x)[0, :, :, 0], [[5, 7], [13, 15]])


def test_resize_bilinear_legacy_kat():
    x = np.array([0.0, 10.0], f32).reshape(1, 1, 2, 1)
    np.testing.assert_array_equal(T.resize_bilinear_tf1(x, 1, 4)[0, 0, :, 0], [0, 5, 10, 10])
    x = np.arange(6 * 8, dtype=f32).reshape(1, 6, 8, 1)
    np.testing.assert_array_equal(T.resize_bilinear_tf1(x, 3, 4)[0, :, :, 0], x[0, ::2, ::2, 0])   # pure subsampling
    assert T.resize_bilinear_tf1(x, 6, 8) is x


def test_resize_x8_edge_replication_and_ramp():
    x = np.arange(4, dtype=f32).reshape(1, 1, 4, 1) * f32(8.0)
    y = T.resize_bilinear_tf1(np.repeat(x, 4, axis=1), 32, 32)[0, 5, :, 0]
    np.testing.assert_array_equal(y[:25], np.arange(25, dtype=f32))
    np.testing.assert_array_equal(y[24:], np.full(8, 24.0, f32))


def test_round_half_even():
    np.testing.assert_array_equal(T.round_half_even(np.array([0.5, 1.5, 2.5, 0.50000006, 1 + 1 / 441, 1 / 441], f32)),
                                  [0, 2, 2, 1, 1, 0])


def test_softmax_two_class():
    p = T.softmax_last(np.array([[0.0, 0.0], [0.0, 20.0], [1.0, -1.0]], f32))
    assert p[0, 1] == f32(0.5) and p[1, 1] == f32(1.0)           # saturates to exactly 1
    np.testing.assert_allclose(p[2, 1], 1 / (1 + np.exp(2.0)), rtol=1e-6)


def test_dilation_kat():
    o = np.zeros((40, 50), f32); o[0, 25] = 1.0
    d = T.dilation2d_21(o)
    exp = np.full((40, 50), f32(1 / 441), f32); exp[0:11, 15:36] += 1.0
    np.testing.assert_allclose(d, exp, rtol=1e-6)


def test_crop_and_resize_kats():
    v = np.arange(5, dtype=f32).reshape(1, 5, 1, 1) * np.ones((1, 5, 5, 1), f32)
    out = T.crop_and_resize(v, np.array([[0.2, 0.0, 0.8, 1.0]], f32), 3, 3)
    np.testing.assert_allclose(out[0, :, 0, 0], [0.8, 2.0, 3.2], rtol=1e-6)
    out = T.crop_and_resize(v, np.array([[-0.1, 0.0, 0.8, 1.0]], f32), 3, 3)
    assert out[0, 0, 0, 0] == 0.0 and out[0, 1, 0, 0] > 0


def test_crop_boxes_use_H_not_Hminus1():
    boxes = O.crop_boxes(np.array([[160.0, 160.0]], f32), 256, np.array([1.0], f32), 320, 320)
    y1, x1, y2, x2 = boxes[0]
    assert y1 == f32(32.0) / f32(320.0) and y2 == f32(288.0) / f32(320.0)
    in_y0 = y1 * f32(319)
    scale = (y2 - y1) * f32(319) / f32(255)
    np.testing.assert_allclose([in_y0, scale], [31.9, 0.8 * 319 / 255], rtol=1e-6)   # 31.9 + 1.000784*y


def test_crop_ramp_returns_sampling_coordinates():
    H = W = 320
    img = np.zeros((1, H, W, 3), f32)
    img[0, :, :, 0] = np.arange(H, dtype=f32)[:, None]
    img[0, :, :, 1] = np.arange(W, dtype=f32)[None, :]
    crop = O.crop_image_from_xy(img, np.array([[100.0, 200.0]], f32), 256, np.array([[2.0]], f32))
    in_y = f32(36.0 / 320.0) * f32(319) + np.arange(256, dtype=f32) * (f32(128.0 / 320.0) * f32(319) / f32(255))
    np.testing.assert_allclose(crop[0, :, 0, 0], in_y, atol=1e-4)


def test_find_max_location_first_occurrence():
    s = np.zeros((2, 6, 7), f32); s[0, 2, 3] = 1; s[0, 4, 1] = 1; s[1, 5, 6] = 2
    np.testing.assert_array_equal(O.find_max_location(s), [[2, 3], [5, 6]])


def _brute_grow(det, seed, passes):
    H, W = det.shape
    o = np.zeros((H, W), bool); o[seed] = True
    for _ in range(passes):
        n = np.zeros_like(o)
        ys, xs = np.nonzero(o)
        for y, x in zip(ys, xs):
            n[max(0, y - 10):y + 11, max(0, x - 10):x + 11] = True
        o = n & det
    return o


@pytest.mark.parametrize("seed", [0, 1])
def test_single_obj_scoremap_literal_vs_bool_vs_brute(seed):
    rng = np.random.default_rng(seed)
    H, W = 48, 64
    low = rng.normal(size=(1, 6, 8, 2)).astype(f32) * 2
    sm = T.resize_bilinear_tf1(low, H, W)
    a = O.single_obj_scoremap(sm, literal=True)
    b = O.single_obj_scoremap(sm, literal=False)
    np.testing.assert_array_equal(a, b)
    fg, det = O.seg_fg_det(sm)
    loc = tuple(O.find_max_location(fg)[0])
    np.testing.assert_array_equal(a[0, :, :, 0] > 0.5, _brute_grow(det[0] > 0.5, loc, max(H, W) // 10))


def test_grower_jumps_gaps_up_to_10_and_stops_after_passes():
    H, W = 320, 320
    sm = np.zeros((1, H, W, 2), f32); sm[..., 0] = 5.0
    sm[0, 100, 5:8, 1] = 10.0       # seed blob (largest prob at first pixel)
    sm[0, 100, 5, 1] = 11.0
    sm[0, 100, 17, 1] = 10.0        # gap of 9 empty pixels (8..16) -> Chebyshev distance 10 from x=7: reached
    sm[0, 100, 28, 1] = 10.0        # distance 11 from 17: not reached
    m = O.single_obj_scoremap(sm, literal=False)[0, :, :, 0]
    assert m[100, 5] == 1 and m[100, 17] == 1 and m[100, 28] == 0
    # a 1-pixel-wide line of 320 px grows 10 px per pass: 32 passes cover 5+320 px -> fully covered
    sm2 = np.zeros((1, H, W, 2), f32); sm2[..., 0] = 5.0; sm2[0, 7, :, 1] = 10.0; sm2[0, 7, 0, 1] = 11.0
    assert O.single_obj_scoremap(sm2, literal=False)[0, 7].sum() == 320


def test_calc_center_bb_and_fallback():
    m = np.zeros((2, 320, 320, 1), f32)
    m[0, 10:21, 30:71] = 1
    c, bb, s = O.calc_center_bb(m)
    np.testing.assert_array_equal(c[0], [15.0, 50.0]); assert s[0, 0] == 40.0
    np.testing.assert_array_equal(c[1], [160.0, 160.0]); assert s[1, 0] == 100.0
    sc = O.crop_scale(s)
    np.testing.assert_allclose(sc[:, 0], [5.0, 2.048], rtol=1e-6)       # 256/50=5.12 -> clipped to 5
    assert O.crop_scale(np.array([[0.0]], f32))[0, 0] == 5.0             # size 0 -> inf -> 5
    assert O.crop_scale(np.array([[10000.0]], f32))[0, 0] == 0.25


def test_rot_mat_rodrigues():
    z = np.zeros((1, 1), f32)
    R = O.get_rot_mat(z, z, np.full((1, 1), np.pi / 2, f32))
    np.testing.assert_allclose(R[0], [[0, -1, 0], [1, 0, 0], [0, 0, 1]], atol=1e-6)
    R0 = O.get_rot_mat(z, z, z)
    np.testing.assert_allclose(R0[0], np.eye(3), atol=1e-6)


def test_flip_and_matmul_order():
    c = np.arange(63, dtype=f32).reshape(1, 21, 3)
    f = O.flip_right_hand(np.repeat(c, 2, 0), np.array([[1, 0], [0, 1]], f32))
    np.testing.assert_array_equal(f[0], c[0]); np.testing.assert_array_equal(f[1, :, 2], -c[0, :, 2])


def test_detect_keypoints_and_trafo():
    s = np.zeros((256, 256, 21), f32)
    for i in range(21):
        s[3 * i, 5 * i + 1, i] = 1.0
        s[3 * i + 1, 0, i] = 1.0        # later duplicate of the max: first occurrence must win
    kp = O.detect_keypoints(s[None])
    assert kp.dtype == np.float64
    np.testing.assert_array_equal(kp[:, 0], 3 * np.arange(21)); np.testing.assert_array_equal(kp[:, 1], 5 * np.arange(21) + 1)
    t = O.trafo_coords(kp, np.array([[100.0, 50.0]]), np.array([[2.0]]), 256)
    np.testing.assert_allclose(t, (kp - 128) / 2 + [100, 50])


def test_fc_check_numerics():
    with pytest.raises(FloatingPointError):
        T.fully_connected(np.zeros((1, 2), f32), np.array([[np.nan], [0.0]], f32), np.zeros(1, f32))


def test_avg_pool():
    x = np.arange(16 * 16, dtype=f32).reshape(1, 16, 16, 1)
    np.testing.assert_allclose(T.avg_pool_8x8(x)[0, :, :, 0], [[x[0, :8, :8].mean(), x[0, :8, 8:].mean()],
                                                             [x[0, 8:, :8].mean(), x[0, 8:, 8:].mean()]], rtol=1e-6)


def test_bone_rel_trafo_round_trip_and_kat():
    """utils/relative_trafo.py: inv(fwd(xyz)) == xyz; a single bone of length L with zero angles points along +z."""
    rng = np.random.default_rng(3)
    xyz = rng.normal(size=(2, 21, 3))
    np.testing.assert_allclose(O.bone_rel_trafo_inv(O.bone_rel_trafo(xyz)), xyz, atol=1e-6)   # the reference atan2 adds 1e-8 to x
    rel = np.zeros((1, 21, 3)); rel[0, :, 0] = 1.0           # every bone: length 1, no articulation
    out = O.bone_rel_trafo_inv(rel)[0]
    np.testing.assert_allclose(out[0], [0, 0, 1], atol=1e-12)          # root key-point
    np.testing.assert_allclose(out[4], [0, 0, 1], atol=1e-12)          # first bone of a finger
    np.testing.assert_allclose(out[1], [0, 0, 4], atol=1e-12)          # finger tip: 4 bones stacked along z
    rel[0, 4, 2] = np.pi / 2                                           # rotate the first thumb bone about y by 90 degrees
    np.testing.assert_allclose(O.bone_rel_trafo_inv(rel)[0, 4], [1, 0, 0], atol=1e-12)


def test_rhd_record_layout_kat():
    """data/BinaryDbReader.py:103-208: 42x3 f32 | 42x2 f32 | 9 f32 | 2 B | 320x320x3 u8 | 320x320 u8 | 42 u8 = 410520 B."""
    xyz = np.arange(126, dtype=f32).reshape(42, 3); uv = (np.arange(84, dtype=f32) + 0.75).reshape(42, 2); K = np.arange(9, dtype=f32)
    img = np.zeros((320, 320, 3), np.uint8); img[1, 2, 0] = 255; img[3, 4, 1] = 51
    parts = np.zeros((320, 320), np.uint8); parts[5, 6] = 1; parts[7, 8] = 2
    vis = np.zeros(42, np.uint8); vis[41] = 1
    rec = xyz.tobytes() + uv.tobytes() + K.tobytes() + b"\xff\xff" + img.tobytes() + parts.tobytes() + vis.tobytes()
    d = O.decode_rhd_record(rec)
    np.testing.assert_array_equal(d["keypoint_xyz"], xyz)
    np.testing.assert_array_equal(d["keypoint_uv"], np.floor(uv))          # cast to int32 and back (:151-154)
    np.testing.assert_array_equal(d["cam_mat"].ravel(), K)
    assert d["image"][1, 2, 0] == f32(0.5) and d["image"][0, 0, 0] == f32(-0.5) and d["image"][3, 4, 1] == f32(51) / f32(255) - f32(0.5)
    assert d["hand_mask"][7, 8, 1] == 1 and d["hand_mask"][5, 6, 1] == 0 and d["hand_mask"][5, 6, 0] == 1   # hand = parts > 1
    assert d["keypoint_vis"][41] and not d["keypoint_vis"][0]
    with pytest.raises(AssertionError):
        O.decode_rhd_record(rec[:-1])


def test_eval_util_kat():
    e = O.EvalUtil(num_kp=2)
    e.feed(np.array([[0.0, 0.0], [0.0, 0.0]]), np.array([1, 0]), np.array([[3.0, 4.0], [9.0, 9.0]]))    # dist 5, second invisible
    e.feed(np.array([[0.0, 0.0], [0.0, 0.0]]), np.array([1, 0]), np.array([[0.0, 1.0], [9.0, 9.0]]))    # dist 1
    mean, med, auc, curve, th = e.get_measures(0.0, 10.0, 11)
    assert mean == 3.0 and med == 3.0
    np.testing.assert_allclose(curve, [0, .5, .5, .5, .5, 1, 1, 1, 1, 1, 1])
    np.testing.assert_allclose(auc, (0.25 + 0.5 * 3 + 0.75 + 5.0) / 10.0)
