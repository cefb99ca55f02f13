"""Op-level parity pins for oracle/tf1_ops.py that the builder of the oracle did not derive:

(1) the input / expected-output vectors of TensorFlow's OWN op tests (r1.3 tree), transcribed from the public test
    sources named next to each case -- TensorFlow itself is not installable here, its published test tables are the
    closest thing to running it;
(2) independent implementations of the same ops: scipy.signal.correlate2d (conv2d), scipy.ndimage.grey_dilation
    (dilation2d), scipy.special.softmax, and scalar-loop transcriptions of the published CPU kernels
    (crop_and_resize_op.cc, resize_bilinear_op.cc) that share no code with the vectorised oracle.

Reference call sites whose semantics hang on these ops: utils/general.py:46 (conv2d SAME), :195 (crop_and_resize),
:259 (dilation2d), :240-242 (softmax / round), nets/ColorHandPose3DNetwork.py:97,128,166 (legacy resize).
The same vectors are run through the CUDA kernels in tests/test_gpu_tf_vectors.py.
"""
import math

import numpy as np
import pytest

from oracle import tf1_ops as T

f32 = np.float32


def _seq(shape):
    """TensorFlow's op tests fill tensors with 1, 2, 3, ... in row-major order."""
    return np.arange(1, int(np.prod(shape)) + 1, dtype=f32).reshape(shape)


# ------------------------------------------------------------------------------------------ conv2d
# tensorflow/python/kernel_tests/conv_ops_test.py (Conv2DTest); inputs and filters are 1..N row-major, NHWC / HWIO
TF_CONV_SAME = [
    # name, in shape, filter shape, stride, expected (flattened NHWC)
    ("testConv2D1x1Filter", (1, 2, 3, 3), (1, 1, 3, 3), 1,
     [30, 36, 42, 66, 81, 96, 102, 126, 150, 138, 171, 204, 174, 216, 258, 210, 261, 312]),
    ("testConv2D2x2FilterStride2Same", (1, 2, 3, 3), (2, 2, 3, 3), 2, [2271, 2367, 2463, 1230, 1305, 1380]),
    ("testConv2DKernelSmallerThanStrideSame_3x3", (1, 3, 3, 1), (1, 1, 1, 1), 2, [1, 3, 7, 9]),
    ("testConv2DKernelSmallerThanStrideSame_4x4", (1, 4, 4, 1), (1, 1, 1, 1), 2, [1, 3, 9, 11]),
    ("testConv2DKernelSmallerThanStrideSame_2x2s3", (1, 4, 4, 1), (2, 2, 1, 1), 3, [44, 28, 41, 16]),
]
TF_CONV_VALID = [
    ("testConv2D2x2Filter", (1, 2, 3, 3), (2, 2, 3, 3), 1, [2271, 2367, 2463, 2901, 3033, 3165]),
    ("testConv2D1x2Filter", (1, 2, 3, 3), (1, 2, 3, 3), 1,
     [231, 252, 273, 384, 423, 462, 690, 765, 840, 843, 936, 1029]),
    ("testConv2D2x2FilterStride2", (1, 2, 3, 3), (2, 2, 3, 3), 2, [2271, 2367, 2463]),
]


def _conv_scipy(x, w, stride, padding):
    """Independent conv2d: zero padding per TF's SAME rule, then scipy.signal.correlate2d per (cin, cout) pair."""
    from scipy.signal import correlate2d
    B, H, W, Cin = x.shape
    kh, kw, _, Cout = w.shape
    if padding == "SAME":
        Ho, Wo = -(-H // stride), -(-W // stride)
        th = max((Ho - 1) * stride + kh - H, 0); tw = max((Wo - 1) * stride + kw - W, 0)
        xp = np.pad(x.astype(np.float64), ((0, 0), (th // 2, th - th // 2), (tw // 2, tw - tw // 2), (0, 0)))
    else:
        xp = x.astype(np.float64)
    out = []
    for b in range(B):
        maps = []
        for co in range(Cout):
            acc = sum(correlate2d(xp[b, :, :, ci], w[:, :, ci, co].astype(np.float64), mode="valid") for ci in range(Cin))
            maps.append(acc[::stride, ::stride])
        out.append(np.stack(maps, -1))
    return np.stack(out)


@pytest.mark.parametrize("case", TF_CONV_SAME, ids=[c[0] for c in TF_CONV_SAME])
def test_tf_conv2d_same_vectors(case):
    _, xs, ws, stride, expected = case
    x, w = _seq(xs), _seq(ws)
    y = T.conv2d_same(x, w, np.zeros(ws[3], f32), stride)
    np.testing.assert_array_equal(y.reshape(-1), np.array(expected, f32))
    np.testing.assert_allclose(_conv_scipy(x, w, stride, "SAME").reshape(-1), expected, rtol=0, atol=0)


@pytest.mark.parametrize("case", TF_CONV_VALID, ids=[c[0] for c in TF_CONV_VALID])
def test_tf_conv2d_valid_vectors_pin_the_independent_implementation(case):
    """The reference only uses SAME; TF's VALID tables pin the scipy cross-check (tap order, HWIO, no kernel flip)."""
    _, xs, ws, stride, expected = case
    np.testing.assert_allclose(_conv_scipy(_seq(xs), _seq(ws), stride, "VALID").reshape(-1), expected, rtol=0, atol=0)


@pytest.mark.parametrize("geom", [(1, 7, 9, 3, 5, 3, 1), (2, 8, 8, 4, 6, 3, 2), (1, 9, 7, 2, 3, 3, 2), (1, 10, 12, 3, 4, 7, 1),
                                  (1, 6, 6, 5, 2, 1, 1), (2, 5, 8, 2, 2, 2, 2), (1, 11, 11, 1, 1, 5, 3)])
def test_conv2d_same_vs_scipy_random(geom):
    """even / odd sizes x stride 1 / 2 / 3 x kernel 1 / 2 / 3 / 5 / 7: the oracle equals the scipy implementation."""
    B, H, W, Cin, Cout, k, s = geom
    rng = np.random.default_rng(sum(geom))
    x = rng.normal(size=(B, H, W, Cin)).astype(f32); w = rng.normal(size=(k, k, Cin, Cout)).astype(f32)
    b = rng.normal(size=Cout).astype(f32)
    ref = _conv_scipy(x, w, s, "SAME") + b.astype(np.float64)
    np.testing.assert_allclose(T.conv2d_same(x, w, b, s, np.float64), ref, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(T.conv2d_same(x, w, b, s), ref, rtol=2e-5, atol=2e-5)


# ------------------------------------------------------------------------------------------ resize_bilinear (legacy, align_corners=False)
# tensorflow/core/kernels/resize_bilinear_op_test.cc (ResizeBilinearOpTest) and
# tensorflow/python/ops/image_ops_test.py (ResizeImagesTest.testResizeUp)
TF_RESIZE = [
    ("TestBilinear2x2To1x1", (2, 2), (1, 1), [1, 2, 3, 4], [1]),
    ("TestBilinear2x2To3x3", (2, 2), (3, 3), [1, 2, 3, 4], [1, 5 / 3, 2, 7 / 3, 3, 10 / 3, 3, 11 / 3, 4]),
    ("TestBilinear3x3To2x2", (3, 3), (2, 2), list(range(1, 10)), [1, 2.5, 5.5, 7]),
    ("TestBilinear3x3To4x4", (3, 3), (4, 4), list(range(1, 10)),
     [1, 1.75, 2.5, 3, 3.25, 4, 4.75, 5.25, 5.5, 6.25, 7, 7.5, 7, 7.75, 8.5, 9]),
    ("TestBilinear4x4To3x3", (4, 4), (3, 3), list(range(1, 17)), [1, 7 / 3, 11 / 3, 19 / 3, 23 / 3, 9, 35 / 3, 13, 43 / 3]),
    ("TestBilinear2x2To4x4", (2, 2), (4, 4), [1, 2, 3, 4], [1, 1.5, 2, 2, 2, 2.5, 3, 3, 3, 3.5, 4, 4, 3, 3.5, 4, 4]),
    ("ResizeImagesTest.testResizeUp", (3, 2), (6, 4), [64, 32, 32, 64, 50, 100],
     [64, 48, 32, 32, 48, 48, 48, 48, 32, 48, 64, 64, 41, 61.5, 82, 82, 50, 75, 100, 100, 50, 75, 100, 100]),
]


def _resize_scalar(x, oh, ow):
    """Scalar transcription of resize_bilinear_op.cc (TF 1.3, align_corners = false): in = out * (in_size / out_size),
    lower = floor, upper = min(lower + 1, size - 1), lerp = in - lower; top / bottom interpolation then vertical."""
    B, H, W, C = x.shape
    hs, ws = f32(H) / f32(oh), f32(W) / f32(ow)
    y = np.zeros((B, oh, ow, C), f32)
    for oy in range(oh):
        iy = f32(oy) * hs; y0 = int(math.floor(iy)); y1 = min(y0 + 1, H - 1); ly = f32(iy - f32(y0))
        for ox in range(ow):
            ix = f32(ox) * ws; x0 = int(math.floor(ix)); x1 = min(x0 + 1, W - 1); lx = f32(ix - f32(x0))
            top = x[:, y0, x0] + (x[:, y0, x1] - x[:, y0, x0]) * lx
            bot = x[:, y1, x0] + (x[:, y1, x1] - x[:, y1, x0]) * lx
            y[:, oy, ox] = top + (bot - top) * ly
    return y


@pytest.mark.parametrize("case", TF_RESIZE, ids=[c[0] for c in TF_RESIZE])
def test_tf_resize_bilinear_vectors(case):
    _, ishape, oshape, data, expected = case
    x = np.array(data, f32).reshape(1, ishape[0], ishape[1], 1)
    y = T.resize_bilinear_tf1(x, *oshape)
    np.testing.assert_allclose(y.reshape(-1), np.array(expected, np.float64), rtol=1e-6, atol=1e-6)
    np.testing.assert_array_equal(y, _resize_scalar(x, *oshape))


def test_tf_resize_bilinear_2x2x2_to_3x3x2():
    """ResizeBilinearOpTest.TestBilinear2x2x2To3x3x2: two channels, the second the negation of the first."""
    x = np.array([1, -1, 2, -2, 3, -3, 4, -4], f32).reshape(1, 2, 2, 2)
    e = np.array([1, 5 / 3, 2, 7 / 3, 3, 10 / 3, 3, 11 / 3, 4])
    y = T.resize_bilinear_tf1(x, 3, 3)
    np.testing.assert_allclose(y[0, :, :, 0].reshape(-1), e, rtol=1e-6)
    np.testing.assert_allclose(y[0, :, :, 1].reshape(-1), -e, rtol=1e-6)


@pytest.mark.parametrize("geom", [(40, 40, 320, 320, 2), (32, 32, 256, 256, 21), (30, 40, 240, 320, 2), (5, 7, 13, 9, 3)])
def test_resize_bilinear_vs_scalar_transcription(geom):
    """the reference's x8 up-sampling shapes (nets/ColorHandPose3DNetwork.py:97,128,166) and an odd down-sampling one"""
    H, W, oh, ow, C = geom
    x = np.random.default_rng(3).normal(size=(1, H, W, C)).astype(f32)
    np.testing.assert_array_equal(T.resize_bilinear_tf1(x, oh, ow), _resize_scalar(x, oh, ow))


# ------------------------------------------------------------------------------------------ crop_and_resize
# tensorflow/core/kernels/crop_and_resize_op_test.cc (CropAndResizeOpTest)
TF_CROP = [
    # name, image shape (H, W), image, boxes, crop size, extrapolation value, expected
    ("TestCropAndResize2x2To1x1", (2, 2), [1, 2, 3, 4], [[0, 0, 1, 1]], (1, 1), 0.0, [2.5]),
    ("TestCropAndResize2x2To1x1Flipped", (2, 2), [1, 2, 3, 4], [[1, 1, 0, 0]], (1, 1), 0.0, [2.5]),
    ("TestCropAndResize2x2To3x3", (2, 2), [1, 2, 3, 4], [[0, 0, 1, 1]], (3, 3), 0.0, [1, 1.5, 2, 2, 2.5, 3, 3, 3.5, 4]),
    ("TestCropAndResize2x2To3x3Flipped", (2, 2), [1, 2, 3, 4], [[1, 1, 0, 0]], (3, 3), 0.0, [4, 3.5, 3, 3, 2.5, 2, 2, 1.5, 1]),
    ("TestCropAndResize3x3To2x2", (3, 3), list(range(1, 10)), [[0, 0, 1, 1], [0, 0, 0.5, 0.5]], (2, 2), 0.0,
     [1, 3, 7, 9, 1, 2, 4, 5]),
    ("TestCropAndResize3x3To2x2Flipped", (3, 3), list(range(1, 10)), [[1, 1, 0, 0], [0.5, 0.5, 0, 0]], (2, 2), 0.0,
     [9, 7, 3, 1, 5, 4, 2, 1]),
    ("TestCropAndResize2x2To3x3Extrapolated", (2, 2), [1, 2, 3, 4], [[-1, -1, 1, 1]], (3, 3), -1.0,
     [-1, -1, -1, -1, 1, 2, -1, 3, 4]),
    # the same box with the reference's extrapolation value 0 (tf.image.crop_and_resize default, utils/general.py:195)
    ("Extrapolated_default0", (2, 2), [1, 2, 3, 4], [[-1, -1, 1, 1]], (3, 3), 0.0, [0, 0, 0, 0, 1, 2, 0, 3, 4]),
]


def _crop_scalar(image, boxes, ch, cw, extrapolation):
    """Scalar transcription of the CPU functor in crop_and_resize_op.cc (TF 1.3), box_ind = range(B)."""
    B, H, W, C = image.shape
    out = np.zeros((B, ch, cw, C), f32)
    for b in range(B):
        y1, x1, y2, x2 = [f32(v) for v in boxes[b]]
        hscale = (y2 - y1) * f32(H - 1) / f32(ch - 1) if ch > 1 else f32(0)
        wscale = (x2 - x1) * f32(W - 1) / f32(cw - 1) if cw > 1 else f32(0)
        for y in range(ch):
            in_y = f32(y1 * f32(H - 1) + f32(y) * hscale) if ch > 1 else f32(f32(0.5) * (y1 + y2) * f32(H - 1))
            if in_y < 0 or in_y > H - 1:
                out[b, y] = extrapolation
                continue
            top, bot = int(math.floor(in_y)), int(math.ceil(in_y))
            ylerp = f32(in_y - f32(top))
            for x in range(cw):
                in_x = f32(x1 * f32(W - 1) + f32(x) * wscale) if cw > 1 else f32(f32(0.5) * (x1 + x2) * f32(W - 1))
                if in_x < 0 or in_x > W - 1:
                    out[b, y, x] = extrapolation
                    continue
                lef, rig = int(math.floor(in_x)), int(math.ceil(in_x))
                xlerp = f32(in_x - f32(lef))
                t = image[b, top, lef] + (image[b, top, rig] - image[b, top, lef]) * xlerp
                bo = image[b, bot, lef] + (image[b, bot, rig] - image[b, bot, lef]) * xlerp
                out[b, y, x] = t + (bo - t) * ylerp
    return out


@pytest.mark.parametrize("case", TF_CROP, ids=[c[0] for c in TF_CROP])
def test_tf_crop_and_resize_vectors(case):
    _, (H, W), data, boxes, (ch, cw), extrap, expected = case
    nb = len(boxes)
    img = np.repeat(np.array(data, f32).reshape(1, H, W, 1), nb, axis=0)     # TF's box_ind = [0, 0] -> one image copy per box
    y = T.crop_and_resize(img, np.array(boxes, f32), ch, cw, extrapolation_value=extrap)
    np.testing.assert_array_equal(y.reshape(-1), np.array(expected, f32))
    np.testing.assert_array_equal(y, _crop_scalar(img, boxes, ch, cw, f32(extrap)))


def test_crop_and_resize_vs_scalar_transcription_random_boxes():
    """boxes as crop_image_from_xy builds them (utils/general.py:181-191), partly outside the image"""
    rng = np.random.default_rng(5)
    img = rng.normal(size=(6, 20, 24, 3)).astype(f32)
    boxes = np.stack([rng.uniform(-0.3, 0.5, 6), rng.uniform(-0.3, 0.5, 6), rng.uniform(0.5, 1.3, 6), rng.uniform(0.5, 1.3, 6)], -1).astype(f32)
    np.testing.assert_array_equal(T.crop_and_resize(img, boxes, 16, 16), _crop_scalar(img, boxes, 16, 16, f32(0)))


# ------------------------------------------------------------------------------------------ dilation2d
# tensorflow/python/kernel_tests/morphological_ops_test.py (DilationTest)
TF_DILATION = [
    # name, image [B,H,W,C], filter [h,w,C], strides, rates, padding, expected
    ("_testDilationValidPadding", [[[[.1], [.2]], [[.3], [.4]]]], [[[.4], [.3]], [[.1], [.0]]], (1, 1), (1, 1), "VALID", [[[[.5]]]]),
    ("_testDilationSamePadding", [[[[.1], [.2]], [[.3], [.4]]]], [[[.4], [.3]], [[.1], [.0]]], (1, 1), (1, 1), "SAME",
     [[[[.5], [.6]], [[.7], [.8]]]]),
    ("_testDilationSamePaddingDepth", [[[[.1, .2, .0], [.2, .3, .1]], [[.3, .4, .2], [.4, .5, .3]]]],
     [[[.4, .5, .3], [.3, .4, .2]], [[.1, .2, .0], [.0, .1, -.1]]], (1, 1), (1, 1), "SAME",
     [[[[.5, .7, .3], [.6, .8, .4]], [[.7, .9, .5], [.8, 1., .6]]]]),
    ("_testDilationSamePaddingBatch", [[[[.1], [.2]], [[.3], [.4]]], [[[.2], [.3]], [[.4], [.5]]]],
     [[[.4], [.3]], [[.1], [.0]]], (1, 1), (1, 1), "SAME", [[[[.5], [.6]], [[.7], [.8]]], [[[.6], [.7]], [[.8], [.9]]]]),
    ("_testDilationValidPaddingNonSquareWindow", [[[[.1], [.2]], [[.3], [.4]]]], [[[.4], [.3]]], (1, 1), (1, 1), "VALID",
     [[[[.5]], [[.7]]]]),
    ("_testDilationSamePaddingRate", [[[[.1], [.2], [.3]], [[.4], [.5], [.6]], [[.7], [.8], [.9]]]],
     [[[.4], [.3]], [[.1], [.2]]], (1, 1), (2, 2), "SAME", [[[[.7], [.8], [.6]], [[1.0], [1.1], [.9]], [[.8], [.9], [.9]]]]),
    ("_testDilationValidPaddingUnevenStride", [[[[.1], [.2], [.3], [.4]], [[.5], [.6], [.7], [.8]], [[.9], [1.0], [1.1], [1.2]]]],
     [[[.4], [.3]], [[.1], [.2]]], (1, 2), (1, 1), "VALID", [[[[.8], [1.0]], [[1.2], [1.4]]]]),
]


@pytest.mark.parametrize("case", TF_DILATION, ids=[c[0] for c in TF_DILATION])
def test_tf_dilation2d_vectors(case):
    _, img, filt, strides, rates, padding, expected = case
    y = T.dilation2d(np.array(img, f32), np.array(filt, f32), strides, rates, padding)
    np.testing.assert_allclose(y, np.array(expected, f32), rtol=1e-6, atol=1e-7)


def test_dilation2d_21_equals_general_and_scipy():
    """utils/general.py:249,259: the fast 21x21 / 441 form used by the oracle == the general restatement == scipy's grey
    dilation (flat structuring element, out-of-bounds taps ignored = constant -inf border)."""
    from scipy.ndimage import grey_dilation
    rng = np.random.default_rng(9)
    for H, W in [(40, 50), (24, 31), (64, 64)]:
        obj = (rng.uniform(size=(H, W)) > 0.97).astype(f32)
        fast = T.dilation2d_21(obj)
        gen = T.dilation2d(obj.reshape(1, H, W, 1), np.full((21, 21, 1), f32(1.0) / f32(441.0), f32))[0, :, :, 0]
        np.testing.assert_array_equal(fast, gen)
        sp = grey_dilation(obj.astype(np.float64), size=(21, 21), mode="constant", cval=-np.inf) + 1.0 / 441.0
        np.testing.assert_allclose(fast, sp, rtol=1e-6)


def test_dilation2d_general_vs_scipy_nonflat():
    """non-flat structuring element: TF correlates (no flip); scipy's grey_dilation flips the structure, origin centred"""
    from scipy.ndimage import grey_dilation
    rng = np.random.default_rng(10)
    x = rng.normal(size=(1, 9, 11, 1)).astype(f32)
    filt = rng.normal(size=(3, 5, 1)).astype(f32)
    ours = T.dilation2d(x, filt)[0, :, :, 0]
    sp = grey_dilation(x[0, :, :, 0].astype(np.float64), structure=filt[::-1, ::-1, 0].astype(np.float64), mode="constant", cval=-np.inf)
    np.testing.assert_allclose(ours, sp, rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------------------------------------ softmax / round
def test_softmax_vs_scipy_and_round_half_even():
    from scipy.special import softmax
    x = np.random.default_rng(11).normal(scale=4.0, size=(1000, 2)).astype(f32)
    np.testing.assert_allclose(T.softmax_last(x), softmax(x.astype(np.float64), axis=-1), rtol=3e-7, atol=1e-7)
    # tf.round documents round-half-to-even ("Rounds half to even. Also known as bankers rounding."): its doc example
    np.testing.assert_array_equal(T.round_half_even(np.array([0.9, 2.5, 2.3, 1.5, -4.5], f32)), [1.0, 2.0, 2.0, 2.0, -4.0])


def test_tf_max_pool_valid_vector():
    """tensorflow/python/kernel_tests/pooling_ops_test.py (_testMaxPoolValidPadding): input 1..27 as [1,3,3,3], 2x2 window,
    stride 2, VALID -> [13, 14, 15]  (NetworkOps.max_pool, utils/general.py:62-65)."""
    x = _seq((1, 3, 3, 3))
    np.testing.assert_array_equal(T.max_pool_2x2(x).reshape(-1), [13, 14, 15])
