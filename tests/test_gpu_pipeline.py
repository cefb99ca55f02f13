"""GPU parity of the network stages and of the full pipeline against the CPU oracle (reference API level).

Tolerances follow BASELINE.json: key-point indices bit-exact, fp32 score maps / 3-D coordinates within
1e-3 abs for the fp32-parity modes (fp32_ffma, bf16x3, fp16x3) and 1e-2 for the single-pass fp16 mode."""
import numpy as np
import pytest
import torch

from hand3d_b200 import weights as Wt
from oracle import hand3d_oracle as O
from oracle import tf1_ops as T

pytestmark = pytest.mark.gpu
f32 = np.float32
PARITY_MODES = ["fp32_ffma", "bf16x3", "fp16x3"]
TOL = {"fp32_ffma": 1e-3, "bf16x3": 1e-3, "fp16x3": 1e-3, "fp16": 1e-2}


@pytest.fixture(scope="module")
def wd():
    return Wt.synthetic_weights(0)


@pytest.fixture(scope="module")
def net(wd):
    from hand3d_b200.nets.ColorHandPose3DNetwork import ColorHandPose3DNetwork
    n = ColorHandPose3DNetwork()
    n.init(None, weights=wd)
    return n


@pytest.fixture(scope="module")
def ctx(net):
    from hand3d_b200 import runtime
    return runtime.default_context()


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def seg_ref(wd):
    img = np.concatenate([Wt.synthetic_images(1, 320, 320, seed=1), Wt.synthetic_blob_images(1, 320, 320, seed=5)], 0)
    return img, O.inference_detection(img, wd)[-1]


@pytest.mark.parametrize("prec", PARITY_MODES + ["fp16"])
def test_handsegnet_stage(net, ctx, seg_ref, prec):
    img, ref = seg_ref
    ctx.set_precision(prec)
    out = net.inference_detection(_dev(img))
    assert isinstance(out, list) and len(out) == 1 and tuple(out[0].shape) == (2, 320, 320, 2)
    err = np.abs(out[0].cpu().numpy() - ref).max()
    assert err < TOL[prec], "HandSegNet %s: max abs err %.3e" % (prec, err)


def test_handsegnet_240x320(net, ctx, wd):
    img = Wt.synthetic_images(1, 240, 320, seed=7)
    ctx.set_precision("bf16x3")
    out = net.inference_detection(_dev(img))[0].cpu().numpy()
    ref = O.inference_detection(img, wd)[-1]
    assert out.shape == (1, 240, 320, 2)
    assert np.abs(out - ref).max() < 1e-3


@pytest.fixture(scope="module")
def pose_ref(wd):
    crop = Wt.synthetic_images(2, 256, 256, seed=11)
    return crop, O.inference_pose2d(crop, wd)


@pytest.mark.parametrize("prec", PARITY_MODES + ["fp16"])
def test_posenet_stage(net, ctx, pose_ref, prec):
    crop, ref = pose_ref
    ctx.set_precision(prec)
    outs = net.inference_pose2d(_dev(crop))
    assert len(outs) == 3
    for i in range(3):
        err = np.abs(outs[i].cpu().numpy() - ref[i]).max()
        assert err < TOL[prec], "PoseNet %s stage %d: max abs err %.3e" % (prec, i, err)


@pytest.mark.parametrize("variant", ["proposed", "direct"])
def test_lifting_stage(ctx, wd, variant):
    rng = np.random.default_rng(13)
    B = 5
    sm = rng.normal(size=(B, 32, 32, 21)).astype(f32)
    hs = Wt.synthetic_hand_side(B, seed=3)
    out, can, rot = ctx.lifting(_dev(sm), _dev(hs), variant)
    if variant == "proposed":
        r_out, r_can, r_R = O.inference_pose3d(sm, hs, wd)
        np.testing.assert_allclose(rot.cpu().numpy(), r_R, atol=1e-4)
    else:
        r_can = O.inference_pose3d_can(sm, hs, wd)
        r_out = r_can
    np.testing.assert_allclose(can.cpu().numpy(), r_can, atol=1e-4)
    np.testing.assert_allclose(out.cpu().numpy(), r_out, atol=1e-4)


def test_pose_prior_network_variants(wd):
    from hand3d_b200 import runtime
    from hand3d_b200.nets.PosePriorNetwork import PosePriorNetwork
    rng = np.random.default_rng(14)
    B = 3
    sm = rng.normal(size=(B, 256, 256, 21)).astype(f32)
    hs = Wt.synthetic_hand_side(B, seed=4)
    for variant in ("proposed", "direct"):
        p = PosePriorNetwork(variant)
        p.init(None, weights={k: v for k, v in wd.items() if k.startswith(("PosePrior", "ViewpointNet"))})
        out, c3, R = p.inference(_dev(sm), _dev(hs), True)
        ref = O.pose_prior_inference(sm, hs, wd, variant)
        np.testing.assert_allclose(out.cpu().numpy(), ref[0], atol=1e-4)
        np.testing.assert_allclose(c3.cpu().numpy(), ref[1], atol=1e-4)
        assert (R is None) == (ref[2] is None)
    wb = Wt.synthetic_weights(0, bottleneck=True)
    p = PosePriorNetwork("bottleneck")
    p.init(None, weights={k: v for k, v in wb.items() if k.startswith("PosePrior")})
    out, _, _ = p.inference(_dev(sm), _dev(hs), True)
    np.testing.assert_allclose(out.cpu().numpy(), O.pose_prior_inference(sm, hs, wb, "bottleneck")[0], atol=1e-4)
    # restore the standard weights for the tests that follow
    runtime.default_context().load_weights({k: v for k, v in wd.items() if k.startswith("PosePrior")})
    with pytest.raises(NotImplementedError):
        PosePriorNetwork("local").inference(_dev(sm), _dev(hs), True)


def _pipeline_case(kind):
    if kind == "noise":
        return Wt.synthetic_images(3, 320, 320, seed=1), Wt.synthetic_weights(0)
    return Wt.synthetic_blob_images(3, 320, 320, seed=5), Wt.synthetic_weights(0, seg_shift=0.15)


@pytest.mark.parametrize("kind", ["noise", "blobs"])
@pytest.mark.parametrize("prec", PARITY_MODES)
def test_full_pipeline(kind, prec):
    """inference(): every output against the oracle; crop parameters and key-point indices bit-exact."""
    from hand3d_b200 import runtime
    from hand3d_b200.nets.ColorHandPose3DNetwork import ColorHandPose3DNetwork
    from hand3d_b200.utils.general import detect_keypoints, trafo_coords
    img, w = _pipeline_case(kind)
    hs = Wt.synthetic_hand_side(3, seed=2)
    net = ColorHandPose3DNetwork()
    net.init(None, weights=w)
    ctx = runtime.default_context()
    ctx.set_precision(prec)
    out = net.inference(_dev(img), _dev(hs), True)
    g = [o.cpu().numpy() for o in out]
    ref = O.inference(img, hs, w, literal_mask=False)
    assert np.abs(g[0] - ref[0]).max() < 1e-3                       # hand_scoremap
    np.testing.assert_array_equal(g[3], ref[3])                      # center (discrete decision: exact)
    np.testing.assert_array_equal(g[2], ref[2])                      # scale_crop
    np.testing.assert_array_equal(g[1], ref[1])                      # image_crop: same fp32 op order -> bit-exact
    assert np.abs(g[4] - ref[4]).max() < 1e-3                       # keypoints_scoremap
    assert np.abs(g[5] - ref[5]).max() < 1e-3                       # keypoint_coord3d
    uv = net.last_keypoints_uv.cpu().numpy()
    for b in range(3):
        kp_ref = O.detect_keypoints(ref[4][b])
        np.testing.assert_array_equal(uv[b], kp_ref.astype(np.int32))
        np.testing.assert_array_equal(detect_keypoints(g[4][b]), O.detect_keypoints(g[4][b]))
        np.testing.assert_allclose(trafo_coords(detect_keypoints(g[4][b]), g[3][b:b + 1], g[2][b:b + 1], 256),
                                   O.trafo_coords(kp_ref, ref[3][b:b + 1], ref[2][b:b + 1], 256))
    if kind == "blobs":
        assert len(np.unique(g[2])) > 1, "the blob set is meant to give varied crops"


def test_inference2d_tuple_order_and_teacher_forcing(wd):
    from hand3d_b200 import runtime
    from hand3d_b200.nets.ColorHandPose3DNetwork import ColorHandPose3DNetwork
    img = Wt.synthetic_images(2, 240, 320, seed=9)
    net = ColorHandPose3DNetwork()
    net.init(None, weights=wd, exclude_var_list=['PosePrior', 'ViewpointNet'])     # eval2d.py:78-79
    ctx = runtime.default_context()
    ctx.set_precision("bf16x3")
    kps, crop, scale, center = net.inference2d(_dev(img))
    r = O.inference2d(img, wd, literal_mask=False)
    assert tuple(kps.shape) == (2, 256, 256, 21) and tuple(crop.shape) == (2, 256, 256, 3)
    np.testing.assert_array_equal(center.cpu().numpy(), r[3])
    np.testing.assert_array_equal(scale.cpu().numpy(), r[2])
    assert np.abs(kps.cpu().numpy() - r[0]).max() < 1e-3
    # teacher-forced crop parameters
    fc = np.array([[100.0, 120.0], [200.0, 50.0]], f32); fs = np.array([[1.5], [0.7]], f32)
    hs = Wt.synthetic_hand_side(2, seed=1)
    res = ctx.pipeline(_dev(img), _dev(hs), True, force_center=_dev(fc), force_scale=_dev(fs))
    ref = O.inference(img, hs, wd, forced_crop=(fc, fs))
    np.testing.assert_array_equal(res["image_crop"].cpu().numpy(), ref[1])
    assert np.abs(res["keypoint_coord3d"].cpu().numpy() - ref[5]).max() < 1e-3


def test_unknown_variable_and_bad_numerics(ctx):
    with pytest.raises(ValueError):
        ctx.load_weights({"HandSegNet/conv9_9/weights": np.zeros((3, 3, 3, 3), f32)})
    with pytest.raises(ValueError):
        ctx.load_weights({"HandSegNet/conv1_1/weights": np.zeros((3, 3, 4, 64), f32)})
    bad = np.zeros((512, 63), f32); bad[0, 0] = np.nan
    with pytest.raises(ValueError):
        ctx.load_weights({"PosePrior/fc_xyz/weights": bad})


def test_fp16_fast_mode_tolerance(wd):
    """BASELINE config 5: single-pass fp16 tensor-core path within 1e-2."""
    from hand3d_b200 import runtime
    img = Wt.synthetic_images(2, 320, 320, seed=1)
    hs = Wt.synthetic_hand_side(2, seed=2)
    ctx = runtime.default_context()
    ctx.load_weights(wd)
    ctx.set_precision("bf16x3")
    base = ctx.pipeline(_dev(img), _dev(hs), True)
    ctx.set_precision("fp16")
    fast = ctx.pipeline(_dev(img), _dev(hs), True, force_center=base["center"], force_scale=base["scale_crop"])
    assert (fast["keypoints_scoremap"] - base["keypoints_scoremap"]).abs().max().item() < 1e-2
    assert (fast["keypoint_coord3d"] - base["keypoint_coord3d"]).abs().max().item() < 1e-2
    ctx.set_precision("bf16x3")
