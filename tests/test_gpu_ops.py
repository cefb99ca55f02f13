"""GPU parity of the HBM-bound / CUDA-core operators against the oracle (through the C ABI)."""
import numpy as np
import pytest
import torch

from oracle import hand3d_oracle as O
from oracle import tf1_ops as T

pytestmark = pytest.mark.gpu
f32 = np.float32


@pytest.fixture(scope="module")
def ctx():
    from hand3d_b200 import runtime
    return runtime.default_context()


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _smooth_logits(rng, B, H, W, amp=3.0, bias=-1.0):
    low = rng.normal(size=(B, H // 8, W // 8, 2)).astype(f32) * amp
    low[..., 1] += bias
    return T.resize_bilinear_tf1(low, H, W)


@pytest.mark.parametrize("shape,out", [((2, 40, 40, 2), (320, 320)), ((2, 32, 32, 21), (256, 256)), ((1, 30, 40, 2), (240, 320)),
                                       ((1, 12, 10, 3), (30, 17)), ((1, 48, 64, 5), (24, 32)), ((1, 8, 8, 1), (8, 8))])
def test_resize_bilinear_bit_exact(ctx, shape, out):
    x = np.random.default_rng(0).normal(size=shape).astype(f32)
    y = ctx.resize_bilinear(_dev(x), *out).cpu().numpy()
    np.testing.assert_array_equal(y, T.resize_bilinear_tf1(x, *out))


@pytest.mark.parametrize("C", [64, 21])
def test_maxpool(ctx, C):
    x = np.random.default_rng(1).normal(size=(2, 16, 24, C)).astype(f32)
    np.testing.assert_array_equal(ctx.max_pool(_dev(x)).cpu().numpy(), T.max_pool_2x2(x))


def test_avgpool8(ctx):
    x = np.random.default_rng(2).normal(size=(2, 256, 256, 21)).astype(f32)
    np.testing.assert_allclose(ctx.avg_pool8(_dev(x)).cpu().numpy(), T.avg_pool_8x8(x), atol=1e-6)


@pytest.mark.parametrize("H,W,seed", [(320, 320, 0), (240, 320, 1), (320, 320, 2), (64, 96, 3), (320, 320, 4)])
def test_seg_postprocess_matches_oracle(ctx, H, W, seed):
    rng = np.random.default_rng(seed)
    B = 4
    sm = _smooth_logits(rng, B, H, W, amp=2.0 + seed, bias=-1.5)
    if seed == 4:
        sm[1, ..., 1] = -50.0          # empty mask -> (160,160)/100 fallbacks
        sm[2, ..., 1] = 50.0           # saturated soft-max: seed = first pixel, mask = everything
    r = ctx.seg_postprocess(_dev(sm))
    fg, det = O.seg_fg_det(sm)
    mask = O.single_obj_scoremap(sm, literal=False)
    center, _, size = O.calc_center_bb(mask)
    np.testing.assert_array_equal(r["max_loc"].cpu().numpy(), O.find_max_location(fg))
    np.testing.assert_array_equal(r["hand_mask"].cpu().numpy(), mask[..., 0].astype(np.uint8))
    np.testing.assert_array_equal(r["center"].cpu().numpy(), center)
    np.testing.assert_array_equal(r["crop_size"].cpu().numpy(), size)
    np.testing.assert_array_equal(r["scale_crop"].cpu().numpy(), O.crop_scale(size))


def test_seg_postprocess_needs_all_32_passes(ctx):
    H = W = 320
    sm = np.zeros((1, H, W, 2), f32); sm[..., 0] = 5.0
    sm[0, 7, :, 1] = 10.0; sm[0, 7, 0, 1] = 11.0           # 1-px line: grows 10 px per pass
    sm[0, 7:300, 319, 1] = 10.0                             # continues down the right edge: not fully reachable in 32 passes
    r = ctx.seg_postprocess(_dev(sm))
    mask = O.single_obj_scoremap(sm, literal=False)
    np.testing.assert_array_equal(r["hand_mask"].cpu().numpy(), mask[..., 0].astype(np.uint8))
    assert 0 < mask.sum() < (sm[..., 1] > 5).sum()          # the pass limit really truncated the growth


@pytest.mark.parametrize("H,W", [(320, 320), (240, 320)])
def test_crop_image_from_xy(ctx, H, W):
    rng = np.random.default_rng(5)
    B = 6
    img = rng.uniform(-0.5, 0.5, size=(B, H, W, 3)).astype(f32)
    center = np.stack([rng.uniform(-20, H + 20, B), rng.uniform(-20, W + 20, B)], 1).astype(f32)
    scale = np.array([0.25, 0.64, 1.0, 2.048, 5.0, 1.7], f32).reshape(B, 1)
    out = ctx.crop_image_from_xy(_dev(img), _dev(center), 256, _dev(scale)).cpu().numpy()
    ref = O.crop_image_from_xy(img, center, 256, scale)
    np.testing.assert_array_equal(out, ref)


def test_detect_keypoints_first_occurrence(ctx):
    rng = np.random.default_rng(6)
    s = rng.normal(size=(3, 256, 256, 21)).astype(f32)
    s[0, 10, 20, 3] = 9.0; s[0, 200, 5, 3] = 9.0          # duplicate maximum: the first one wins
    s[1, :, :, 7] = -1.0                                   # constant map -> (0,0)
    uv = ctx.detect_keypoints(_dev(s)).cpu().numpy()
    for b in range(3):
        np.testing.assert_array_equal(uv[b], O.detect_keypoints(s[b]).astype(np.int32))
    from hand3d_b200.utils.general import detect_keypoints
    kp = detect_keypoints(s[2])
    assert kp.dtype == np.float64
    np.testing.assert_array_equal(kp, O.detect_keypoints(s[2]))


CONV_CASES = [  # B,H,W,Cin,Cout,k,stride
    (2, 32, 32, 3, 64, 3, 1), (1, 17, 23, 21, 32, 3, 1), (2, 32, 32, 32, 32, 3, 2), (1, 16, 16, 64, 64, 3, 2),
    (1, 8, 8, 128, 256, 3, 1), (2, 16, 16, 512, 2, 1, 1), (1, 16, 16, 128, 21, 1, 1), (1, 12, 12, 149, 128, 7, 1),
    (1, 9, 7, 16, 8, 3, 2), (1, 40, 40, 64, 128, 3, 1), (2, 19, 45, 3, 64, 3, 1), (3, 4, 4, 256, 256, 3, 1), (2, 8, 8, 128, 128, 3, 2),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_f32_vs_oracle(ctx, case):
    B, H, W, Cin, Cout, k, s = case
    rng = np.random.default_rng(7)
    x = rng.normal(size=(B, H, W, Cin)).astype(f32)
    w = (rng.normal(size=(k, k, Cin, Cout)) / np.sqrt(k * k * Cin)).astype(f32)
    b = rng.normal(size=Cout).astype(f32)
    for leaky in (False, True):
        y = ctx.conv2d(_dev(x), _dev(w), _dev(b), stride=s, leaky=leaky).cpu().numpy()
        ref = T.conv2d_same(x, w, b, s, np.float64)
        if leaky:
            ref = T.leaky_relu(ref)
        np.testing.assert_allclose(y, ref, atol=2e-5, rtol=1e-5)


def test_fully_connected(ctx):
    rng = np.random.default_rng(8)
    for B, i, o in [(3, 2050, 512), (33, 512, 63), (5, 128, 3), (1, 4098, 256)]:
        x = rng.normal(size=(B, i)).astype(f32); w = (rng.normal(size=(i, o)) / np.sqrt(i)).astype(f32); b = rng.normal(size=o).astype(f32)
        y = ctx.fully_connected(_dev(x), _dev(w), _dev(b), leaky=True).cpu().numpy()
        ref = T.fully_connected(x, w, b, np.float64)
        np.testing.assert_allclose(y, np.maximum(ref, 0.01 * ref), atol=2e-5, rtol=1e-5)


def test_rotate_canonical(ctx):
    rng = np.random.default_rng(9)
    B = 7
    can = rng.normal(size=(B, 21, 3)).astype(f32)
    u = rng.normal(size=(B, 3)).astype(f32); u[0] = 0.0
    hs = np.zeros((B, 2), f32); hs[np.arange(B), rng.integers(0, 2, B)] = 1
    rot, out = ctx.rotate_canonical(_dev(can), _dev(u), _dev(hs))
    R = O.get_rot_mat(u[:, 0:1], u[:, 1:2], u[:, 2:3])
    np.testing.assert_allclose(rot.cpu().numpy(), R, atol=2e-6)
    np.testing.assert_allclose(out.cpu().numpy(), np.matmul(O.flip_right_hand(can, hs), R), atol=1e-5)


def test_bone_rel_trafo_inv(ctx):
    """Forward kinematics kernel vs the oracle, and the round trip xyz -> bone_rel_trafo (oracle) -> device -> xyz."""
    rng = np.random.default_rng(15)
    rel = np.stack([rng.uniform(0.1, 1.5, (9, 21)), rng.uniform(-3, 3, (9, 21)), rng.uniform(-3, 3, (9, 21))], -1).astype(f32)
    out = ctx.bone_rel_trafo_inv(_dev(rel)).cpu().numpy()
    np.testing.assert_allclose(out, O.bone_rel_trafo_inv(rel), atol=2e-5)
    xyz = rng.normal(size=(4, 21, 3)).astype(f32)
    back = ctx.bone_rel_trafo_inv(_dev(O.bone_rel_trafo(xyz))).cpu().numpy()
    np.testing.assert_allclose(back, xyz, atol=2e-5)
    from hand3d_b200.utils.relative_trafo import bone_rel_trafo_inv
    np.testing.assert_allclose(bone_rel_trafo_inv(_dev(rel[0])).cpu().numpy()[0], out[0], atol=0)


def test_network_ops_mirror(ctx):
    """NetworkOps.conv_relu / max_pool / fully_connected_relu with tf-style variable scopes."""
    from hand3d_b200 import weights as Wt
    from hand3d_b200.utils.general import NetworkOps as ops, variable_scope
    wd = Wt.synthetic_weights(0)
    ctx.load_weights({k: v for k, v in wd.items() if k.startswith("PosePrior/")})
    x = np.random.default_rng(10).normal(size=(2, 32, 32, 21)).astype(f32)
    with variable_scope("PosePrior"):
        y = ops.conv_relu(_dev(x), "conv_pose_0_1", kernel_size=3, stride=1, out_chan=32)
        y = ops.conv_relu(y, "conv_pose_0_2", kernel_size=3, stride=2, out_chan=32)
        p = ops.max_pool(y)
    r = T.conv_relu(x, wd["PosePrior/conv_pose_0_1/weights"], wd["PosePrior/conv_pose_0_1/biases"], 1)
    r = T.conv_relu(r, wd["PosePrior/conv_pose_0_2/weights"], wd["PosePrior/conv_pose_0_2/biases"], 2)
    np.testing.assert_allclose(y.cpu().numpy(), r, atol=1e-5)
    np.testing.assert_allclose(p.cpu().numpy(), T.max_pool_2x2(r), atol=1e-5)


def _fake_rhd_record(rng):
    xyz = rng.normal(size=(42, 3)).astype(f32); uv = rng.uniform(-5, 330, size=(42, 2)).astype(f32); K = rng.normal(size=9).astype(f32)
    img = rng.integers(0, 256, size=(320, 320, 3), dtype=np.uint8); parts = rng.integers(0, 34, size=(320, 320), dtype=np.uint8)
    vis = rng.integers(0, 2, size=42, dtype=np.uint8)
    rec = xyz.tobytes() + uv.tobytes() + K.tobytes() + bytes([255, 255]) + img.tobytes() + parts.tobytes() + vis.tobytes()
    assert len(rec) == 410520
    return rec


def test_decode_rhd_and_stb_records(ctx):
    """On-device decode of the readers' binary records vs the oracle restatement (bit-exact)."""
    from hand3d_b200.data.records import decode_rhd_records, decode_stb_records
    rng = np.random.default_rng(16)
    recs = [_fake_rhd_record(rng) for _ in range(3)]
    out = decode_rhd_records(b"".join(recs))
    for b, r in enumerate(recs):
        ref = O.decode_rhd_record(r)
        for k in ref:
            np.testing.assert_array_equal(out[k][b].cpu().numpy(), ref[k], err_msg=k)
    xyz = rng.normal(size=(21, 3)).astype(f32); uvv = np.concatenate([rng.uniform(0, 640, (21, 2)), rng.integers(0, 2, (21, 1))], 1).astype(f32)
    img = rng.integers(0, 256, size=(480, 640, 3), dtype=np.uint8)
    rec = xyz.tobytes() + uvv.tobytes() + img.tobytes()
    for sub in (1, 2):
        out = decode_stb_records(rec + rec, subsample=sub)
        ref = O.decode_stb_record(rec, subsample=sub)
        for k in ref:
            np.testing.assert_array_equal(out[k][1].cpu().numpy(), ref[k], err_msg=k)


def test_eval_util_device_matches_reference_semantics(ctx):
    from hand3d_b200.utils.general import EvalUtil
    rng = np.random.default_rng(17)
    B = 40
    gt = rng.normal(size=(B, 21, 2)).astype(f32) * 20; pred = gt + rng.normal(size=(B, 21, 2)).astype(f32) * 6
    vis = rng.integers(0, 2, size=(B, 21)).astype(bool); vis[:, 5] = False
    ours, ours_np, ref = EvalUtil(), EvalUtil(), O.EvalUtil()
    ours.feed(_dev(gt), _dev(vis), _dev(pred))                        # one batched device call
    for b in range(B):
        ours_np.feed(gt[b], vis[b], pred[b]); ref.feed(gt[b], vis[b], pred[b])
    for a in (ours, ours_np):
        got, exp = a.get_measures(0.0, 30.0, 20), ref.get_measures(0.0, 30.0, 20)
        for x, y in zip(got, exp):
            np.testing.assert_allclose(x, y, rtol=1e-6)


@pytest.mark.parametrize("H,W,seed", [(100, 70, 7), (512, 512, 8), (41, 500, 9)])
def test_seg_postprocess_odd_sizes(ctx, H, W, seed):
    """widths that are not multiples of 32, the 512 x 512 limit, a map wider than tall: the log-step / van Herk dilation handles
    partial words, partial 21-row blocks and the image borders like the literal 21 x 21 dilation"""
    rng = np.random.default_rng(seed)
    sm = _smooth_logits(rng, 2, H, W, amp=3.0, bias=-1.0)
    r = ctx.seg_postprocess(_dev(sm))
    mask = O.single_obj_scoremap(sm, literal=False)
    center, _, size = O.calc_center_bb(mask)
    np.testing.assert_array_equal(r["hand_mask"].cpu().numpy(), mask[..., 0].astype(np.uint8))
    np.testing.assert_array_equal(r["center"].cpu().numpy(), center)
    np.testing.assert_array_equal(r["crop_size"].cpu().numpy(), size)


@pytest.mark.parametrize("shape,out", [((3, 32, 32), (256, 256)), ((2, 30, 40), (240, 320)), ((2, 16, 16), (32, 32)), ((1, 20, 12), (60, 36))])
def test_upsample_detect_keypoints_fused(ctx, shape, out):
    """h3d_upsample_detect_keypoints (the row-group kernel for power-of-two factors, the per-pixel kernel otherwise) against the oracle:
    bit-exact up-sampled map and first-occurrence indices, incl. the plateaus the edge replication creates."""
    rng = np.random.default_rng(21)
    B, H, W = shape
    s = rng.normal(size=(B, H, W, 21)).astype(f32)
    s[0, H - 1, W - 1, 4] = 50.0        # maximum in the replicated corner: a plateau whose first occurrence is its top-left pixel
    s[0, 5, :, 2] = 7.0                 # a whole input row at the maximum
    s[B - 1, :, :, 9] = 0.25            # constant channel -> (0, 0)
    up, uv = ctx.upsample_detect_keypoints(_dev(s), *out)
    ref = T.resize_bilinear_tf1(s, *out)
    np.testing.assert_array_equal(up.cpu().numpy(), ref)
    for b in range(B):
        np.testing.assert_array_equal(uv.cpu().numpy()[b], O.detect_keypoints(ref[b]).astype(np.int32))


def test_calc_center_bb_leaky_relu_flip_pack(ctx):
    """the reference helpers that were eager torch in round 1 now run one kernel each"""
    from hand3d_b200.utils.general import NetworkOps, calc_center_bb
    from hand3d_b200.nets.ColorHandPose3DNetwork import ColorHandPose3DNetwork
    from hand3d_b200.distributed import pack_records
    rng = np.random.default_rng(22)
    mask = (rng.uniform(size=(4, 60, 80, 1)) > 0.995).astype(f32)
    mask[1] = 0.0                                               # empty -> fall-backs
    mask[2] = 0.0
    mask[2, 10:30, 5:60, 0] = 1.0
    mask[3] *= 2.0                                              # values != 1 do not count (tf.equal(mask, 1))
    c, bb, sz = calc_center_bb(_dev(mask))
    rc, rbb, rsz = O.calc_center_bb(mask)
    np.testing.assert_array_equal(c.cpu().numpy(), rc)
    np.testing.assert_array_equal(sz.cpu().numpy(), rsz)
    assert tuple(bb.shape) == (4, 2, 2) and torch.isinf(bb[1]).all()
    np.testing.assert_array_equal(bb[2].cpu().numpy(), [[10, 29], [5, 59]])
    x = rng.normal(size=(3, 7, 5, 13)).astype(f32)
    np.testing.assert_array_equal(NetworkOps.leaky_relu(_dev(x)).cpu().numpy(), T.leaky_relu(x))
    xyz = rng.normal(size=(5, 21, 3)).astype(f32)
    cond = np.array([True, False, True, True, False])
    out = ColorHandPose3DNetwork._flip_right_hand(_dev(xyz), torch.from_numpy(cond).cuda().reshape(5, 1, 1).expand(5, 21, 3)).cpu().numpy()
    ref = xyz.copy(); ref[cond, :, 2] *= -1
    np.testing.assert_array_equal(out, ref)
    uvk = rng.integers(0, 256, size=(5, 21, 2)).astype(np.int32)
    cen = rng.normal(size=(5, 2)).astype(f32); scl = rng.uniform(0.5, 2, size=(5, 1)).astype(f32)
    rec = pack_records(_dev(xyz), torch.from_numpy(uvk).cuda(), _dev(cen), _dev(scl)).cpu()
    ref_rec = pack_records(torch.from_numpy(xyz), torch.from_numpy(uvk), torch.from_numpy(cen), torch.from_numpy(scl))
    assert torch.equal(rec.view(torch.int32), ref_rec.view(torch.int32))


def test_operator_entries_enqueue_only(ctx):
    """include/hand3d_b200.h: operator entries never allocate or synchronise per call.  (1) After one warm-up round the driver's free
    memory does not move across a second round of direct C-ABI calls; (2) the calls can be captured into a CUDA graph (stream capture
    rejects cudaMalloc / cudaFree / synchronisation) and the replay reproduces the eager results."""
    import ctypes as C
    from hand3d_b200 import _lib
    rng = np.random.default_rng(23)
    L = ctx.lib
    logits = _dev(_smooth_logits(rng, 2, 64, 96, amp=2.0, bias=-1.0))
    sm = _dev(rng.normal(size=(2, 32, 32, 21)).astype(f32))
    fx = _dev(rng.normal(size=(4, 300)).astype(f32)); fw = _dev(rng.normal(size=(300, 40)).astype(f32)); fb = _dev(rng.normal(size=40).astype(f32))
    cx = _dev(rng.normal(size=(2, 16, 16, 64)).astype(f32))
    wk = (rng.normal(size=(3, 3, 64, 64)) / 24).astype(f32)
    pk = ctx.pack_conv(wk, np.zeros(64, f32), "bf16x3")
    outs = {"mask": torch.empty((2, 64, 96), dtype=torch.uint8, device="cuda"), "loc": torch.empty((2, 2), dtype=torch.int32, device="cuda"),
            "center": torch.empty((2, 2), device="cuda"), "size": torch.empty((2, 1), device="cuda"), "scale": torch.empty((2, 1), device="cuda"),
            "uv": torch.empty((2, 21, 2), dtype=torch.int32, device="cuda"), "fy": torch.empty((4, 40), device="cuda"),
            "cy": torch.empty((2, 16, 16, 64), device="cuda"), "ty": torch.empty((2, 8, 8, 64), device="cuda")}
    p = lambda t: C.c_void_p(t.data_ptr())     # noqa: E731

    def round_(stream):
        st = C.c_void_p(stream)
        _lib.check(L.h3d_seg_postprocess(ctx.h, p(logits), 2, 64, 96, p(outs["mask"]), p(outs["loc"]), p(outs["center"]), p(outs["size"]), p(outs["scale"]), st))
        _lib.check(L.h3d_detect_keypoints(ctx.h, p(sm), 2, 32, 32, 21, p(outs["uv"]), st))
        _lib.check(L.h3d_fully_connected_f32(ctx.h, p(fx), p(fw), p(fb), p(outs["fy"]), 4, 300, 40, 1, st))
        _lib.check(L.h3d_conv2d_tc_packed(ctx.h, p(cx), pk.h, p(outs["cy"]), 2, 16, 16, 1, 1, st))
        _lib.check(L.h3d_conv2d_f32(ctx.h, p(cx), p(_w), p(_b), p(outs["ty"]), 2, 16, 16, 64, 64, 3, 2, 1, st))

    _w = _dev(wk); _b = _dev(np.zeros(64, f32))
    cur = torch.cuda.current_stream().cuda_stream
    round_(cur); torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(3):
        round_(cur)
    free1 = torch.cuda.mem_get_info()[0]
    torch.cuda.synchronize()
    assert free0 == free1, "an operator entry allocated device memory per call (%d -> %d bytes free)" % (free0, free1)
    eager = {k: v.clone() for k, v in outs.items()}
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for v in outs.values():
            v.zero_()
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(g):
        round_(torch.cuda.current_stream().cuda_stream)
    g.replay(); torch.cuda.synchronize()
    for k in outs:
        assert torch.equal(outs[k], eager[k]), k
    ref = T.leaky_relu(T.conv2d_same(cx.cpu().numpy(), wk, np.zeros(64, f32), 1, np.float64))
    assert np.abs(outs["cy"].cpu().numpy() - ref).max() < 5e-5
    ctx.check_errors()
