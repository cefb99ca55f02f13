"""GPU parity of the tcgen05 implicit-GEMM convolution (operator entry h3d_conv2d_tc) against the fp64 oracle."""
import numpy as np
import pytest
import torch

from oracle import tf1_ops as T

pytestmark = pytest.mark.gpu
f32 = np.float32

# tolerance on outputs of unit scale: 3-pass split modes are fp32-grade, single-pass modes are 16-bit grade
TOL = {"bf16x3": 5e-5, "fp16x3": 2e-5, "fp16": 6e-3, "bf16": 5e-2, "fp16_f8c": 2e-4}

CASES = [  # B,H,W,Cin,Cout,k
    (1, 16, 8, 64, 64, 1),      # one exact tile, one K block: the smallest possible case
    (1, 16, 8, 64, 64, 3),      # halo / zero padding
    (2, 32, 32, 128, 128, 3),   # several tiles, BN = 128
    (3, 40, 40, 64, 256, 3),    # (8,8,2) tiles with a ragged batch, 2 N tiles
    (1, 24, 40, 192, 128, 7),   # 7x7, 3 channel chunks, partial tiles
    (2, 20, 12, 100, 72, 3),    # channel padding on both sides
    (1, 64, 64, 256, 512, 3),   # long K loop, pipeline wrap-around, many tiles per CTA
    (2, 20, 40, 64, 64, 3),     # 64 -> 64 specialisation (weights-resident, patch re-use): ragged rows and columns, several tiles
    (3, 64, 64, 64, 64, 3),     # 64 -> 64: more tiles than fit one wave of the A ring
    (2, 24, 48, 64, 128, 3),    # 64 -> 128 (conv2_1): two 64-channel groups, CTAs split between them
    (4, 80, 80, 128, 256, 3),   # enough tiles (200) for the CTA-pair kernel to be chosen by the policy itself
    (1, 24, 16, 64, 64, 3),     # 64 -> 64 on a CTA pair: odd number of pixel tiles (the last pair's second tile is out of range)
    (2, 48, 32, 128, 128, 3),   # Cout = 128: N-stacked CTA-pair kernel chosen by the policy, 4 K blocks per tap
]


@pytest.fixture
def tuning(ctx):
    """Sets kernel-selection switches for one test and restores the policy defaults afterwards."""
    defaults = {"tc_2cta": -1, "tc_bn": 0, "tc_c64": 1, "tc_c64x2": 1, "tc_pair128": 1, "tc_stack": 1}
    yield ctx.set_tuning
    for k, v in defaults.items():
        ctx.set_tuning(k, v)


@pytest.fixture(scope="module")
def ctx():
    from hand3d_b200 import runtime
    return runtime.default_context()


@pytest.mark.parametrize("prec", ["bf16x3", "fp16x3", "fp16", "bf16", "fp16_f8c"])
@pytest.mark.parametrize("case", CASES)
def test_conv2d_tc_vs_oracle(ctx, case, prec):
    B, H, W, Cin, Cout, k = case
    rng = np.random.default_rng(11)
    x = rng.normal(size=(B, H, W, Cin)).astype(f32)
    w = (rng.normal(size=(k, k, Cin, Cout)) / np.sqrt(k * k * Cin)).astype(f32)
    b = rng.normal(size=Cout).astype(f32)
    y = ctx.conv2d_tc(torch.from_numpy(x).cuda(), w, b, leaky=True, precision=prec).cpu().numpy()
    ref = T.leaky_relu(T.conv2d_same(x, w, b, 1, np.float64))
    err = np.abs(y - ref).max()
    assert err < TOL[prec], "max abs err %.3e (tolerance %.1e)" % (err, TOL[prec])


@pytest.mark.parametrize("prec", ["bf16x3", "fp16", "fp16_f8c"])
@pytest.mark.parametrize("case", [CASES[3], CASES[4], CASES[6]])
def test_conv2d_tc_forced_cta_pair(ctx, case, prec, tuning):
    """Small problems normally fall back to single-CTA tiles; tc_2cta = 1 forces the cta_group::2 kernel onto them
    (ragged pairs, odd tile counts, N = 128 (N-stacked in the 3-pass mode) / 256 pair tiles)."""
    tuning("tc_2cta", 1)
    B, H, W, Cin, Cout, k = case
    rng = np.random.default_rng(15)
    x = rng.normal(size=(B, H, W, Cin)).astype(f32)
    w = (rng.normal(size=(k, k, Cin, Cout)) / np.sqrt(k * k * Cin)).astype(f32)
    b = rng.normal(size=Cout).astype(f32)
    y = ctx.conv2d_tc(torch.from_numpy(x).cuda(), w, b, leaky=True, precision=prec).cpu().numpy()
    ref = T.leaky_relu(T.conv2d_same(x, w, b, 1, np.float64))
    err = np.abs(y - ref).max()
    assert err < TOL[prec], "max abs err %.3e (tolerance %.1e)" % (err, TOL[prec])


@pytest.mark.parametrize("switch", ["tc_pair128", "tc_c64x2", "tc_c64", "tc_stack"])
@pytest.mark.parametrize("case", [CASES[2], CASES[4], CASES[7], CASES[9], CASES[11]])
def test_conv2d_tc_single_cta_variants(ctx, case, switch, tuning):
    """The kernels the policy no longer picks for these shapes (single-CTA stacked N = 128, single-CTA 64-channel kernel, the
    generic kernel for 64 -> 64, un-stacked passes) stay correct: they serve the small maps and the single-pass modes."""
    tuning(switch, 0)
    B, H, W, Cin, Cout, k = case
    rng = np.random.default_rng(16)
    x = rng.normal(size=(B, H, W, Cin)).astype(f32)
    w = (rng.normal(size=(k, k, Cin, Cout)) / np.sqrt(k * k * Cin)).astype(f32)
    b = rng.normal(size=Cout).astype(f32)
    y = ctx.conv2d_tc(torch.from_numpy(x).cuda(), w, b, leaky=True, precision="bf16x3").cpu().numpy()
    ref = T.leaky_relu(T.conv2d_same(x, w, b, 1, np.float64))
    err = np.abs(y - ref).max()
    assert err < TOL["bf16x3"], "max abs err %.3e (tolerance %.1e)" % (err, TOL["bf16x3"])


@pytest.mark.parametrize("pair", [1, 0])
def test_conv2d_tc_c64_two_channel_groups(ctx, pair, tuning):
    """64 -> 128 channels on the 64-channel kernels (two resident 64-channel weight groups, CTAs / clusters split between them):
    no longer the policy's choice for conv2_1 in the 3-pass modes (tc_c64 = 2 forces it), still the fp16 single-pass path."""
    tuning("tc_c64", 2)
    tuning("tc_c64x2", pair)
    B, H, W, Cin, Cout, k = CASES[9]
    rng = np.random.default_rng(17)
    x = rng.normal(size=(B, H, W, Cin)).astype(f32)
    w = (rng.normal(size=(k, k, Cin, Cout)) / np.sqrt(k * k * Cin)).astype(f32)
    b = rng.normal(size=Cout).astype(f32)
    for prec in ("bf16x3", "fp16"):
        y = ctx.conv2d_tc(torch.from_numpy(x).cuda(), w, b, leaky=True, precision=prec).cpu().numpy()
        ref = T.leaky_relu(T.conv2d_same(x, w, b, 1, np.float64))
        assert np.abs(y - ref).max() < TOL[prec]


STRIDED = [  # B,H,W,Cin,Cout: the stride-2 layers of the lifting pyramids (nets/ColorHandPose3DNetwork.py:255-258,291-294)
    (2, 32, 32, 32, 32),      # conv_pose_0_2: Cin / Cout padded 32 -> 64
    (3, 16, 16, 64, 64),      # conv_pose_1_2 / conv_vp_0_2 geometry
    (5, 8, 8, 128, 128),      # conv_pose_2_2: (8,8,2) tiles, ragged batch
    (2, 8, 8, 256, 256),      # conv_vp_2_2: CTA-pair kernel
    (1, 12, 20, 21, 40),      # odd channel counts, partial tiles
    (2, 32, 32, 64, 64),      # 64 -> 64 on a CTA pair with the stride-2 epilogue
    (2, 32, 32, 128, 128),    # N-stacked CTA-pair kernel with the stride-2 epilogue
]


@pytest.mark.parametrize("prec", ["bf16x3", "fp16x3", "fp16"])
@pytest.mark.parametrize("case", STRIDED)
def test_conv2d_tc_stride2_vs_oracle(ctx, case, prec):
    """stride 2 'SAME' on an even-sized map (pads 0 before / 1 after) = the odd pixels of the stride-1 result."""
    B, H, W, Cin, Cout = case
    rng = np.random.default_rng(13)
    x = rng.normal(size=(B, H, W, Cin)).astype(f32)
    w = (rng.normal(size=(3, 3, Cin, Cout)) / np.sqrt(9 * Cin)).astype(f32)
    b = rng.normal(size=Cout).astype(f32)
    y = ctx.conv2d_tc(torch.from_numpy(x).cuda(), w, b, leaky=True, precision=prec, stride=2).cpu().numpy()
    ref = T.leaky_relu(T.conv2d_same(x, w, b, 2, np.float64))
    assert y.shape == ref.shape
    err = np.abs(y - ref).max()
    assert err < TOL[prec], "max abs err %.3e (tolerance %.1e)" % (err, TOL[prec])


def test_conv2d_tc_padded_planes_chain(ctx):
    """Cout = 32 feeding a second layer: the padding channels of the split planes must be exact zeros (they are the next K)."""
    rng = np.random.default_rng(14)
    x = rng.normal(size=(2, 16, 16, 21)).astype(f32)
    w1 = (rng.normal(size=(3, 3, 21, 32)) / np.sqrt(9 * 21)).astype(f32); b1 = rng.normal(size=32).astype(f32)
    y1 = ctx.conv2d_tc(torch.from_numpy(x).cuda(), w1, b1, leaky=True, precision="fp16x3")
    ref = T.leaky_relu(T.conv2d_same(x, w1, b1, 1, np.float64))
    assert np.abs(y1.cpu().numpy() - ref).max() < TOL["fp16x3"]


def test_conv2d_tc_identity_weights(ctx):
    """Delta kernel = identity: any layout / swizzle / descriptor mistake shows up as permuted channels or pixels."""
    B, H, W, C = 1, 16, 16, 64
    x = np.random.default_rng(12).normal(size=(B, H, W, C)).astype(f32)
    w = np.zeros((3, 3, C, C), f32)
    w[1, 1] = np.eye(C, dtype=f32)
    y = ctx.conv2d_tc(torch.from_numpy(x).cuda(), w, np.zeros(C, f32), leaky=False, precision="fp16x3").cpu().numpy()
    np.testing.assert_allclose(y, x, atol=1e-6)
    w2 = np.zeros((3, 3, C, C), f32)
    w2[0, 2] = np.eye(C, dtype=f32)         # tap (kh=0, kw=2): y[h,w] = x[h-1, w+1]
    y2 = ctx.conv2d_tc(torch.from_numpy(x).cuda(), w2, np.zeros(C, f32), leaky=False, precision="fp16x3").cpu().numpy()
    ref = np.zeros_like(x); ref[:, 1:, :-1] = x[:, :-1, 1:]
    np.testing.assert_allclose(y2, ref, atol=1e-6)
