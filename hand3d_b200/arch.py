"""Layer tables of the four sub-networks (variable names / shapes the reference's weight pickles use).

Follows nets/ColorHandPose3DNetwork.py:142-161 (HandSegNet), :181-215 (PoseNet2D),
:251-267 (PosePrior), :276-308 (ViewpointNet) and SURVEY.md section 8a.  Each entry is
(layer_name, kernel, stride, cin, cout, leaky_relu).  FC entries have kernel == 0 and cin == fan-in.
"""
from __future__ import annotations

HANDSEGNET = [
    ("conv1_1", 3, 1, 3, 64, True), ("conv1_2", 3, 1, 64, 64, True),
    ("conv2_1", 3, 1, 64, 128, True), ("conv2_2", 3, 1, 128, 128, True),
    ("conv3_1", 3, 1, 128, 256, True), ("conv3_2", 3, 1, 256, 256, True),
    ("conv3_3", 3, 1, 256, 256, True), ("conv3_4", 3, 1, 256, 256, True),
    ("conv4_1", 3, 1, 256, 512, True), ("conv4_2", 3, 1, 512, 512, True),
    ("conv4_3", 3, 1, 512, 512, True), ("conv4_4", 3, 1, 512, 512, True),
    ("conv5_1", 3, 1, 512, 512, True), ("conv5_2", 3, 1, 512, 128, True),
    ("conv6_1", 1, 1, 128, 512, True), ("conv6_2", 1, 1, 512, 2, False),
]
HANDSEGNET_POOL_AFTER = ("conv1_2", "conv2_2", "conv3_4")

POSENET2D = [
    ("conv1_1", 3, 1, 3, 64, True), ("conv1_2", 3, 1, 64, 64, True),
    ("conv2_1", 3, 1, 64, 128, True), ("conv2_2", 3, 1, 128, 128, True),
    ("conv3_1", 3, 1, 128, 256, True), ("conv3_2", 3, 1, 256, 256, True),
    ("conv3_3", 3, 1, 256, 256, True), ("conv3_4", 3, 1, 256, 256, True),
    ("conv4_1", 3, 1, 256, 512, True), ("conv4_2", 3, 1, 512, 512, True),
    ("conv4_3", 3, 1, 512, 256, True), ("conv4_4", 3, 1, 256, 256, True),
    ("conv4_5", 3, 1, 256, 256, True), ("conv4_6", 3, 1, 256, 256, True),
    ("conv4_7", 3, 1, 256, 128, True),
    ("conv5_1", 1, 1, 128, 512, True), ("conv5_2", 1, 1, 512, 21, False),
]
for _u in (6, 7):
    POSENET2D += [("conv%d_1" % _u, 7, 1, 149, 128, True)]
    POSENET2D += [("conv%d_%d" % (_u, i), 7, 1, 128, 128, True) for i in range(2, 6)]
    POSENET2D += [("conv%d_6" % _u, 1, 1, 128, 128, True), ("conv%d_7" % _u, 1, 1, 128, 21, False)]
POSENET2D_POOL_AFTER = ("conv1_2", "conv2_2", "conv3_4")

POSEPRIOR = [
    ("conv_pose_0_1", 3, 1, 21, 32, True), ("conv_pose_0_2", 3, 2, 32, 32, True),
    ("conv_pose_1_1", 3, 1, 32, 64, True), ("conv_pose_1_2", 3, 2, 64, 64, True),
    ("conv_pose_2_1", 3, 1, 64, 128, True), ("conv_pose_2_2", 3, 2, 128, 128, True),
    ("fc_rel0", 0, 0, 2050, 512, True), ("fc_rel1", 0, 0, 512, 512, True),
    ("fc_xyz", 0, 0, 512, 63, False),
]
POSEPRIOR_BOTTLENECK = ("fc_bottleneck", 0, 0, 512, 30, False)   # nets/PosePriorNetwork.py:115-116

VIEWPOINT = [
    ("conv_vp_0_1", 3, 1, 21, 64, True), ("conv_vp_0_2", 3, 2, 64, 64, True),
    ("conv_vp_1_1", 3, 1, 64, 128, True), ("conv_vp_1_2", 3, 2, 128, 128, True),
    ("conv_vp_2_1", 3, 1, 128, 256, True), ("conv_vp_2_2", 3, 2, 256, 256, True),
    ("fc_vp0", 0, 0, 4098, 256, True), ("fc_vp1", 0, 0, 256, 128, True),
    ("fc_vp_ux", 0, 0, 128, 1, False), ("fc_vp_uy", 0, 0, 128, 1, False), ("fc_vp_uz", 0, 0, 128, 1, False),
]

NETS = {"HandSegNet": HANDSEGNET, "PoseNet2D": POSENET2D, "PosePrior": POSEPRIOR, "ViewpointNet": VIEWPOINT}


def variable_shapes(bottleneck: bool = False):
    """Ordered {variable_name: shape} for all four scopes."""
    out = {}
    for scope, layers in NETS.items():
        layers = list(layers)
        if scope == "PosePrior" and bottleneck:
            layers.insert(8, POSEPRIOR_BOTTLENECK)
            layers[9] = ("fc_xyz", 0, 0, 30, 63, False)
        for name, k, s, cin, cout, _ in layers:
            out["%s/%s/weights" % (scope, name)] = (k, k, cin, cout) if k else (cin, cout)
            out["%s/%s/biases" % (scope, name)] = (cout,)
    return out


def conv_flops_per_image(H=320, W=320, crop=256):
    """2*MAC of all conv+FC layers for one image (142.258 GFLOP at 320x320)."""
    total = 0
    for scope, layers, (h, w) in (("HandSegNet", HANDSEGNET, (H, W)), ("PoseNet2D", POSENET2D, (crop, crop)),
                                  ("PosePrior", POSEPRIOR, (32, 32)), ("ViewpointNet", VIEWPOINT, (32, 32))):
        pools = HANDSEGNET_POOL_AFTER if scope in ("HandSegNet", "PoseNet2D") else ()
        for name, k, s, cin, cout, _ in layers:
            if k == 0:
                total += 2 * cin * cout
                continue
            h, w = -(-h // s), -(-w // s)
            total += 2 * h * w * k * k * cin * cout
            if name in pools:
                h, w = h // 2, w // 2
    return total
