"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the TensorFlow-1.3 op semantics the
reference (lmb-freiburg/hand3d) relies on.

PARITY UNPINNED: the reference has no tests / golden tensors and TensorFlow 1.3 cannot be
installed in this image, so this file restates the *published* TF 1.3
kernel behaviour and is pinned by the hand-derivable known-answer
vectors in tests/test_oracle_kat.py, by the vectors of TensorFlow's own op tests and by independent
implementations (scipy, scalar transcriptions of the published kernels) in tests/test_tf_published_vectors.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (hand3d_b200/) never does.

All functions take / return numpy float32 NHWC arrays (or torch CPU tensors where noted) and
use separate multiply / add (numpy never contracts to FMA), matching the SSE4-only TF 1.3 wheels.
Set ``dtype=np.float64`` where offered to obtain the high-precision twin used to bound fp32 error.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

NEG_SLOPE = 0.01  # utils/general.py:28


def _t(x, dtype):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dtype)


def leaky_relu(x):
    """utils/general.py:31-33  tf.maximum(x, 0.01*x)."""
    return np.maximum(x, x.dtype.type(NEG_SLOPE) * x)


def same_pad(in_size: int, k: int, s: int):
    """TF 'SAME' padding: out=ceil(in/s); total=max((out-1)*s+k-in,0); before=total//2."""
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    return total // 2, total - total // 2


def conv2d_same(x, w, b, stride=1, dtype=np.float32):
    """utils/general.py:36-53  tf.nn.conv2d(x, W[kh,kw,Cin,Cout], [1,s,s,1], 'SAME') + bias.

    x: [B,H,W,Cin] NHWC, w: HWIO, b: [Cout].  Cross-correlation (no kernel flip)."""
    td = torch.float64 if dtype == np.float64 else torch.float32
    B, H, W, Cin = x.shape
    kh, kw, ci, co = w.shape
    assert ci == Cin
    pt, pb = same_pad(H, kh, stride)
    pl, pr = same_pad(W, kw, stride)
    xt = _t(x, td).permute(0, 3, 1, 2)
    xt = F.pad(xt, (pl, pr, pt, pb))
    wt = _t(w, td).permute(3, 2, 0, 1).contiguous()
    y = F.conv2d(xt, wt, _t(b, td), stride=stride)
    return y.permute(0, 2, 3, 1).contiguous().numpy()


def conv_relu(x, w, b, stride=1, dtype=np.float32):
    """utils/general.py:56-59."""
    return leaky_relu(conv2d_same(x, w, b, stride, dtype))


def max_pool_2x2(x):
    """utils/general.py:62-65  tf.nn.max_pool 2x2 / stride 2 / VALID."""
    B, H, W, C = x.shape
    Ho, Wo = (H - 2) // 2 + 1, (W - 2) // 2 + 1
    v = x[:, : Ho * 2, : Wo * 2, :].reshape(B, Ho, 2, Wo, 2, C)
    return v.max(axis=(2, 4))


def avg_pool_8x8(x):
    """nets/PosePriorNetwork.py:61  tf.nn.avg_pool 8x8 / 8 / SAME on sizes divisible by 8."""
    B, H, W, C = x.shape
    assert H % 8 == 0 and W % 8 == 0
    v = x.reshape(B, H // 8, 8, W // 8, 8, C)
    # Eigen sums the window then divides by the (full) window count.
    return (v.sum(axis=(2, 4), dtype=x.dtype) / x.dtype.type(64.0)).astype(x.dtype)


def fully_connected(x, w, b, dtype=np.float32):
    """utils/general.py:113-130  x @ W[in,out] + b (check_numerics on W and b)."""
    if not (np.all(np.isfinite(w)) and np.all(np.isfinite(b))):
        raise FloatingPointError("check_numerics: FC weights/biases contain NaN/Inf")
    td = torch.float64 if dtype == np.float64 else torch.float32
    return (_t(x, td) @ _t(w, td) + _t(b, td)).numpy()


def resize_bilinear_tf1(x, out_h: int, out_w: int):
    """tf.image.resize_images(x,(h,w)): bilinear, align_corners=False, legacy (no half-pixel
    centres) -- SURVEY 9.3; call sites nets/ColorHandPose3DNetwork.py:97,128,166."""
    B, H, W, C = x.shape
    if (H, W) == (out_h, out_w):
        return x
    ft = x.dtype.type
    hs = ft(H) / ft(out_h)
    ws = ft(W) / ft(out_w)
    in_y = np.arange(out_h, dtype=x.dtype) * hs
    in_x = np.arange(out_w, dtype=x.dtype) * ws
    y0 = np.floor(in_y).astype(np.int64)
    x0 = np.floor(in_x).astype(np.int64)
    y1 = np.minimum(y0 + 1, H - 1)
    x1 = np.minimum(x0 + 1, W - 1)
    ly = (in_y - y0.astype(x.dtype)).reshape(1, out_h, 1, 1)
    lx = (in_x - x0.astype(x.dtype)).reshape(1, 1, out_w, 1)
    rows0 = x[:, y0]
    rows1 = x[:, y1]
    tl, tr = rows0[:, :, x0], rows0[:, :, x1]
    bl, br = rows1[:, :, x0], rows1[:, :, x1]
    top = tl + (tr - tl) * lx
    bot = bl + (br - bl) * lx
    return (top + (bot - top) * ly).astype(x.dtype)


def softmax_last(x):
    """tf.nn.softmax (Eigen, TF 1.3 softmax_op_functor.h): e=exp(x-max); p=e*(1/sum e)."""
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m)
    inv = x.dtype.type(1.0) / e.sum(axis=-1, keepdims=True, dtype=x.dtype)
    return e * inv


def round_half_even(x):
    """tf.round = round half to even (np.round/np.rint do the same)."""
    return np.rint(x)


def dilation2d_21(obj):
    """tf.nn.dilation2d(obj, ones(21,21,1)/441, strides 1, rates 1, 'SAME')  (utils/general.py:249,259).

    out[y,x] = max over in-bounds taps (dy,dx in [-10,10]) of in + 1/441.  obj: [H,W] float32."""
    t = torch.from_numpy(np.ascontiguousarray(obj))[None, None]
    # max_pool2d pads with -inf == "out-of-bounds taps are skipped"
    d = F.max_pool2d(t, kernel_size=21, stride=1, padding=10)[0, 0].numpy()
    return d + obj.dtype.type(1.0 / 441.0)


def dilation2d(x, filt, strides=(1, 1), rates=(1, 1), padding="SAME"):
    """tf.nn.dilation2d, general form (TF 1.3 core/kernels/dilation_ops.cc, DilationOp / ParseSizes): grey-scale max-sum
    correlation  out[b,y,x,c] = max_{dy,dx in bounds} in[b, y*sy - pad_top + ry*dy, x*sx - pad_left + rx*dx, c] + filter[dy,dx,c],
    out-of-bounds taps skipped, SAME padding computed on the dilated ("effective") filter size, before = total // 2.
    Brute-force loops: used to pin dilation2d_21 (the 21x21 / 441 case of utils/general.py:249,259) and for TF's op-test vectors."""
    x = np.asarray(x); filt = np.asarray(filt)
    B, H, W, C = x.shape
    fh, fw, fc = filt.shape
    assert fc == C
    sy, sx = strides; ry, rx = rates
    eh, ew = (fh - 1) * ry + 1, (fw - 1) * rx + 1
    if padding == "SAME":
        Ho, Wo = -(-H // sy), -(-W // sx)
        pt = max((Ho - 1) * sy + eh - H, 0) // 2
        pl = max((Wo - 1) * sx + ew - W, 0) // 2
    else:
        Ho, Wo = (H - eh) // sy + 1, (W - ew) // sx + 1
        pt = pl = 0
    out = np.full((B, Ho, Wo, C), -np.inf, dtype=x.dtype)
    for dy in range(fh):
        for dx in range(fw):
            for y in range(Ho):
                iy = y * sy - pt + ry * dy
                if iy < 0 or iy >= H:
                    continue
                for xo in range(Wo):
                    ix = xo * sx - pl + rx * dx
                    if ix < 0 or ix >= W:
                        continue
                    out[:, y, xo, :] = np.maximum(out[:, y, xo, :], x[:, iy, ix, :] + filt[dy, dx, :])
    return out


def crop_and_resize(image, boxes, crop_h: int, crop_w: int, extrapolation_value=0.0):
    """tf.image.crop_and_resize(image, boxes, box_ind=range(B), [crop_h,crop_w]) -- bilinear,
    extrapolation_value 0 at the reference's call site (TF 1.3 crop_and_resize_op.cc).
    boxes [B,4]=(y1,x1,y2,x2) normalised."""
    B, H, W, C = image.shape
    f = np.float32
    out = np.zeros((B, crop_h, crop_w, C), dtype=np.float32)
    for b in range(B):
        y1, x1, y2, x2 = [f(v) for v in boxes[b]]
        hs = (y2 - y1) * f(H - 1) / f(crop_h - 1) if crop_h > 1 else f(0)
        ws = (x2 - x1) * f(W - 1) / f(crop_w - 1) if crop_w > 1 else f(0)
        ys = np.arange(crop_h, dtype=np.float32)
        xs = np.arange(crop_w, dtype=np.float32)
        in_y = y1 * f(H - 1) + ys * hs if crop_h > 1 else np.full(1, f(0.5) * (y1 + y2) * f(H - 1), f)
        in_x = x1 * f(W - 1) + xs * ws if crop_w > 1 else np.full(1, f(0.5) * (x1 + x2) * f(W - 1), f)
        in_y = in_y.astype(f)
        in_x = in_x.astype(f)
        vy = ~((in_y < 0) | (in_y > f(H - 1)))
        vx = ~((in_x < 0) | (in_x > f(W - 1)))
        cy = np.where(vy, in_y, f(0))
        cx = np.where(vx, in_x, f(0))
        top = np.floor(cy).astype(np.int64)
        bot = np.ceil(cy).astype(np.int64)
        lef = np.floor(cx).astype(np.int64)
        rig = np.ceil(cx).astype(np.int64)
        ly = (cy - top.astype(f)).reshape(crop_h, 1, 1)
        lx = (cx - lef.astype(f)).reshape(1, crop_w, 1)
        img = image[b]
        tl, tr = img[top][:, lef], img[top][:, rig]
        bl, br = img[bot][:, lef], img[bot][:, rig]
        t = tl + (tr - tl) * lx
        bo = bl + (br - bl) * lx
        val = (t + (bo - t) * ly).astype(f)
        mask = (vy.reshape(crop_h, 1, 1) & vx.reshape(1, crop_w, 1))
        out[b] = np.where(mask, val, f(extrapolation_value))
    return out
