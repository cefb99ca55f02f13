#!/usr/bin/env python
"""Single-layer micro-benchmark of the tensor-core convolution through the enqueue-only operator entry (h3d_conv2d_tc_packed):
    python scripts/experiments/mb_conv.py B H W Cin Cout k [key=value ...]     (tuning keys of h3d_set_tuning)
Prints the median CUDA-event time of the conv kernel alone (the plane conversion kernels are timed separately and subtracted)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hand3d_b200 import runtime  # noqa: E402


def main():
    B, H, W, Cin, Cout, k = [int(v) for v in sys.argv[1:7]]
    ctx = runtime.default_context()
    for kv in sys.argv[7:]:
        key, val = kv.split("=")
        ctx.set_tuning(key, int(val))
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.normal(size=(B, H, W, Cin)).astype(np.float32)).cuda()
    w = (rng.normal(size=(k, k, Cin, Cout)) / np.sqrt(k * k * Cin)).astype(np.float32)
    pk = ctx.pack_conv(w, np.zeros(Cout, np.float32), "bf16x3")
    for _ in range(3):
        ctx.conv2d_tc_packed(x, pk, leaky=True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.conv2d_tc_packed(x, pk, leaky=True); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    flops = 2.0 * B * H * W * k * k * Cin * Cout
    t = float(np.median(ts))
    print("B=%d %dx%d %d->%d k=%d %s: %.1f us total (conv + plane conversions), %.0f TFLOP/s algorithmic over the total" % (
        B, H, W, Cin, Cout, k, " ".join(sys.argv[7:]), 1000 * t, flops / t / 1e9))


if __name__ == "__main__":
    main()
