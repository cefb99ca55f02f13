import numpy as np, torch, sys
from hand3d_b200 import runtime
from oracle import tf1_ops as T
ctx = runtime.default_context()
rng = np.random.default_rng(11)
B,H,W,Cin,Cout,k = 1,16,8,64,64,1
x = rng.normal(size=(B,H,W,Cin)).astype(np.float32)
w = (rng.normal(size=(k,k,Cin,Cout))/np.sqrt(k*k*Cin)).astype(np.float32)
b = np.zeros(Cout, np.float32)
y = ctx.conv2d_tc(torch.from_numpy(x).cuda(), w, b, leaky=False, precision="fp16_f8c").cpu().numpy()
ref = T.conv2d_same(x, w, b, 1, np.float64)
err = np.abs(y-ref)
print("max err", err.max(), "ratio y/ref median", np.median(y/ref), "nan", np.isnan(y).sum(), "y[0,0,0,:4]", y[0,0,0,:4], "ref", ref[0,0,0,:4])
