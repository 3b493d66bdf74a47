import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np
import fyn_oracle as fo
from fyusenet_b200 import capi
from gpu_util import conv_gpu, half, random_wb, rel_l2
k, ds, ci, co, inp, outp, size, batch = 3, 1, 512, 512, 1, 0, 7, 1
rng = np.random.default_rng(k * 100 + ci + co + size)
x = half(rng.normal(size=(batch, ci, size, size)).astype(np.float32))
wb = random_wb(rng, ci, co, k, post_bn=True)
kw = dict(out_channels=co, kernel=k, downsample=ds, in_pad=inp, out_pad=outp, flags=capi.FLAG_PRE_RELU | capi.FLAG_POST_BATCHNORM, residual=None, deep=True, backend=capi.BACKEND_TC)
os.environ["FYN_DEEP_HALO"] = "0"
os.environ["FYN_DEEP_PERSIST"] = "0"
old = conv_gpu(x, wb, **kw)
ref = np.stack([fo.conv2d(x[i], wb, co, k, downsample=ds, in_pad=inp, out_pad=outp, act=fo.ACT_RELU, flags=fo.POST_BATCHNORM, deep=True, residual=None, prec=fo.FP16_STORE) for i in range(1)])
print("old vs oracle: max", np.abs(old - ref).max(), "at", np.unravel_index(np.abs(old - ref).argmax(), ref.shape))
os.environ["FYN_DEEP_PERSIST"] = "2"
del os.environ["FYN_DEEP_HALO"]
y = conv_gpu(x, wb, **kw)
print("h3 vs oracle: max", np.abs(y - ref).max(), "h3 vs old", np.abs(y - old).max())
i = (0, 480, 6, 5)
print("values at", i, "old", old[i], "h3", y[i], "oracle", ref[i])
