import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np
from fyusenet_b200 import capi
from gpu_util import conv_gpu, half, random_wb, rel_l2
rng = np.random.default_rng(0)
os.environ["FYN_DEEP_PERSIST"] = "2"
for (ci, co, size, batch) in [(64, 64, 28, 2), (64, 64, 7, 1), (128, 64, 7, 1), (64, 128, 7, 1), (64, 256, 7, 1), (128, 128, 7, 3), (512, 512, 7, 1), (512, 128, 7, 1), (256, 256, 14, 2), (128, 128, 28, 1), (64, 64, 8, 1), (64, 64, 6, 1)]:
    x = half(rng.normal(size=(batch, ci, size, size)).astype(np.float32))
    wb = random_wb(rng, ci, co, 3, post_bn=True)
    kw = dict(out_channels=co, kernel=3, in_pad=1, flags=capi.FLAG_PRE_RELU | capi.FLAG_POST_BATCHNORM, deep=True, backend=capi.BACKEND_TC)
    os.environ["FYN_DEEP_HALO"] = "0"
    ref = conv_gpu(x, wb, **kw)
    os.environ["FYN_DEEP_HALO"] = "1"
    for sets in ("1", "3"):
        os.environ["FYN_DEEP_SETS"] = sets
        y = conv_gpu(x, wb, **kw)
        err = np.abs(y - ref)
        bad = np.argwhere(err > 0.05)
        print(ci, co, size, batch, "sets", sets, "rel_l2 %.2e" % rel_l2(y, ref), "nbad", len(bad), "first bad", bad[:3].tolist(), "chan set", sorted(set(bad[:, 1].tolist()))[:12] if len(bad) else "")
