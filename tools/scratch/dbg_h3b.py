import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np
from fyusenet_b200 import capi
from gpu_util import conv_gpu, half, random_wb, rel_l2
k, ds, ci, co, inp, outp, size, batch = 3, 1, 512, 512, 1, 0, 7, 1
rng = np.random.default_rng(k * 100 + ci + co + size)
x = half(rng.normal(size=(batch, ci, size, size)).astype(np.float32))
wb = random_wb(rng, ci, co, k, post_bn=True)
kw = dict(out_channels=co, kernel=k, downsample=ds, in_pad=inp, out_pad=outp, flags=capi.FLAG_PRE_RELU | capi.FLAG_POST_BATCHNORM, residual=None, deep=True, backend=capi.BACKEND_TC)
os.environ["FYN_DEEP_HALO"] = "0"
os.environ["FYN_DEEP_PERSIST"] = "0"
ref = conv_gpu(x, wb, **kw)
os.environ["FYN_DEEP_PERSIST"] = "2"
print("persist2 halo0", rel_l2(conv_gpu(x, wb, **kw), ref))
for ring, sets in (("2", "2"), ("3", "1"), ("5", "4")):
    os.environ["FYN_DEEP_PRING"] = ring
    os.environ["FYN_DEEP_SETS"] = sets
    print("ring", ring, "sets", sets, rel_l2(conv_gpu(x, wb, **kw), ref))
del os.environ["FYN_DEEP_HALO"]
del os.environ["FYN_DEEP_PRING"]
for rep in range(3):
    for sets in ("1", "2", "3"):
        os.environ["FYN_DEEP_SETS"] = sets
        y = conv_gpu(x, wb, **kw)
        err = np.abs(y - ref)
        bad = np.argwhere(err > 0.05)
        print("H3 sets", sets, "rel_l2 %.2e" % rel_l2(y, ref), "nbad", len(bad), "chans", sorted(set(bad[:, 1].tolist()))[:20] if len(bad) else "", "pos", sorted(set(map(tuple, bad[:, 2:].tolist())))[:10] if len(bad) else "")
