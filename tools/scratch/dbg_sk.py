import os, sys
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/oracle")
import gpu_util
from gpu_util import conv_gpu, half, random_wb, rel_l2
from fyusenet_b200 import capi
k, ds, ci, co, inp, outp, postbn, size, batch = 1, 1, 2048, 512, 0, 1, True, 7, 1
rng = np.random.default_rng(1)
x = half(rng.normal(size=(batch, ci, size, size)).astype(np.float32))
wb = random_wb(rng, ci, co, k, post_bn=postbn)
fl = capi.FLAG_PRE_RELU | capi.FLAG_POST_BATCHNORM
kw = dict(out_channels=co, kernel=k, downsample=ds, in_pad=inp, out_pad=outp, flags=fl, deep=True, backend=capi.BACKEND_TC)
os.environ["FYN_DEEP_SPLITK"] = "0"
base = conv_gpu(x, wb, **kw)
os.environ["FYN_DEEP_SPLITK"] = "2"
for nt, split in (("64", "2"),):
    for name, v in (("FYN_DEEP_SK_NT", nt), ("FYN_DEEP_SK_SPLIT", split)):
        if v is None: os.environ.pop(name, None)
        else: os.environ[name] = v
    for rep in range(1):
        y = conv_gpu(x, wb, **kw).reshape(base.shape)
        bad = ~np.isfinite(y)
        d = np.abs(np.nan_to_num(y) - base)
        chans = np.unique(np.argwhere(d > 1e-2)[:, -3]) if d.ndim >= 3 else []
        print(nt, split, hex(gpu_util.LAST_KERNEL), "nan", int(bad.sum()), "maxdiff", float(d.max()), "bad chans", list(chans)[:40], flush=True)
