import os, sys
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/oracle")
import gpu_util
from gpu_util import conv_gpu
from fyusenet_b200 import capi
ci, co, size = 256, 64, 7
x = np.ones((1, ci, size, size), np.float32)
W = np.zeros((co, ci), np.float32)
for st in range(4):
    W[:, st * 64:(st + 1) * 64] = (10.0 ** st) / 64
for o in range(co):
    W[o] *= 1 + o / 64.0
wb = np.concatenate([np.zeros(co, np.float32), W.reshape(-1)])
kw = dict(out_channels=co, kernel=1, flags=0, deep=True, backend=capi.BACKEND_TC)
os.environ["FYN_DEEP_SPLITK"] = "2"
np.set_printoptions(linewidth=250, precision=1, suppress=True)
for nt, split in (("64", "1"), ("64", "2"), ("64", "4"), ("32", "2")):
    os.environ["FYN_DEEP_SK_NT"] = nt; os.environ["FYN_DEEP_SK_SPLIT"] = split
    y = conv_gpu(x, wb, **kw).reshape(co, size, size)
    print(nt, split, hex(gpu_util.LAST_KERNEL))
    print("pixel(0,0) / (1+o/64):", y[:, 0, 0] / (1 + np.arange(co) / 64.0))
    print("chan 0:", y[0].reshape(-1))
