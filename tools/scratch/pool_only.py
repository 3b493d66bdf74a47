import sys
sys.path.insert(0, "/root/repo")
from fyusenet_b200 import capi
ctx = capi.Context(0)
B = 128
tin = ctx.tensor(112, 112, 64, 1, capi.ORDER_DEEP, capi.F16, B)
tout = ctx.tensor(56, 56, 64, 1, capi.ORDER_DEEP, capi.F16, B)
op = capi.Pool2d(ctx, width=112, height=112, channels=64, pool=3, downsample=2, in_padding=1, out_padding=1, is_max=True, flags=capi.FLAG_DEEP | capi.FLAG_PRE_RELU)
for _ in range(3):
    op.run(tin, tout)
ctx.stream_sync()
