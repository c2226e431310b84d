"""Fixed cost of the dW kernels: time pmgt_dw_tile / fused dX+dW for 1, 2, 4, 8, 16 tiles per CTA; the intercept of the
line is prologue + TMEM flush (148-way same-address red.add) + tail."""
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
from pmgt_b200 import ops

BF16 = torch.bfloat16
dev = torch.device("cuda", 0)
sms = torch.cuda.get_device_properties(0).multi_processor_count


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


ops.set_pdl(False)
for N in (128, 512):
    for tiles in (1, 2, 4, 8, 16):
        T = sms * 128 * tiles
        dy = torch.randn(T, N, device=dev).to(BF16)
        x = torch.randn(T, 128, device=dev).to(BF16)
        dw = torch.zeros(N, 128, device=dev)
        db = torch.zeros(N, device=dev)
        us = timeit(lambda: ops.dw_tile(dy, x, dw, db))
        print(f"dw_tile N={N:3d} tiles/CTA={tiles:2d}  {us:7.2f} us   {2 * T * (N + 128) / us / 1e6:6.2f} TB/s")
w = torch.randn(128, 128, device=dev).to(BF16)
for tiles in (1, 2, 4, 8, 16):
    T = sms * 128 * tiles
    dy = torch.randn(T, 128, device=dev).to(BF16)
    x = torch.randn(T, 128, device=dev).to(BF16)
    out = torch.empty(T, 128, device=dev, dtype=BF16)
    dw = torch.zeros(128, 128, device=dev)
    db = torch.zeros(128, device=dev)
    us = timeit(lambda: ops.linear_tile(dy, w, out, ops.LT_PLAIN, w_mn=True, dw_x=x, dw=dw, dbias=db))
    us2 = timeit(lambda: ops.linear_tile(dy, w, out, ops.LT_PLAIN, w_mn=True))
    print(f"dxdw tiles/CTA={tiles:2d}  fused {us:7.2f} us   plain dX {us2:7.2f} us")
