"""Times the attention core at BASELINE config-5 shape (256 targets: 3072 sequences x 12 heads, L = 33, dh = 64)."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmgt_b200 import ops
R, L, H, heads = 3072, 33, 768, 12
p = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
T = R * L; BF16 = torch.bfloat16
qkvc = (torch.randn(T, 4 * H, device="cuda") * 0.7).to(BF16)
mask = torch.ones(R, L, device="cuda"); mask[::3, 20:] = 0
ctx = torch.empty(T, H, device="cuda", dtype=BF16); dctx = torch.randn(T, H, device="cuda").to(BF16)
dqkvc = torch.empty_like(qkvc)
fa = ops.attn_args(R, L, H, heads, 0.5, qkvc, mask, p, 1, 10, ctx=ctx)
ba = ops.attn_args(R, L, H, heads, 0.5, qkvc, mask, p, 1, 10, dctx=dctx, dqkvc=dqkvc)
res = {"R": R, "L": L, "H": H, "heads": heads, "p": p}
for name, fn, nbytes in (("fwd", lambda: ops.attn_core_fwd(fa), T * H * 2 * 5), ("bwd", lambda: ops.attn_core_bwd(ba), T * H * 2 * 9)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    res[name + "_us"] = round(us, 1); res[name + "_gbs"] = round(nbytes / us / 1e3)
print(json.dumps(res))
