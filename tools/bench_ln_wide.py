"""Times res_ln fwd / bwd at BASELINE config-5 width (H = 768, 101,376 tokens).  python tools/bench_ln_wide.py [p]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pmgt_b200 import ops
BF16 = torch.bfloat16
T, H = 3072 * 33, 768
p = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
o, res, dy = (torch.randn(T, H, device="cuda").to(BF16) for _ in range(3))
g, b = torch.ones(H, device="cuda"), torch.zeros(H, device="cuda")
y, dz, d_o = (torch.empty(T, H, device="cuda", dtype=BF16) for _ in range(3))
dg, db, dbias = (torch.zeros(H, device="cuda") for _ in range(3))
fa = ops.resln_args(T, H, o, res, g, b, 1e-12, p, 5, 7, y=y)
ba = ops.resln_args(T, H, o, res, g, None, 1e-12, p, 5, 7, dy=dy, dz=dz, d_o=d_o, d_g=dg, d_b=db, d_bias=dbias)
out = {"T": T, "H": H, "p": p}
for name, fn, rows in (("fwd", lambda: ops.res_ln_fwd(fa), 3), ("bwd", lambda: ops.res_ln_bwd(ba), 5)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    out[name + "_us"] = round(us, 1); out[name + "_gbs"] = round(rows * T * H * 2 / us / 1e3)
print(json.dumps(out))
