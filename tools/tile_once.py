"""Each token-tile kernel variant twice at the bench shape (for `ncu -k regex:... ` captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmgt_b200 import ops
BF16 = torch.bfloat16
R, L, H = 49152, 6, 128
T = R * L
def r(*s, sc=1.0): return (torch.randn(*s, device="cuda") * sc).to(BF16)
x, w4, w = r(T, H), r(4 * H, H, sc=0.05), r(H, H, sc=0.05)
b4 = torch.zeros(4 * H, device="cuda"); g = torch.ones(H, device="cuda"); be = torch.zeros(H, device="cuda")
o4 = torch.empty(T, 4 * H, device="cuda", dtype=BF16)
o, o2, res = torch.empty(T, H, device="cuda", dtype=BF16), torch.empty(T, H, device="cuda", dtype=BF16), r(T, H)
dw4, db4 = torch.zeros(4 * H, H, device="cuda"), torch.zeros(4 * H, device="cuda")
dz, do = torch.empty_like(o), torch.empty_like(o)
dg, db = torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda")
for _ in range(2):
    ops.linear_tile(x, w4, o4, ops.LT_BIAS, bias=b4)
    ops.linear_tile(x, w, o, ops.LT_GELU, bias=b4[:H], aux_out=o2)
    ops.linear_tile(x, w, o, ops.LT_RES_LN, bias=b4[:H], aux_out=o2, e_in=res, ln_g=g, ln_b=be, ln_eps=1e-12, p=0.1, seed=1, site=3)
    ops.linear_tile(x, w, o, ops.LT_PLAIN, w_mn=True)
    ops.linear_tile(o4, w4, o, ops.LT_PLAIN, w_mn=True)
    ops.linear_tile(x, w, o, ops.LT_GELU_BWD, w_mn=True, e_in=res)
    ops.dw_tile(o4, x, dw4, db4)
    ops.dw_tile(res, x, dw4[:H], db4[:H])
    ops.ln_bwd(T, H, res, g, 1e-12, 0.1, 1, 3, dz, do, dg, db, dy_a=x, dy_b=o)
torch.cuda.synchronize()
