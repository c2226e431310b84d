import os, sys, torch
sys.path.insert(0, os.getcwd())
from pmgt_b200 import ops
R, L, H = 49152, 6, 128; T = R * L; BF16 = torch.bfloat16
qkvc = (torch.randn(T, 4 * H, device="cuda") * 0.7).to(BF16)
mask = torch.ones(R, L, device="cuda"); mask[::3, 4:] = 0
ctx = torch.empty(T, H, device="cuda", dtype=BF16); dctx = torch.randn(T, H, device="cuda").to(BF16)
dqkvc = torch.empty_like(qkvc)
for _ in range(2):
    ops.attn_core_fwd(ops.attn_args(R, L, H, 1, 0.5, qkvc, mask, 0.1, 1, 10, ctx=ctx))
    ops.attn_core_bwd(ops.attn_args(R, L, H, 1, 0.5, qkvc, mask, 0.1, 1, 10, dctx=dctx, dqkvc=dqkvc))
torch.cuda.synchronize()
