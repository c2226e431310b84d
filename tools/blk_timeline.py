"""Per-stage timeline of CTA 0 of the fused block kernels (pmgt_block_set_trace): cycles relative to the first stamp.
    python tools/blk_timeline.py fwd|bwd [ffn=1] [p=0.1]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pmgt_b200 import ops, _lib

which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
ffn = int(sys.argv[2]) if len(sys.argv) > 2 else 1
p = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1
T = 294912
BF16 = torch.bfloat16
r = lambda *s, k=1.0: (torch.randn(*s, device="cuda") * k).to(BF16)
w1, w2 = r(128, 128, k=0.1), r(128, 128, k=0.1)
b1, b2 = torch.randn(128, device="cuda") * 0.3, torch.randn(128, device="cuda") * 0.3
g, be = 1 + 0.1 * torch.randn(128, device="cuda"), 0.1 * torch.randn(128, device="cuda")
a, x, dy = r(T, 128), r(T, 128), r(T, 128, k=0.5)
out, dx, dz = (torch.empty_like(a) for _ in range(3))
G = [torch.zeros(128, 128, device="cuda"), torch.zeros(128, 128, device="cuda")] + [torch.zeros(128, device="cuda") for _ in range(4)]
ops.set_pdl(False)
fa = ops.block_args(a, w2, b2, g, be, 1e-12, p, 77, 14, w1=w1, b1=b1) if ffn else ops.block_args(a, w2, b2, g, be, 1e-12, p, 77, 13, res=x)
sv = ops.BlockSaved(T, bool(ffn), p, "cuda")
trace = torch.zeros(8 * 32 * 4, dtype=torch.int64, device="cuda")


def run():
    if which == "fwd":
        ops.block_fwd(fa, out, None, sv)
    elif ffn:
        ops.block_bwd(fa, sv, dy, dx, G[1], G[3], G[4], G[5], dw1=G[0], db1=G[2])
    else:
        ops.block_bwd(fa, sv, dy, dx, G[1], G[3], G[4], G[5], dz=dz)


ops.block_fwd(fa, out, None, sv)
for _ in range(3):
    run()
torch.cuda.synchronize()
_lib.lib().pmgt_block_set_trace(trace.data_ptr())
run()
torch.cuda.synchronize()
_lib.lib().pmgt_block_set_trace(None)
t = trace.view(8, 32, 4).cpu()
t0 = int(t[t > 0].min())
names = {"fwd": ["MMA(m1,m2)", "G0(wait,go,done)", "G1", "L0(wait,go,p1,done)", "L1", "res", "acc", "done"],
         "bwd": ["MMA(34,56)", "E2_0(start,p1,dofree,done)", "E2_1", "E3(wait,go,done)", "E4(wait,go,done)", "", "", ""]}[which]
for n in range(12):
    line = [f"n={n:2d}"]
    for role in range(8):
        ev = [int(v) - t0 for v in t[role, n] if int(v) > 0]
        if ev:
            line.append(f"{names[role].split('(')[0]}:" + "/".join(str(e) for e in ev))
    print("  ".join(line))
