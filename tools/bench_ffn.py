"""Times the fused post-attention block kernels against the unfused token-tile chain at the bench token count."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pmgt_b200 import ops

BF16 = torch.bfloat16
T = int(sys.argv[1]) if len(sys.argv) > 1 else 294912
p = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
torch.manual_seed(0)
r = lambda *s, k=1.0: (torch.randn(*s, device="cuda") * k).to(BF16)
w1, w2 = r(128, 128, k=0.1), r(128, 128, k=0.1)
b1, b2 = torch.randn(128, device="cuda") * 0.3, torch.randn(128, device="cuda") * 0.3
g, be = 1 + 0.1 * torch.randn(128, device="cuda"), 0.1 * torch.randn(128, device="cuda")
a, x, dy = r(T, 128), r(T, 128), r(T, 128, k=0.5)
out, da, dzd = (torch.empty_like(a) for _ in range(3))
h_pre, h, z, y, dz, do, dh_pre, da2 = (torch.empty_like(a) for _ in range(8))
G = [torch.zeros(128, 128, device="cuda"), torch.zeros(128, 128, device="cuda")] + [torch.zeros(128, device="cuda") for _ in range(4)]
ops.set_pdl(False)
fa = ops.block_args(a, w2, b2, g, be, 1e-12, p, 77, 14, w1=w1, b1=b1)
fd = ops.block_args(a, w2, b2, g, be, 1e-12, p, 77, 13, res=x)
sv, svd = ops.BlockSaved(T, True, p, "cuda"), ops.BlockSaved(T, False, p, "cuda")


def fused_fwd():
    ops.block_fwd(fa, out, None, sv)


def fused_bwd():
    ops.block_bwd(fa, sv, dy, da, G[1], G[3], G[4], G[5], dw1=G[0], db1=G[2])


def dense_fwd():
    ops.block_fwd(fd, out, None, svd)


def dense_bwd():
    ops.block_bwd(fd, svd, dy, da, G[1], G[3], G[4], G[5], dz=dzd)


def chain_fwd():
    ops.linear_tile(a, w1, h, ops.LT_GELU, bias=b1, aux_out=h_pre)
    ops.linear_tile(h, w2, y, ops.LT_RES_LN, bias=b2, aux_out=z, e_in=a, ln_g=g, ln_b=be, ln_eps=1e-12, p=p, seed=77, site=14)


def chain_bwd():
    ops.ln_bwd(T, 128, z, g, 1e-12, p, 77, 14, dz, do, G[4], G[5], dy_a=dy)
    ops.linear_tile(do, w2, dh_pre, ops.LT_GELU_BWD, w_mn=True, e_in=h_pre, dw_x=h, dw=G[1], dbias=G[3])
    ops.linear_tile(dh_pre, w1, da2, ops.LT_PLAIN, w_mn=True, dw_x=a, dw=G[0], dbias=G[2])


def chain_dense_fwd():
    ops.linear_tile(a, w2, y, ops.LT_RES_LN, bias=b2, aux_out=z, e_in=x, ln_g=g, ln_b=be, ln_eps=1e-12, p=p, seed=77, site=13)


def chain_dense_bwd():
    ops.ln_bwd(T, 128, z, g, 1e-12, p, 77, 13, dz, do, G[4], G[5], dy_a=dy)
    ops.linear_tile(do, w2, da2, ops.LT_PLAIN, w_mn=True, dw_x=a, dw=G[1], dbias=G[3])


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


res = {"T": T, "p": p}
only = sys.argv[3].split(",") if len(sys.argv) > 3 else None
for name, fn in (("chain_fwd", chain_fwd), ("fused_fwd", fused_fwd), ("chain_bwd", chain_bwd), ("fused_bwd", fused_bwd),
                 ("chain_dense_fwd", chain_dense_fwd), ("dense_fwd", dense_fwd), ("chain_dense_bwd", chain_dense_bwd),
                 ("dense_bwd", dense_bwd)):
    if only is None or name in only:
        res[name + "_us"] = round(timeit(fn), 2)
print(json.dumps(res))
