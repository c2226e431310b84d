"""Diagnostic (not a test): runs every GEMM operand-layout variant once and prints
error statistics, so a single GPU call tells which variant (if any) is wrong."""
import os
import sys
import traceback

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from pmgt_b200 import ops  # noqa: E402

BF16 = torch.bfloat16
torch.manual_seed(0)


def report(name, got, want):
    got = got.float()
    scale = float(want.abs().max())
    err = float((got - want).abs().max())
    bad = int(((got - want).abs() > 2e-2 * scale).sum())
    print(f"{name:40s} max_err={err:.4g} scale={scale:.4g} bad={bad}/{want.numel()} finite={bool(torch.isfinite(got).all())}",
          flush=True)
    if bad:
        idx = ((got - want).abs() > 2e-2 * scale).nonzero()[:6].tolist()
        print("   first bad:", [(i, float(got[tuple(i)]), float(want[tuple(i)])) for i in idx], flush=True)


def main():
    M, N, K = 256, 128, 128
    x = torch.randn(M, K, device="cuda").to(BF16)
    w = (torch.randn(N, K, device="cuda") * 0.1).to(BF16)
    out = torch.empty(M, N, device="cuda", dtype=BF16)
    try:
        ops.linear_fwd(x, w, None, out); torch.cuda.synchronize()
        report("fwd K-major/K-major 256x128x128", out, x.float() @ w.float().t())
    except Exception:
        traceback.print_exc()
    try:
        dy = torch.randn(M, N, device="cuda").to(BF16)
        o2 = torch.empty(M, K, device="cuda", dtype=BF16)
        ops.linear_dx(dy, w, o2); torch.cuda.synchronize()
        report("dx  K-major/MN-major", o2, dy.float() @ w.float())
    except Exception:
        traceback.print_exc()
    try:
        dw = torch.zeros(N, K, device="cuda")
        ops.linear_dw(dy, x, dw); torch.cuda.synchronize()
        report("dw  MN-major/MN-major split-k", dw, dy.float().t() @ x.float())
    except Exception:
        traceback.print_exc()
    try:
        table = torch.randn(100, 256, device="cuda").to(BF16)
        idx = torch.randint(0, 100, (M,), device="cuda")
        w2 = (torch.randn(N, 256, device="cuda") * 0.1).to(BF16)
        ops.linear_fwd(table, w2, None, out, rows=idx, src_rows=100); torch.cuda.synchronize()
        report("fwd gathered A", out, table[idx].float() @ w2.float().t())
    except Exception:
        traceback.print_exc()
    try:
        dw2 = torch.zeros(N, 256, device="cuda")
        ops.linear_dw(dy, table, dw2, rows=idx, src_rows=100, x_cols=256); torch.cuda.synchronize()
        report("dw  gathered B", dw2, dy.float().t() @ table[idx].float())
    except Exception:
        traceback.print_exc()


if __name__ == "__main__":
    main()
