#!/usr/bin/env python
"""Micro-benchmarks of single kernels at the bench shape (R = 49152 sequences, L = 6, H = I = 128), CUDA-event timed.

    python tools/bench_kernels.py [names...]      # attn colsum resln ...

Prints one line per kernel: average us, algorithmic GB/s, fraction of the measured HBM peak.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pmgt_b200 import ops  # noqa: E402

BF16 = torch.bfloat16
R, L, H = 49152, 6, 128
T = R * L
PEAK = 6540.8
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, nbytes, name, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tot = 0.0
    for _ in range(iters):
        flush.zero_()  # L2 flush between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    us = tot / iters * 1e3
    gbs = nbytes / (us * 1e-6) / 1e9
    print(f"{name:28s} {us:9.1f} us  {gbs:8.1f} GB/s  {gbs / PEAK:6.3f} of measured HBM peak", flush=True)
    return us


def r(*shape, s=1.0):
    return (torch.randn(*shape, device="cuda") * s).to(BF16)


def bench_attn(p=0.1):
    qkvc = r(T, 4 * H, s=0.7)
    mask = torch.ones(R, L, device="cuda")
    mask[::3, 4:] = 0
    ctx = torch.empty(T, H, device="cuda", dtype=BF16)
    dctx = r(T, H)
    dqkvc = torch.empty_like(qkvc)
    dbias = torch.zeros(4 * H, device="cuda")
    timeit(lambda: ops.attn_core_fwd(ops.attn_args(R, L, H, 1, 0.5, qkvc, mask, p, 1, 10, ctx=ctx)), T * H * 2 * 5,
           f"attn_core_fwd p={p}")
    timeit(lambda: ops.attn_core_bwd(ops.attn_args(R, L, H, 1, 0.5, qkvc, mask, p, 1, 10, dctx=dctx, dqkvc=dqkvc)),
           T * H * 2 * 9, f"attn_core_bwd p={p}")
    timeit(lambda: ops.colsum(dqkvc, dbias), T * H * 2 * 4, "colsum [T,4H]")
    timeit(lambda: ops.colsum(dctx, dbias[:H]), T * H * 2, "colsum [T,H]")


def bench_resln(p=0.1):
    o, res = r(T, H), r(T, H)
    g = 1 + 0.1 * torch.randn(H, device="cuda")
    b = 0.1 * torch.randn(H, device="cuda")
    y = torch.empty(T, H, device="cuda", dtype=BF16)
    dy = r(T, H)
    dz, do = torch.empty_like(y), torch.empty_like(y)
    dg, db, dbias = (torch.zeros(H, device="cuda") for _ in range(3))
    timeit(lambda: ops.res_ln_fwd(ops.resln_args(T, H, o, res, g, b, 1e-12, p, 1, 13, y=y)), T * H * 2 * 3,
           f"res_ln_fwd p={p}")
    timeit(lambda: ops.res_ln_bwd(ops.resln_args(T, H, o, res, g, None, 1e-12, p, 1, 13, dy=dy, dz=dz, d_o=do, d_g=dg,
                                                 d_b=db, d_bias=dbias)), T * H * 2 * 5, f"res_ln_bwd p={p}")


def bench_gemm():
    x = r(T, H)
    w = r(4 * H, H, s=0.05)
    bias = torch.zeros(4 * H, device="cuda")
    out = torch.empty(T, 4 * H, device="cuda", dtype=BF16)
    timeit(lambda: ops.linear_fwd(x, w, bias, out), T * H * 2 * 5, "gemm qkvc fwd [T,128]x[512,128]")
    w1 = r(H, H, s=0.05)
    o1 = torch.empty(T, H, device="cuda", dtype=BF16)
    timeit(lambda: ops.linear_fwd(x, w1, bias[:H], o1), T * H * 2 * 2, "gemm fwd [T,128]x[128,128]")
    timeit(lambda: ops.linear_dx(out, w, o1), T * H * 2 * 5, "gemm dx [T,512]x[512,128]")
    dw = torch.zeros(4 * H, H, device="cuda")
    timeit(lambda: ops.linear_dw(out, x, dw), T * H * 2 * 5, "gemm dw [512,T]x[T,128]")


def bench_tile(p=0.1):
    x = r(T, H)
    w4 = r(4 * H, H, s=0.05)
    w = r(H, H, s=0.05)
    b4 = torch.zeros(4 * H, device="cuda")
    g = 1 + 0.1 * torch.randn(H, device="cuda")
    be = 0.1 * torch.randn(H, device="cuda")
    o4 = torch.empty(T, 4 * H, device="cuda", dtype=BF16)
    o, o2, res = torch.empty(T, H, device="cuda", dtype=BF16), torch.empty(T, H, device="cuda", dtype=BF16), r(T, H)
    B = T * H * 2
    timeit(lambda: ops.linear_tile(x, w4, o4, ops.LT_BIAS, bias=b4), 5 * B, "tile qkvc fwd (BIAS N=512)")
    timeit(lambda: ops.linear_tile(x, w, o, ops.LT_BIAS, bias=b4[:H]), 2 * B, "tile BIAS N=128")
    timeit(lambda: ops.linear_tile(x, w, o, ops.LT_GELU, bias=b4[:H], aux_out=o2), 3 * B, "tile GELU")
    timeit(lambda: ops.linear_tile(x, w, o, ops.LT_RES_LN, bias=b4[:H], aux_out=o2, e_in=res, ln_g=g, ln_b=be,
                                   ln_eps=1e-12, p=p, seed=1, site=3), 4 * B, f"tile RES_LN p={p}")
    timeit(lambda: ops.linear_tile(x, w, o, ops.LT_RES_LN, bias=b4[:H], aux_out=o2, e_in=res, ln_g=g, ln_b=be,
                                   ln_eps=1e-12, p=0.0, seed=1, site=3), 4 * B, "tile RES_LN p=0")
    timeit(lambda: ops.linear_tile(x, w, o, ops.LT_PLAIN, w_mn=True), 2 * B, "tile dX K=128")
    timeit(lambda: ops.linear_tile(o4, w4, o, ops.LT_PLAIN, w_mn=True), 5 * B, "tile dX K=512")
    timeit(lambda: ops.linear_tile(x, w, o, ops.LT_GELU_BWD, w_mn=True, e_in=res), 3 * B, "tile GELU_BWD")
    dw4, db4 = torch.zeros(4 * H, H, device="cuda"), torch.zeros(4 * H, device="cuda")
    timeit(lambda: ops.dw_tile(o4, x, dw4, db4), 5 * B, "tile dW N=512 + dbias")
    timeit(lambda: ops.dw_tile(res, x, dw4[:H], db4[:H]), 2 * B, "tile dW N=128 + dbias")
    dz, do = torch.empty_like(o), torch.empty_like(o)
    dg, db = torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda")
    timeit(lambda: ops.ln_bwd(T, H, res, g, 1e-12, p, 1, 3, dz, do, dg, db, dy_a=x, dy_b=o), 5 * B, f"ln_bwd p={p} (2 dy)")
    timeit(lambda: ops.ln_bwd(T, H, res, g, 1e-12, 0.0, 1, 3, dz, dz, dg, db, dy_a=x), 3 * B, "ln_bwd p=0 (1 dy)")


ALL = {"tile": bench_tile, "attn": bench_attn, "resln": bench_resln, "gemm": bench_gemm}

if __name__ == "__main__":
    names = sys.argv[1:] or list(ALL)
    for n in names:
        if n in ALL:
            ALL[n]()
        else:
            mod, fn = n.rsplit(".", 1) if "." in n else ("__main__", n)
            globals()[fn]()
