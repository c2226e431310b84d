"""Times pmgt_gemm_bf16 on the Linear shapes of BASELINE config 5 (H = 768, I = 3072, 256 targets: 101,376 tokens) and
checks each result against torch on a slice.  python tools/bench_gemm_wide.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pmgt_b200 import ops

BF16 = torch.bfloat16
T = 3072 * 33
torch.manual_seed(0)
res = {"T": T}


def timeit(fn, n=10):
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


for name, N, K, gelu in (("qkvc", 3072, 768, False), ("out", 768, 768, False), ("ffn1_gelu", 3072, 768, True), ("ffn2", 768, 3072, False)):
    x = torch.randn(T, K, device="cuda").to(BF16)
    w = (torch.randn(N, K, device="cuda") * 0.03).to(BF16)
    b = torch.randn(N, device="cuda") * 0.1
    out = torch.empty(T, N, device="cuda", dtype=BF16)
    aux = torch.empty(T, N, device="cuda", dtype=BF16) if gelu else None
    us = timeit(lambda: ops.linear_fwd(x, w, b, out, gelu_aux=aux))
    ref = x[:300].float() @ w.float().t() + b
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    err = float((out[:300].float() - ref).abs().max() / ref.abs().max())
    ref2 = x[-200:].float() @ w.float().t() + b
    if gelu:
        ref2 = torch.nn.functional.gelu(ref2)
    err = max(err, float((out[-200:].float() - ref2).abs().max() / ref2.abs().max()))
    res[f"fwd_{name}_us"] = round(us, 1)
    res[f"fwd_{name}_tflops"] = round(2 * T * N * K / us / 1e6, 1)
    res[f"fwd_{name}_err"] = round(err, 5)
    # dX = dY W
    dy = torch.randn(T, N, device="cuda").to(BF16)
    dx = torch.empty(T, K, device="cuda", dtype=BF16)
    us = timeit(lambda: ops.linear_dx(dy, w, dx))
    ref = dy[:300].float() @ w.float()
    res[f"dx_{name}_us"] = round(us, 1)
    res[f"dx_{name}_tflops"] = round(2 * T * N * K / us / 1e6, 1)
    res[f"dx_{name}_err"] = round(float((dx[:300].float() - ref).abs().max() / ref.abs().max()), 5)
    del x, w, out, aux, dy, dx
print(json.dumps(res))
