"""Times the gather-fused projection GEMMs of the 1M-node configuration (forward: tokens x D gathered rows -> [T,128];
backward: dW[128, D] += dY^T X_gathered) on a synthetic 1M-row table.  python tools/bench_gather.py [T] [rows]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pmgt_b200 import ops

BF16 = torch.bfloat16
T = int(sys.argv[1]) if len(sys.argv) > 1 else 294912
ROWS = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_002
torch.manual_seed(0)
res = {"T": T, "rows": ROWS}
ops.set_pdl(False)
ops.GATHER_PROJ = os.environ.get("GATHER_PROJ", "1") != "0"
res["gather_proj"] = ops.GATHER_PROJ
for name, D in (("visual", 1536), ("text", 768)):
    table = torch.randn(ROWS, D, device="cuda", dtype=torch.float32).to(BF16)
    # ids with the duplication of a real step (~68 % unique): a mix of uniform and Zipf-like hot rows
    hot = (torch.rand(T, device="cuda") ** 4 * ROWS).long().clamp_(2, ROWS - 1)
    uni = torch.randint(2, ROWS, (T,), device="cuda")
    ids = torch.where(torch.rand(T, device="cuda") < 0.5, hot, uni)
    res[name + "_unique_frac"] = round(ids.unique().numel() / T, 3)
    w = (torch.randn(128, D, device="cuda") * 0.05).to(BF16)
    b = torch.zeros(128, device="cuda")
    out = torch.empty(T, 128, device="cuda", dtype=BF16)
    dy = torch.randn(T, 128, device="cuda").to(BF16)
    dw = torch.zeros(128, D, device="cuda")

    def fwd():
        ops.linear_fwd(table, w, b, out, rows=ids, src_rows=ROWS)

    def bwd():
        ops.linear_dw(dy, table, dw, rows=ids, src_rows=ROWS, x_cols=D)

    for tag, fn, nbytes in (("fwd", fwd, T * D * 2 + T * 256), ("dw", bwd, T * D * 2 + T * 256)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        res[f"{name}_{tag}_us"] = round(us, 1)
        res[f"{name}_{tag}_gbs"] = round(nbytes / us / 1e3, 0)
    # correctness spot check on a slice
    ref = table[ids[:512]].float() @ w.float().t()
    res[name + "_fwd_err"] = float((out[:512].float() - ref).abs().max() / ref.abs().max())
    del table
print(json.dumps(res))
