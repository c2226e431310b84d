#!/usr/bin/env python
"""torch.profiler view of a few real training steps: every CUDA kernel (ours and torch's), GPU busy time vs wall time.

    python tools/profile_step.py [--workload TG] [--batch 4096] [--steps 3]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pmgt_b200 import synthetic, trainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="TG")
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()

dev = torch.device("cuda", 0)
args = trainer.make_args(synthetic=a.workload, train_batch_size=a.batch, seed=0)
args.device = dev
trainer.set_seed(0)
args.graph, args.feat_init_emb = trainer._load_graph_and_features(args)
trainer.init_dataloader(args)
trainer.init_model(args)
tm = trainer.PMGTTrainerModel(args)
ds = args.train_dataset
n = len(ds)
idx = [torch.from_numpy(np.resize(trainer.epoch_permutation(n, 0, s), a.batch).astype(np.int64)).to(dev) for s in range(a.steps + 3)]
for s in range(3):
    tm.train_on_indices(ds, idx[s], epoch=s)
    tm.prefetch(ds, idx[s + 1], epoch=s + 1)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    e0.record()
    for s in range(3, 3 + a.steps):
        tm.train_on_indices(ds, idx[s], epoch=s)
        if s + 1 < 3 + a.steps:
            tm.prefetch(ds, idx[s + 1], epoch=s + 1)
    e1.record()
    torch.cuda.synchronize()
wall = e0.elapsed_time(e1) / a.steps
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
busy = sum(e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total for e in ev) / 1e3 / a.steps
print(f"wall {wall:.3f} ms/step, GPU busy {busy:.3f} ms/step, idle {wall - busy:.3f} ms/step, {len(ev) / a.steps:.0f} GPU ops/step")
agg = {}
for e in ev:
    t = e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
    d = agg.setdefault(e.name[:100], [0, 0.0])
    d[0] += 1
    d[1] += t
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{t / 1e3 / a.steps:8.3f} ms/step {c / a.steps:6.1f}x  {k}")
