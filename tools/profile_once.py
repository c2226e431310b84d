#!/usr/bin/env python
"""One pre-training step inside a cudaProfilerStart/Stop window, for `ncu --profile-from-start off --set full`.

    ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/<tag>_full \
        python tools/profile_once.py [--workload TG] [--batch 4096] [--layers 1]

With --layers 1 every kernel family of the step appears once or twice (45 launches instead of 210), which keeps the
capture short; tile shapes, token counts and grids are those of the bench workload.
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pmgt_b200 import trainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="TG")
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--layers", type=int, default=1)
ap.add_argument("--warmup", type=int, default=2)
a = ap.parse_args()

dev = torch.device("cuda", 0)
args = trainer.make_args(synthetic=a.workload, train_batch_size=a.batch, seed=0, num_hidden_layers=a.layers)
args.device = dev
trainer.set_seed(0)
args.graph, args.feat_init_emb = trainer._load_graph_and_features(args)
trainer.init_dataloader(args)
trainer.init_model(args)
tm = trainer.PMGTTrainerModel(args)
ds = args.train_dataset
n = len(ds)
idx = [torch.from_numpy(np.resize(trainer.epoch_permutation(n, 0, s), a.batch).astype(np.int64)).to(dev)
       for s in range(a.warmup + 1)]
for s in range(a.warmup):
    tm.train_on_indices(ds, idx[s], epoch=s)
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = tm.train_on_indices(ds, idx[a.warmup], epoch=a.warmup)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss))
