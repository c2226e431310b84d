"""BASELINE config 4: bulk item-embedding inference (forward only, eval mode) over all nodes of a synthetic graph through
`trainer.inference` (sample context -> gather -> encode -> write [:, 0]); prints one JSON line."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.getcwd())
from pmgt_b200 import trainer

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="1M")
ap.add_argument("--batch", type=int, default=32768)
a = ap.parse_args()
dev = torch.device("cuda", 0)
args = trainer.make_args(synthetic=a.workload, test_batch_size=a.batch, seed=0, mode="inference")
args.device = dev
trainer.set_seed(0)
args.graph, args.feat_init_emb = trainer._load_graph_and_features(args)
trainer.init_dataloader(args)
trainer.init_model(args)
args.model.bert.use_launch_plans = True
trainer.inference(args)  # warm-up pass: graph upload, launch-plan recording, lazy module loading
torch.cuda.synchronize()
t0 = time.perf_counter()
emb = trainer.inference(args)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
n = emb.shape[0]
print(json.dumps({"metric": "pmgt_inference_node_contexts_per_s", "value": n / dt, "unit": "contexts/s", "nodes": n,
                  "seconds": dt, "batch": a.batch, "workload": a.workload, "hidden": int(emb.shape[1]),
                  "includes": "sampling, encode, D2H of the (N, H) fp32 result", "finite": bool(abs(emb).max() < 1e4)}))
