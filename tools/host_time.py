"""Host-side cost of a step in the end-to-end loop (sync every step): cProfile of `train_on_indices` + `prefetch`."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.getcwd())
from pmgt_b200 import trainer

dev = torch.device("cuda", 0)
WL = sys.argv[1] if len(sys.argv) > 1 else "TG"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
args = trainer.make_args(synthetic=WL, train_batch_size=B, seed=0)
args.device = dev
trainer.set_seed(0)
args.graph, args.feat_init_emb = trainer._load_graph_and_features(args)
trainer.init_dataloader(args)
trainer.init_model(args)
tm = trainer.PMGTTrainerModel(args)
ds = args.train_dataset
N = 40
idx = [torch.from_numpy(np.resize(trainer.epoch_permutation(len(ds), 0, s), B).astype(np.int64)).pin_memory() for s in range(N + 1)]


def loop(lo, hi):
    for s in range(lo, hi):
        loss = tm.train_on_indices(ds, idx[s], epoch=s)
        tm.prefetch(ds, idx[s + 1], epoch=s + 1)
        float(loss)


loop(0, 5)
torch.cuda.synchronize()
t0 = time.perf_counter()
loop(5, 20)
torch.cuda.synchronize()
print(f"e2e-style loop: {1e3 * (time.perf_counter() - t0) / 15:.3f} ms/step")
pr = cProfile.Profile()
pr.enable()
loop(20, 35)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
st.sort_stats("cumulative").print_stats(30)
