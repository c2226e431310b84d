"""How much does the side-stream batch preparation cost the main stream?  Times the 1M-graph step (a) as bench.py runs it
(sampling + NFR corruption of step k + 1 on the side stream while step k runs) and (b) with ONE batch prepared up front
and re-used every step (no sampler, no corruption: the main chain alone, PDL on)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pmgt_b200 import trainer

wl = sys.argv[1] if len(sys.argv) > 1 else "1M"
B = 4096
dev = torch.device("cuda", 0)
args = trainer.make_args(synthetic=wl, train_batch_size=B, seed=0)
args.device = dev
trainer.set_seed(0)
args.graph, args.feat_init_emb = trainer._load_graph_and_features(args)
trainer.init_dataloader(args)
trainer.init_model(args)
tm = trainer.PMGTTrainerModel(args)
ds = args.train_dataset
n = len(ds)
N = 40
idx = [torch.from_numpy(np.resize(trainer.epoch_permutation(n, 0, s), B).astype(np.int64)).to(dev) for s in range(N + 10)]


def timed(fn, steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        fn(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def normal(s):
    tm.train_on_indices(ds, idx[s], epoch=s)
    tm.prefetch(ds, idx[s + 1], epoch=s + 1)


for s in range(8):
    normal(s)
res = {"workload": wl, "with_side_stream_ms": round(timed(lambda s: normal(s + 8), N - 8), 3)}

# one prepared batch, re-used: train_on_indices finds it "prefetched" every step
batch = ds.sample_batch(idx[0], epoch=0)
masked = tm.net.mask_nodes(batch[0]["node_ids"], with_positions=True)
ready = torch.cuda.Event()
ready.record()
torch.cuda.synchronize()


def reuse(s):
    tm._prefetched = (ds, idx[0], 0, batch, masked, ready)
    tm.train_on_indices(ds, idx[0], epoch=0)


for s in range(5):
    reuse(s)
res["main_chain_only_ms"] = round(timed(reuse, N - 8), 3)
print(json.dumps(res))
