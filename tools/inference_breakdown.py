"""Where does bulk inference (BASELINE config 4) spend its time?  Times sampling, the eval forward and the D2H copy apart."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pmgt_b200 import trainer
from pmgt_b200.datasets import PMGTDataset

wl = sys.argv[1] if len(sys.argv) > 1 else "1M"
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
dev = torch.device("cuda", 0)
args = trainer.make_args(synthetic=wl, test_batch_size=bs, seed=0, mode="inference")
args.device = dev
trainer.set_seed(0)
args.graph, args.feat_init_emb = trainer._load_graph_and_features(args)
trainer.init_dataloader(args)
trainer.init_model(args)
args.model.bert.use_launch_plans = True
net = args.model
net.eval()
ds = PMGTDataset(args.graph, max_ctx_neigh=args.max_ctx_neigh, hop_sampling_sizes=args.hop_sampling_sizes,
                 is_training=False, is_inference=True, seed=args.seed)
n = len(ds)
out = torch.empty((n, args.hidden_size), dtype=torch.float32, device=dev)


def sync_time(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    return time.perf_counter() - t0, r


def sample_all():
    return [ds.sample_batch(np.arange(s, min(s + bs, n))) for s in range(0, n, bs)]


def fwd_all(batches):
    with torch.no_grad():
        for i, b in enumerate(batches):
            s = i * bs
            out[s: s + b["node_ids"].shape[0]] = net(b)[0][:, 0]


batches = sample_all()
fwd_all(batches[:2])
res = {"workload": wl, "nodes": n, "batch": bs}
res["sample_s"], batches = sync_time(sample_all)
res["forward_s"], _ = sync_time(lambda: fwd_all(batches))
res["d2h_pageable_s"], _ = sync_time(lambda: out.cpu().numpy())
pin = torch.empty(out.shape, dtype=out.dtype).pin_memory()
res["d2h_pinned_s"], _ = sync_time(lambda: pin.copy_(out, non_blocking=True))
res["whole_inference_s"], _ = sync_time(lambda: trainer.inference(args))
print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items()}))
