"""Times pmgt_sample_contexts on a synthetic graph: python tools/bench_sampler.py TG|VG|1M [n_ctx]  (PMGT_SAMPLER_GUIDE=0 for
the plain binary search)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pmgt_b200 import synthetic
from pmgt_b200.datasets import context_keys, sample_contexts

wl = sys.argv[1] if len(sys.argv) > 1 else "TG"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 45056
t0 = time.time()
g = synthetic.make_item_graph(wl, device="cuda" if wl == "1M" else None)
t_graph = time.time() - t0
dev = torch.device("cuda", 0)
t0 = time.time()
g.device_handle(0)
t_upload = time.time() - t0
rng = np.random.default_rng(0)
roots = torch.from_numpy(rng.integers(2, g.num_nodes + 2, size=n)).to(dev)
keys = context_keys(3, roots, 0)
for _ in range(3):
    sample_contexts(g, roots, keys, [16, 8, 4], 5, 0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    ids, mask = sample_contexts(g, roots, keys, [16, 8, 4], 5, 0)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(json.dumps({"workload": wl, "n_ctx": n, "guide": os.environ.get("PMGT_SAMPLER_GUIDE", "1"), "ms": round(ms, 4),
                  "ns_per_ctx": round(ms * 1e6 / n, 1), "graph_build_s": round(t_graph, 2), "upload_s": round(t_upload, 2),
                  "checksum": int(ids.sum())}))
