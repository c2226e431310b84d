"""Item-graph ingestion at BASELINE config 3's size (1M nodes / 20M undirected edges): host builder (numpy lexsort + fp64
softmax CDF + C++ guide tables + upload) vs the device builder (one stable device sort + CDF / guide kernels)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pmgt_b200 import synthetic
from pmgt_b200.graph import ItemGraph

name = sys.argv[1] if len(sys.argv) > 1 else "1M"
n, m, gseed, _ = synthetic.SHAPES[name]
t0 = time.time()
u, v = synthetic.chung_lu_edges(n, m, gseed)
rng = np.random.default_rng(gseed + 7919)
deg = (np.bincount(u, minlength=n) + np.bincount(v, minlength=n)).astype(np.float64)
r = 2.0 + rng.geometric(0.5, size=m)
w = (np.log(r) + 1.0) / (np.log(np.sqrt(deg[u] * deg[v])) + 1.0)
t_edges = time.time() - t0
torch.cuda.init()
torch.zeros(1, device="cuda")
t0 = time.time()
gh = ItemGraph.from_edge_list(n, u + 2, v + 2, w)
t_host_csr = time.time() - t0
t0 = time.time()
gh.device_handle(0)
torch.cuda.synchronize()
t_host_upload = time.time() - t0
ud, vd, wd = torch.from_numpy(u + 2).cuda(), torch.from_numpy(v + 2).cuda(), torch.from_numpy(w).cuda()
torch.cuda.synchronize()
t0 = time.time()
gd = ItemGraph.from_edge_list_device(n, ud, vd, wd, device="cuda")
torch.cuda.synchronize()
t_dev = time.time() - t0
same = bool(np.array_equal(gh.indptr, gd.indptr) and np.array_equal(gh.indices, gd.indices))
print(json.dumps({"graph": name, "nodes": n, "edges": m, "edge_list_generation_s": round(t_edges, 2),
                  "host_builder_s": round(t_host_csr, 2), "host_lookup_tables_and_upload_s": round(t_host_upload, 2),
                  "device_builder_s_incl_readback": round(t_dev, 2), "same_csr": same,
                  "max_cdf_diff": float(np.abs(gh.cdf.astype(np.float64) - gd.cdf.astype(np.float64)).max())}))
