"""Seeded synthetic inputs of the shapes BASELINE.json names (there is no network
for the Amazon VG/TG data).  Value distributions follow the reference's data
preparation (notebooks/PMGT.ipynb cell 20: edge weight
``(ln r + 1) / (ln sqrt(deg_u * deg_v) + 1)`` with co-review count r >= 3;
cell 30: feature rows 0/1 zero, most rows N(0, 1)).
"""
from typing import Tuple

import numpy as np

from .graph import ItemGraph

# (num_nodes, num_undirected_edges, graph seed, feature seed)  -- SURVEY.md section 8(d)
SHAPES = {
    "VG": (7252, 88606, 0, 1234),
    "TG": (10834, 38252, 1, 1235),
    "1M": (1_000_000, 20_000_000, 2, 1236),
}


def _unique_pairs(u: np.ndarray, v: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Drop self loops and duplicate undirected pairs, keeping first occurrences in order."""
    keep = u != v
    u, v = u[keep], v[keep]
    lo, hi = np.minimum(u, v), np.maximum(u, v)
    code = lo.astype(np.int64) << 32 | hi.astype(np.int64)
    _, first = np.unique(code, return_index=True)
    first.sort()
    return u[first], v[first]


def chung_lu_edges(num_nodes: int, num_edges: int, seed: int, exponent: float = 2.5,
                   max_degree: int = 10_000) -> Tuple[np.ndarray, np.ndarray]:
    """Heavy-tailed undirected edge list with exactly ``num_edges`` unique edges,
    no self loops and no isolated node.  Node indices are 0-based here."""
    if num_edges < num_nodes:
        raise ValueError("need at least one edge per node")
    rng = np.random.default_rng(seed)
    rank = rng.permutation(num_nodes).astype(np.float64)
    w = (rank + 1.0) ** (-1.0 / (exponent - 1.0))
    w *= 2.0 * num_edges / w.sum()
    w = np.minimum(w, float(max_degree))
    p = w / w.sum()
    cdf = np.cumsum(p)
    cdf[-1] = 1.0

    def draw(n):
        return np.searchsorted(cdf, rng.random(n), side="right").astype(np.int64)

    # backbone: every node gets one edge to a degree-biased partner
    bu = rng.permutation(num_nodes).astype(np.int64)
    bv = draw(num_nodes)
    clash = bu == bv
    bv[clash] = (bv[clash] + 1) % num_nodes
    u, v = _unique_pairs(bu, bv)
    # nodes that lost their backbone edge to de-duplication still appear as partner of someone
    while len(u) < num_edges:
        need = num_edges - len(u)
        nu, nv = draw(int(need * 1.2) + 16), draw(int(need * 1.2) + 16)
        u, v = _unique_pairs(np.concatenate([u, nu]), np.concatenate([v, nv]))
    u, v = u[:num_edges], v[:num_edges]
    deg = np.bincount(u, minlength=num_nodes) + np.bincount(v, minlength=num_nodes)
    assert deg.min() >= 1
    return u, v


def make_edge_list(name_or_shape, seed=None):
    """``(num_nodes, src, dst, weight)`` of a synthetic item graph: node ids ``2..N+1``, unique undirected edges in
    insertion order, fp32-rounded weights per notebook cell 20 (``(ln r + 1) / (ln sqrt(deg_u deg_v) + 1)``, r >= 3)."""
    if isinstance(name_or_shape, str):
        n, m, gseed, _ = SHAPES[name_or_shape]
    else:
        n, m = name_or_shape
        gseed = 0
    if seed is not None:
        gseed = seed
    u, v = chung_lu_edges(n, m, gseed)
    rng = np.random.default_rng(gseed + 7919)
    deg = (np.bincount(u, minlength=n) + np.bincount(v, minlength=n)).astype(np.float64)
    r = 2.0 + rng.geometric(0.5, size=m)  # co-review counts, >= 3
    weight = (np.log(r) + 1.0) / (np.log(np.sqrt(deg[u] * deg[v])) + 1.0)
    return n, u + 2, v + 2, weight.astype(np.float32)


def make_item_graph(name_or_shape, seed=None, device=None) -> ItemGraph:
    """``"VG"`` / ``"TG"`` / ``"1M"`` or an explicit ``(num_nodes, num_edges)``.  With ``device`` (a CUDA device) the CSR,
    the softmax CDF and the sampler's lookup tables are built on the GPU (``ItemGraph.from_edge_list_device``: 0.7 s
    instead of 12 s at 1M nodes / 20M edges); the CDF then agrees with the host build to fp32 rounding."""
    n, src, dst, weight = make_edge_list(name_or_shape, seed)
    if device is not None:
        return ItemGraph.from_edge_list_device(n, src, dst, weight.astype(np.float64), device=device)
    return ItemGraph.from_edge_list(n, src, dst, weight)


def make_features(num_nodes: int, dims=(1536, 768), seed: int = 1234):
    """Host fp32 feature tables ``(N+2, D)``; rows 0 (<pad>) and 1 (<mask>) are zero."""
    rng = np.random.default_rng(seed)
    out = []
    for d in dims:
        t = rng.standard_normal((num_nodes + 2, d), dtype=np.float32)
        t[:2] = 0.0
        out.append(t)
    return out


def make_features_device(num_nodes: int, dims=(1536, 768), seed: int = 1234, device="cuda",
                         dtype=None, chunk_rows: int = 65536):
    """Same distribution generated directly on the device in chunks (for the 1M-node
    configuration, whose fp32 tables would be 9.2 GB on the host)."""
    import torch

    dtype = dtype or torch.bfloat16
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    out = []
    for d in dims:
        t = torch.empty((num_nodes + 2, d), dtype=dtype, device=device)
        for lo in range(0, num_nodes + 2, chunk_rows):
            hi = min(lo + chunk_rows, num_nodes + 2)
            t[lo:hi] = torch.randn((hi - lo, d), generator=gen, device=device, dtype=torch.float32).to(dtype)
        t[:2] = 0
        out.append(t)
    return out
