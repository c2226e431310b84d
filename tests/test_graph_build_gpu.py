"""Device-side item-graph ingestion (pmgt_graph_create_device + ItemGraph.from_edge_list_device) against the host
builder (ItemGraph.from_edge_list, which tests/test_host_logic.py pins to networkx adjacency order): same CSR, CDF equal
to fp32 rounding, and the sampler running on the device-built lookup tables reproduces the CPU Philox replay bit for
bit."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _edges(n, m, seed):
    from pmgt_b200 import synthetic
    u, v = synthetic.chung_lu_edges(n, m, seed)
    rng = np.random.default_rng(seed + 1)
    w = rng.uniform(0.2, 1.6, size=m)          # fp64 weights, like the python floats of the reference's nx.Graph
    return u + 2, v + 2, w


@pytest.mark.parametrize("n,m,seed", [(50, 120, 1), (3000, 40000, 2), (20000, 500000, 3)])
def test_device_builder_matches_host_builder(n, m, seed):
    from pmgt_b200.graph import ItemGraph
    u, v, w = _edges(n, m, seed)
    host = ItemGraph.from_edge_list(n, u, v, w)
    dev = ItemGraph.from_edge_list_device(n, u, v, w, device="cuda")
    assert np.array_equal(host.indptr, dev.indptr) and np.array_equal(host.indices, dev.indices)
    assert np.allclose(host.weights, dev.weights)
    # CDF: non-decreasing per row, last entry exactly 1, equal to the host's fp64 build up to fp32 rounding
    assert np.all(dev.cdf[host.indptr[1:][np.diff(host.indptr) > 0] - 1] == 1.0)
    ulp = np.spacing(np.maximum(host.cdf, dev.cdf).astype(np.float32))
    assert np.all(np.abs(host.cdf.astype(np.float64) - dev.cdf.astype(np.float64)) <= ulp), "CDF differs by more than 1 ulp"
    row_start = np.zeros(len(dev.cdf), dtype=bool)
    row_start[host.indptr[:-1][np.diff(host.indptr) > 0]] = True
    d = np.diff(dev.cdf, prepend=0.0)
    assert np.all((d >= 0) | row_start)


def test_sampler_on_device_built_graph_matches_cpu_replay():
    from oracle import sampler_ref
    from pmgt_b200.datasets import context_keys, sample_contexts
    from pmgt_b200.graph import ItemGraph
    n, m = 5000, 60000
    u, v, w = _edges(n, m, 7)
    g = ItemGraph.from_edge_list_device(n, u, v, w, device="cuda")
    roots = torch.arange(2, 2 + 512, device="cuda")
    keys = context_keys(2, roots, 0)
    ids, mask = sample_contexts(g, roots, keys, [16, 8, 4], 5, 11)
    want_ids, want_mask, _ = sampler_ref.philox_sample_contexts(g.indptr, g.indices, g.cdf, g.num_nodes, roots.cpu().numpy(),
                                                                keys.cpu().numpy(), [16, 8, 4], 5, 11)
    assert np.array_equal(ids.cpu().numpy(), want_ids) and np.array_equal(mask.cpu().numpy(), want_mask)


def test_duplicate_edges_fall_back_to_the_host_semantics():
    from pmgt_b200.graph import ItemGraph
    u = np.array([2, 3, 3, 4, 2]); v = np.array([3, 4, 2, 5, 3]); w = np.array([0.5, 1.0, 0.9, 0.2, 0.1])
    a = ItemGraph.from_edge_list(4, u, v, w)
    b = ItemGraph.from_edge_list_device(4, u, v, w, device="cuda")
    assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices) and np.array_equal(a.cdf, b.cdf)
