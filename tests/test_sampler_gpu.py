"""K1 parity: the CUDA sampler must reproduce the CPU Philox replay
(oracle/philox_sampler.c) bit for bit -- ids, masks, pair nodes, labels."""
import numpy as np
import pytest
import torch

from oracle import sampler_ref
from pmgt_b200 import synthetic
from pmgt_b200.datasets import PMGTDataset, context_keys, pmgt_collate_fn, sample_contexts
from pmgt_b200.graph import ItemGraph

pytestmark = pytest.mark.gpu


def _small_graph(n=300, m=1500, seed=4):
    return synthetic.make_item_graph((n, m), seed=seed)


def _run_both(g, roots, keys, hops, max_ctx, seed):
    dev = torch.device("cuda", 0)
    r = torch.as_tensor(roots, dtype=torch.int64, device=dev)
    k = torch.as_tensor(keys, dtype=torch.int64, device=dev)
    ids, mask, vdeg = sample_contexts(g, r, k, hops, max_ctx, seed, want_visited_deg=True)
    want = sampler_ref.philox_sample_contexts(g.indptr, g.indices, g.cdf, g.num_nodes, roots, keys, hops, max_ctx, seed)
    return (ids.cpu().numpy(), mask.cpu().numpy(), vdeg.cpu().numpy()), want


@pytest.mark.parametrize("hops,max_ctx", [([16, 8, 4], 5), ([4, 3, 2], 5), ([16, 8, 4], 32), ([5], 3), ([2, 2, 2, 2], 7)])
def test_contexts_bit_exact_vs_cpu_replay(hops, max_ctx):
    g = _small_graph()
    roots = np.arange(2, g.num_nodes + 2, dtype=np.int64)
    keys = (np.arange(len(roots), dtype=np.int64) << 8) | 3
    got, want = _run_both(g, roots, keys, hops, max_ctx, seed=0x1234_5678_9ABC)
    for a, b, name in zip(got, want, ("ids", "mask", "visited_deg")):
        assert np.array_equal(a, b), name
    assert got[0].dtype == np.int64 and got[1].dtype == np.float32


@pytest.mark.parametrize("hops,max_ctx", [([16, 8, 4], 5), ([6, 5, 4, 3], 7)])
def test_contexts_bit_exact_on_a_graph_larger_than_l2(hops, max_ctx):
    """A CSR above the L2-residency threshold (> 48 MB) switches the late hops to the lock-step variant (four draws of a
    Philox block advanced together); it must replay bit for bit like the sequential one."""
    g = synthetic.make_item_graph((150_000, 3_300_000), seed=11)   # 6.6 M directed entries -> ~54 MB of CSR
    assert (len(g.indices) * 8 + g.num_nodes * 8) / 1e6 > 48.0
    rng = np.random.default_rng(5)
    roots = rng.integers(2, g.num_nodes + 2, size=600).astype(np.int64)
    roots[:3] = [0, 1, g.num_nodes + 5]                            # invalid roots -> all-pad contexts
    keys = (np.arange(len(roots), dtype=np.int64) << 8) | 1
    got, want = _run_both(g, roots, keys, hops, max_ctx, seed=0xABCDEF)
    for a, b, name in zip(got, want, ("ids", "mask", "visited_deg")):
        assert np.array_equal(a, b), name


def test_contexts_ragged_rows_and_padding():
    """TG-like sparse graph: many rows shorter than a hop's sample size -> padded contexts."""
    g = synthetic.make_item_graph((400, 420), seed=9)  # mean degree ~2
    roots = np.arange(2, g.num_nodes + 2, dtype=np.int64)
    keys = np.arange(len(roots), dtype=np.int64) * 977
    got, want = _run_both(g, roots, keys, [16, 8, 4], 5, seed=7)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    assert (got[1] == 0).any(), "expected at least one padded context on a sparse graph"
    assert np.array_equal(got[0] == 0, got[1] == 0)


def test_contexts_invalid_and_isolated_roots():
    """Root ids outside [2, N+2) and isolated nodes give an all-pad context (mask 1 only at position 0)."""
    n = 20
    src = np.arange(2, n + 1)
    dst = src + 1
    indptr_graph = ItemGraph.from_edge_list(n + 1, src, dst, np.ones(len(src)))  # node n+2 is isolated
    roots = np.asarray([0, 1, n + 2, 5], dtype=np.int64)
    keys = np.arange(4, dtype=np.int64)
    got, want = _run_both(indptr_graph, roots, keys, [4, 2], 3, seed=1)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    assert np.all(got[0][:3, 1:] == 0) and np.all(got[1][:3, 1:] == 0)


@pytest.mark.parametrize("max_pos,min_neg,max_total", [(5, 5, 10), (1, 1, 2), (3, 8, 10)])
def test_pairs_bit_exact_vs_cpu_replay(max_pos, min_neg, max_total):
    from pmgt_b200 import ops

    g = _small_graph()
    dev = torch.device("cuda", 0)
    targets = np.arange(2, g.num_nodes + 2, dtype=np.int64)
    keys = np.arange(len(targets), dtype=np.int64) << 8
    P = max(max_pos + min_neg, max_total)
    pairs = torch.empty((len(targets), P), dtype=torch.int64, device=dev)
    labels = torch.empty((len(targets), P), dtype=torch.float32, device=dev)
    num = torch.empty(len(targets), dtype=torch.int64, device=dev)
    ops.sample_pairs(g.device_handle(0), torch.as_tensor(targets, device=dev), torch.as_tensor(keys, device=dev),
                     max_pos, min_neg, max_total, P, 99, pairs, labels, num)
    w_pairs, w_labels, w_num = sampler_ref.philox_sample_pairs(g.indptr, g.indices, g.num_nodes, targets, keys,
                                                               max_pos, min_neg, max_total, P, 99)
    assert np.array_equal(pairs.cpu().numpy(), w_pairs)
    assert np.array_equal(labels.cpu().numpy(), w_labels)
    assert np.array_equal(num.cpu().numpy(), w_num)


def test_dataset_matches_reference_layout_and_replay():
    g = _small_graph()
    ds = PMGTDataset(g, seed=5)
    idx = [0, 3, 17, 250]
    t, p, num_pairs, labels = ds.sample_batch(idx, epoch=2)
    B, L, P = len(idx), 6, 10
    assert t["node_ids"].shape == (B, L) and t["node_ids"].dtype == torch.int64
    assert t["attention_mask"].shape == (B, L) and t["attention_mask"].dtype == torch.float32
    assert p["node_ids"].shape == (B * P, L) and labels.shape == (B * P,) and labels.dtype == torch.float32
    assert num_pairs.dtype == torch.int64 and num_pairs.tolist() == [P] * B
    # replay the whole batch on the CPU
    targets = ds.node_ids[idx]
    tkeys = (2 << 40) | (targets << 8)
    w_pairs, w_labels, _ = sampler_ref.philox_sample_pairs(g.indptr, g.indices, g.num_nodes, targets, tkeys, 5, 5, 10, P, 5)
    pkeys = ((2 << 40) | (targets[:, None] << 8) | np.arange(1, P + 1)[None, :]).reshape(-1)
    roots = np.concatenate([targets, w_pairs.reshape(-1)])
    keys = np.concatenate([tkeys, pkeys])
    w_ids, w_mask, _ = sampler_ref.philox_sample_contexts(g.indptr, g.indices, g.cdf, g.num_nodes, roots, keys,
                                                          [16, 8, 4], 5, 5)
    assert np.array_equal(torch.cat([t["node_ids"], p["node_ids"]]).cpu().numpy(), w_ids)
    assert np.array_equal(torch.cat([t["attention_mask"], p["attention_mask"]]).cpu().numpy(), w_mask)
    assert np.array_equal(labels.cpu().numpy(), w_labels.reshape(-1))
    # item-at-a-time contract + collate (reference layout, CPU tensors)
    items = [ds[i] for i in idx]
    ct, cp, cn, cl = pmgt_collate_fn(items)
    assert ct["node_ids"].shape == (B, L) and not ct["node_ids"].is_cuda
    assert cp["node_ids"].shape == (B * P, L) and cn.tolist() == [P] * B and cl.shape == (B * P,)
    assert ct["node_ids"][:, 0].tolist() == targets.tolist()


def test_dataset_validation_and_inference_modes():
    g = _small_graph()
    v = PMGTDataset(g, is_training=False)
    t, p, n, lab = v.sample_batch([1, 2, 3])
    assert p["node_ids"].shape == (6, 6) and n.tolist() == [2, 2, 2] and lab.tolist() == [1.0, 0.0] * 3
    inf = PMGTDataset(g, is_training=False, is_inference=True)
    out = inf.sample_batch(np.arange(10))
    assert set(out.keys()) == {"node_ids", "attention_mask"} and out["node_ids"].shape == (10, 6)
    assert len(inf[0]) == 1 and pmgt_collate_fn([inf[0], inf[1]])["node_ids"].shape == (2, 6)


def test_results_do_not_depend_on_batching():
    """Draws are keyed on (seed, epoch, node, slot): sharding a batch changes nothing."""
    g = _small_graph()
    ds = PMGTDataset(g, seed=11)
    full = ds.sample_batch(np.arange(64), epoch=1)
    a = ds.sample_batch(np.arange(0, 32), epoch=1)
    b = ds.sample_batch(np.arange(32, 64), epoch=1)
    assert torch.equal(full[0]["node_ids"], torch.cat([a[0]["node_ids"], b[0]["node_ids"]]))
    assert torch.equal(full[1]["node_ids"], torch.cat([a[1]["node_ids"], b[1]["node_ids"]]))
    other = ds.sample_batch(np.arange(64), epoch=2)
    assert not torch.equal(full[1]["node_ids"], other[1]["node_ids"])


def test_full_size_properties_tg_graph():
    """BASELINE config 2 graph at full size: size-independent invariants + replay on a subset."""
    g = synthetic.make_item_graph("TG")
    assert g.num_nodes == 10834 and g.num_edges_directed == 2 * 38252
    ds = PMGTDataset(g, seed=3)
    t, p, num_pairs, labels = ds.sample_batch(np.arange(4096), epoch=0)
    ids = torch.cat([t["node_ids"], p["node_ids"]])
    mask = torch.cat([t["attention_mask"], p["attention_mask"]])
    assert ids.shape == (4096 * 11, 6)
    assert torch.equal(ids == 0, mask == 0)
    assert bool(((ids[:, 1:] != ids[:, :1]) | (ids[:, 1:] == 0)).all())      # root never in its own context
    s = torch.sort(ids[:, 1:], dim=1).values
    assert bool(((s[:, 1:] != s[:, :-1]) | (s[:, 1:] == 0)).all())           # no duplicate neighbours
    m = mask[:, 1:]
    assert bool((m[:, 1:] <= m[:, :-1]).all())                                # pads only on the right
    assert int(ids.max()) <= g.num_nodes + 1
    # positives are neighbours of their target, negatives are not
    indptr, indices = g.indptr, g.indices
    tg = t["node_ids"][:, 0].cpu().numpy()
    pr = p["node_ids"][:, 0].cpu().numpy().reshape(4096, 10)
    lb = labels.cpu().numpy().reshape(4096, 10)
    for i in range(0, 4096, 97):
        nb = set(indices[indptr[tg[i]]: indptr[tg[i] + 1]].tolist())
        assert all((int(x) in nb) == bool(l) for x, l in zip(pr[i], lb[i]))
    sub = np.arange(0, 4096 * 11, 61)
    roots = ids[:, 0].cpu().numpy()[sub]
    # keys as the dataset forms them
    targets = ds.node_ids[np.arange(4096)]
    tkeys = targets << 8
    pkeys = ((targets[:, None] << 8) | np.arange(1, 11)[None, :]).reshape(-1)
    keys = np.concatenate([tkeys, pkeys])[sub]
    w_ids, w_mask, _ = sampler_ref.philox_sample_contexts(indptr, indices, g.cdf, g.num_nodes, roots, keys, [16, 8, 4], 5, 3)
    assert np.array_equal(ids.cpu().numpy()[sub], w_ids) and np.array_equal(mask.cpu().numpy()[sub], w_mask)
