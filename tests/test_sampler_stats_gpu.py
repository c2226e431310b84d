"""Statistical equivalence of the CUDA MCNSampling kernel and the reference sampler (north_star: "its neighbour
distribution is statistically equivalent to the reference MCNSampling").  The reference side is
oracle.sampler_ref.ref_* (bit-exactly pinned to the unmodified reference in tests/test_oracle_sampler.py)."""
import numpy as np
import pytest
import torch

from oracle import sampler_ref
from pmgt_b200 import synthetic
from pmgt_b200.datasets import context_keys, sample_contexts

pytestmark = pytest.mark.gpu


def _chi2_p(obs_a, obs_b):
    from scipy.stats import chi2_contingency
    keep = (obs_a + obs_b) >= 10
    table = np.stack([obs_a[keep], obs_b[keep]])
    table = np.concatenate([table, np.stack([[obs_a[~keep].sum()], [obs_b[~keep].sum()]])], axis=1) if (~keep).any() else table
    table = table[:, table.sum(0) > 0]
    return chi2_contingency(table)[1]


def test_hop1_draw_frequencies_follow_softmax_of_weights():
    """One hop, size 1, max_ctx 1: the context neighbour is a single draw from softmax(edge weights)."""
    g = synthetic.make_item_graph((200, 1500), seed=2)
    node = int(np.argmax(np.diff(g.indptr)))  # highest-degree node
    n = 40000
    roots = torch.full((n,), node, dtype=torch.int64, device="cuda")
    keys = torch.arange(n, dtype=torch.int64, device="cuda")
    ids, _ = sample_contexts(g, roots, keys, [1], 1, seed=5)
    nb = g.neighbors(node)
    counts = np.asarray([(ids[:, 1].cpu().numpy() == v).sum() for v in nb], dtype=np.float64)
    lo, hi = g.indptr[node], g.indptr[node + 1]
    w = g.weights[lo:hi].astype(np.float64)
    p = np.exp(w - w.max())
    p /= p.sum()
    from scipy.stats import chisquare
    keep = p * n >= 5
    exp = np.append(p[keep] * n, max(p[~keep].sum() * n, 1e-9))
    obs = np.append(counts[keep], counts[~keep].sum())
    if exp[-1] < 1e-6:
        exp, obs = exp[:-1], obs[:-1]
    exp *= obs.sum() / exp.sum()
    assert chisquare(obs, exp)[1] > 1e-3


@pytest.mark.parametrize("target_rank", [0, 5, 60])
def test_context_membership_matches_reference_sampler(target_rank):
    """Full [16, 8, 4] sampler: how often each node lands in the context, and at which position, vs the reference."""
    g = synthetic.make_item_graph((200, 1500), seed=2)
    order = np.argsort(-np.diff(g.indptr))
    node = int(order[target_rank])
    n_gpu, n_ref = 6000, 1500
    roots = torch.full((n_gpu,), node, dtype=torch.int64, device="cuda")
    keys = context_keys(3, roots, torch.arange(n_gpu, device="cuda") % 200) + (torch.arange(n_gpu, device="cuda") << 45)
    ids, mask = sample_contexts(g, roots, keys, [16, 8, 4], 5, seed=9)
    got = ids.cpu().numpy()
    np.random.seed(123)
    w64 = g.weights.astype(np.float64)
    ref = np.stack([sampler_ref.ref_input_tensor(g.indptr, g.indices, w64, node, [16, 8, 4], 5)[0] for _ in range(n_ref)])
    nn = g.num_nodes + 2
    # membership frequency of every node in the context (any position)
    a = np.bincount(got[:, 1:].reshape(-1), minlength=nn)[2:].astype(np.float64)
    b = np.bincount(ref[:, 1:].reshape(-1), minlength=nn)[2:].astype(np.float64)
    assert _chi2_p(a, b) > 1e-3
    # top-1 neighbour distribution (exercises scoring + tie-breaking)
    a1 = np.bincount(got[:, 1], minlength=nn)[2:].astype(np.float64)
    b1 = np.bincount(ref[:, 1], minlength=nn)[2:].astype(np.float64)
    assert _chi2_p(a1, b1) > 1e-3
    # number of real (non-pad) neighbours
    assert abs((got[:, 1:] != 0).sum(1).mean() - (ref[:, 1:] != 0).sum(1).mean()) < 0.05


def test_pair_selection_distribution():
    """Positives uniform over neighbours without replacement; negatives uniform over non-neighbours."""
    from pmgt_b200 import ops
    g = synthetic.make_item_graph((120, 700), seed=6)
    node = int(np.argsort(-np.diff(g.indptr))[3])
    n = 20000
    dev = torch.device("cuda", 0)
    t = torch.full((n,), node, dtype=torch.int64, device=dev)
    keys = torch.arange(n, dtype=torch.int64, device=dev) << 8
    pairs = torch.empty((n, 10), dtype=torch.int64, device=dev)
    labels = torch.empty((n, 10), dtype=torch.float32, device=dev)
    num = torch.empty(n, dtype=torch.int64, device=dev)
    ops.sample_pairs(g.device_handle(0), t, keys, 5, 5, 10, 10, 4, pairs, labels, num)
    pr = pairs.cpu().numpy()
    nb = g.neighbors(node)
    from scipy.stats import chisquare
    pos = np.asarray([(pr[:, :5] == v).sum() for v in nb], dtype=np.float64)
    assert len(nb) >= 5 and chisquare(pos)[1] > 1e-3
    nonnb = np.setdiff1d(np.arange(2, g.num_nodes + 2), nb)
    neg = np.asarray([(pr[:, 5:] == v).sum() for v in nonnb], dtype=np.float64)
    assert neg.sum() == 5 * n and chisquare(neg)[1] > 1e-3
