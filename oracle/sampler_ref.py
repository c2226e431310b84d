"""TEST INFRASTRUCTURE ONLY -- CPU oracles for the MCNSampling path.

Two oracles live here:

1. ``ref_*``: a numpy restatement of the reference sampler
   (``pmgt/pmgt/datasets.py:14-53,113-183``) on CSR arrays, consuming the legacy
   ``np.random`` stream exactly as the reference does (``ss.softmax`` ->
   ``RandomState.choice(p=)`` = ``cdf.searchsorted(random_sample(n), "right")``;
   ``choice(replace=False)`` = ``permutation(n)[:k]``; ``randint``).  Pinned in
   ``tests/test_oracle_sampler.py`` against the unmodified reference under the
   same seed (bit-exact) and against ``tests/golden/sampler_ref_golden.npz``.
   It is the sample generator for the statistical-equivalence tests and the
   sampler half of the CPU baseline.

2. ``philox_*``: ctypes wrappers over ``oracle/philox_sampler.c``, the CPU
   replay of the Philox stream the CUDA kernels must reproduce bit for bit.
"""
import ctypes
from collections import Counter, defaultdict

import numpy as np
import scipy.special as ss

from . import build as _build


# --------------------------------------------------------------------------
# 1. reference algorithm on CSR, np.random stream
# --------------------------------------------------------------------------
def ref_sample_context(indptr, indices, weights, target, hops, max_ctx, rng=np.random):
    """datasets.py:14-53 on CSR rows (row order == nx adjacency insertion order)."""
    depth = len(hops)
    scores = defaultdict(int)
    frontier = [int(target)]
    for k, size in enumerate(hops, start=1):
        drawn = []
        for node in frontier:
            lo, hi = int(indptr[node]), int(indptr[node + 1])
            p = ss.softmax(np.asarray(weights[lo:hi], dtype=np.float64))
            # RandomState.choice(a, size, replace=True, p): inverse CDF
            cdf = p.cumsum()
            cdf /= cdf[-1]
            u = rng.random_sample(size)
            pos = cdf.searchsorted(u, side="right")
            drawn.extend(np.asarray(indices[lo:hi])[pos].tolist())
        for node, freq in Counter(drawn).items():
            if node != target:
                scores[node] += freq * (depth - k + 1)
        frontier = drawn
    ranked = [n for n, _ in sorted(scores.items(), key=lambda kv: kv[1], reverse=True)]
    n_real = min(len(ranked), max_ctx)
    ranked = ranked[:max_ctx] + [0] * max(0, max_ctx - len(ranked))
    return ranked, n_real


def ref_input_tensor(indptr, indices, weights, target, hops, max_ctx, rng=np.random):
    """datasets.py:56-79 -> (ids int64 (L,), mask float32 (L,))."""
    ctx, n_real = ref_sample_context(indptr, indices, weights, target, hops, max_ctx, rng)
    mask = np.zeros(max_ctx + 1, dtype=np.float32)
    mask[: n_real + 1] = 1
    return np.asarray([int(target)] + ctx, dtype=np.int64), mask


def ref_getitem(indptr, indices, weights, num_nodes, target, hops=(16, 8, 4), max_ctx=5,
                max_total=10, min_neg=5, is_training=True, is_inference=False, rng=np.random):
    """PMGTDataset.__getitem__ (datasets.py:113-183) on CSR."""
    tgt = ref_input_tensor(indptr, indices, weights, target, hops, max_ctx, rng)
    if is_inference:
        return (tgt,)
    lo, hi = int(indptr[target]), int(indptr[target + 1])
    neigh = np.asarray(indices[lo:hi])
    n_pos = min((max_total - min_neg) if is_training else 1, len(neigh))
    # RandomState.choice(a, k, replace=False) == a[permutation(len(a))[:k]]
    pos_nodes = neigh[rng.permutation(len(neigh))[:n_pos]].tolist()
    pos = [ref_input_tensor(indptr, indices, weights, n, hops, max_ctx, rng) for n in pos_nodes]
    n_neg = max(min_neg, max_total - len(pos_nodes)) if is_training else 1
    neigh_set = set(neigh.tolist())
    neg_nodes = []
    for _ in range(n_neg):
        cand = rng.randint(num_nodes) + 2
        while cand in neigh_set:
            cand = rng.randint(num_nodes) + 2
        neg_nodes.append(cand)
    neg = [ref_input_tensor(indptr, indices, weights, n, hops, max_ctx, rng) for n in neg_nodes]
    ids = np.stack([x[0] for x in pos + neg])
    mask = np.stack([x[1] for x in pos + neg])
    labels = np.asarray([1.0] * len(pos) + [0.0] * len(neg), dtype=np.float32)
    return tgt, (ids, mask), labels


def ref_collate(batch):
    """pmgt_collate_fn (datasets.py:186-208) on numpy items."""
    target = {
        "node_ids": np.stack([b[0][0] for b in batch]),
        "attention_mask": np.stack([b[0][1] for b in batch]),
    }
    if len(batch[0]) == 1:
        return target
    pair = {
        "node_ids": np.concatenate([b[1][0] for b in batch]),
        "attention_mask": np.concatenate([b[1][1] for b in batch]),
    }
    num_pairs = np.asarray([len(b[1][0]) for b in batch], dtype=np.int64)
    labels = np.concatenate([b[2] for b in batch])
    return target, pair, num_pairs, labels


def softmax_cdf_f32(indptr, weights):
    """Per-row softmax CDF exactly as the reference forms it (fp64), rounded to fp32,
    last entry of each row forced to 1.0.  Independent restatement of
    ``pmgt_b200.graph.ItemGraph``'s CDF builder, used to cross-check it."""
    cdf = np.zeros(len(weights), dtype=np.float32)
    for r in range(len(indptr) - 1):
        lo, hi = int(indptr[r]), int(indptr[r + 1])
        if hi > lo:
            c = ss.softmax(np.asarray(weights[lo:hi], dtype=np.float64)).cumsum()
            c /= c[-1]
            c32 = c.astype(np.float32)
            c32[-1] = 1.0
            cdf[lo:hi] = c32
    return cdf


# --------------------------------------------------------------------------
# 2. Philox replay (plain C)
# --------------------------------------------------------------------------
def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def philox4x32_10(counter, key):
    lib = _build.load()
    out = (ctypes.c_uint32 * 4)()
    lib.pmgt_oracle_philox(*(ctypes.c_uint32(int(c)) for c in counter),
                           ctypes.c_uint32(int(key[0])), ctypes.c_uint32(int(key[1])), out)
    return [int(x) for x in out]


def philox_sample_contexts(indptr, indices, cdf, num_nodes, roots, keys, hops, max_ctx, seed):
    lib = _build.load()
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    cdf = np.ascontiguousarray(cdf, dtype=np.float32)
    roots = np.ascontiguousarray(roots, dtype=np.int64)
    keys = np.ascontiguousarray(keys, dtype=np.int64)
    hops = np.ascontiguousarray(hops, dtype=np.int32)
    n = len(roots)
    L = max_ctx + 1
    ids = np.zeros((n, L), dtype=np.int64)
    mask = np.zeros((n, L), dtype=np.float32)
    vdeg = np.zeros(n, dtype=np.int64)
    rc = lib.pmgt_oracle_sample_contexts(
        _p(indptr, ctypes.c_int64), _p(indices, ctypes.c_int32), _p(cdf, ctypes.c_float),
        ctypes.c_int64(num_nodes), _p(roots, ctypes.c_int64), _p(keys, ctypes.c_int64),
        ctypes.c_int64(n), _p(hops, ctypes.c_int32), ctypes.c_int(len(hops)), ctypes.c_int(max_ctx),
        ctypes.c_uint64(seed), _p(ids, ctypes.c_int64), _p(mask, ctypes.c_float),
        _p(vdeg, ctypes.c_int64))
    assert rc == 0
    return ids, mask, vdeg


def philox_sample_pairs(indptr, indices, num_nodes, targets, keys, max_pos, min_neg, max_total,
                        stride, seed):
    lib = _build.load()
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    targets = np.ascontiguousarray(targets, dtype=np.int64)
    keys = np.ascontiguousarray(keys, dtype=np.int64)
    n = len(targets)
    pairs = np.zeros((n, stride), dtype=np.int64)
    labels = np.zeros((n, stride), dtype=np.float32)
    num = np.zeros(n, dtype=np.int64)
    rc = lib.pmgt_oracle_sample_pairs(
        _p(indptr, ctypes.c_int64), _p(indices, ctypes.c_int32), ctypes.c_int64(num_nodes),
        _p(targets, ctypes.c_int64), _p(keys, ctypes.c_int64), ctypes.c_int64(n),
        ctypes.c_int(max_pos), ctypes.c_int(min_neg), ctypes.c_int(max_total), ctypes.c_int(stride),
        ctypes.c_uint64(seed), _p(pairs, ctypes.c_int64), _p(labels, ctypes.c_float),
        _p(num, ctypes.c_int64))
    assert rc == 0
    return pairs, labels, num
