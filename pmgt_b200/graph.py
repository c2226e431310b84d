"""Item graph in CSR form -- the device-resident replacement for the weighted
``nx.Graph`` the reference samples from (``pmgt/pmgt/trainer.py:34-41``,
``pmgt/pmgt/datasets.py:27-32``).

Node-id space is the reference's: 0 = ``<pad>``, 1 = ``<mask>``, real nodes
``2..N+1`` (datasets.py:96-102).  ``indptr`` is indexed by node id and has
``N + 3`` entries.  Row order is the graph's adjacency *insertion* order, so an
inverse-CDF position means the same neighbour as in the reference.
"""
from typing import Optional

import numpy as np


def _row_softmax_cdf(indptr: np.ndarray, weights: np.ndarray) -> np.ndarray:
    """Running softmax CDF per CSR row (vectorised), fp64 math, fp32 result.

    Reference: ``ss.softmax(weights)`` (datasets.py:27-29) followed by numpy's
    legacy ``choice``: ``cdf = p.cumsum(); cdf /= cdf[-1]``.
    """
    n_rows = len(indptr) - 1
    deg = np.diff(indptr)
    if len(weights) == 0:
        return np.zeros(0, dtype=np.float32)
    w = np.asarray(weights, dtype=np.float64)
    row_of = np.repeat(np.arange(n_rows), deg)
    row_max = np.full(n_rows, -np.inf)
    np.maximum.at(row_max, row_of, w)
    e = np.exp(w - row_max[row_of])
    c = np.cumsum(e)
    starts = indptr[:-1]
    nz = deg > 0
    # subtract the running total before each row
    before = np.zeros(n_rows)
    before[nz] = np.where(starts[nz] > 0, c[np.maximum(starts[nz] - 1, 0)], 0.0)
    c = c - before[row_of]
    row_sum = np.zeros(n_rows)
    row_sum[nz] = c[indptr[1:][nz] - 1]
    cdf = (c / row_sum[row_of]).astype(np.float32)
    cdf[indptr[1:][nz] - 1] = 1.0
    return cdf


def _device_array_as_tensor(ptr: int, n: int, dtype, device):
    """A torch view of ``n`` elements of library-owned device memory (``__cuda_array_interface__``)."""
    import torch

    itemsize = torch.empty((), dtype=dtype).element_size()
    typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.int64: "<i8"}[dtype]

    class _Wrap:
        __cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2,
                                    "strides": (itemsize,)}

    return torch.as_tensor(_Wrap(), device=device)


class ItemGraph:
    """CSR item graph + per-row softmax CDF; optionally resident on a GPU."""

    def __init__(self, num_nodes: int, indptr: np.ndarray, indices: np.ndarray,
                 weights: np.ndarray, cdf: Optional[np.ndarray] = None):
        indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        weights_in = np.asarray(weights)  # the softmax CDF is formed from the weights at the precision they arrive in
        weights = np.ascontiguousarray(weights, dtype=np.float32)
        if indptr.shape != (num_nodes + 3,):
            raise ValueError(f"indptr must have num_nodes+3={num_nodes + 3} entries, got {indptr.shape}")
        if indptr[0] != 0 or indptr[-1] != len(indices) or np.any(np.diff(indptr) < 0):
            raise ValueError("indptr is not a valid CSR row-pointer array")
        if indptr[2] != 0:
            raise ValueError("rows 0 (<pad>) and 1 (<mask>) must be empty")
        if len(indices) and (indices.min() < 2 or indices.max() > num_nodes + 1):
            raise ValueError("neighbour ids must lie in [2, num_nodes+1]")
        if len(weights) != len(indices):
            raise ValueError("weights and indices differ in length")
        self.num_nodes = int(num_nodes)
        self.indptr = indptr
        self.indices = indices
        self.weights = weights
        self.cdf = (np.ascontiguousarray(cdf, dtype=np.float32) if cdf is not None
                    else _row_softmax_cdf(indptr, weights_in))  # fp64 math on fp64 weights like ss.softmax (datasets.py:27-29)
        self._handles = {}  # device index -> opaque pmgt_graph*

    # -- nx.Graph-like surface used by the reference's callers -----------------
    def __len__(self) -> int:  # len(graph) == number of nodes (trainer.py:127, datasets.py:101)
        return self.num_nodes

    @property
    def num_edges_directed(self) -> int:
        return int(len(self.indices))

    def degree(self, node: int) -> int:
        return int(self.indptr[node + 1] - self.indptr[node])

    def neighbors(self, node: int) -> np.ndarray:
        return self.indices[self.indptr[node]: self.indptr[node + 1]]

    # -- constructors -----------------------------------------------------------
    @classmethod
    def from_networkx(cls, graph) -> "ItemGraph":
        """Convert a weighted ``nx.Graph`` whose nodes are already relabelled to
        ints ``2..N+1`` (trainer.py:38-41), keeping adjacency insertion order."""
        n = len(graph)
        indptr = np.zeros(n + 3, dtype=np.int64)
        idx, wts = [], []
        for node in range(2, n + 2):
            if node not in graph:
                raise ValueError(f"graph nodes must be labelled 2..{n + 1}; {node} is missing")
            adj = graph[node]
            idx.extend(adj.keys())
            wts.extend(d["weight"] for d in adj.values())
            indptr[node + 1] = len(idx)
        return cls(n, indptr, np.asarray(idx, dtype=np.int32), np.asarray(wts, dtype=np.float64))

    @classmethod
    def from_edge_list(cls, num_nodes: int, src: np.ndarray, dst: np.ndarray, weight: np.ndarray) -> "ItemGraph":
        """Undirected edge list (node ids in ``2..N+1``) -> CSR, reproducing the
        adjacency insertion order ``nx.Graph.add_weighted_edges_from`` would give
        (edge i appends dst to src's row and src to dst's row, in edge order)."""
        src = np.asarray(src, dtype=np.int64)
        dst = np.asarray(dst, dtype=np.int64)
        weight = np.asarray(weight)
        # nx.Graph keeps ONE edge per unordered pair: a repeated (u, v) updates the weight of the existing edge and
        # leaves its position in both adjacency lists where the first occurrence put it
        code = np.minimum(src, dst) << 32 | np.maximum(src, dst)
        uniq, first, inverse = np.unique(code, return_index=True, return_inverse=True)
        if len(uniq) != len(code):
            last = np.zeros(len(uniq), dtype=np.int64)
            np.maximum.at(last, inverse, np.arange(len(code)))       # last occurrence carries the surviving weight
            keep = np.sort(first)
            w_last = weight[last]
            weight = w_last[inverse[keep]]
            src, dst = src[keep], dst[keep]
        m = len(src)
        order = np.arange(m, dtype=np.int64)
        rows = np.concatenate([src, dst])
        cols = np.concatenate([dst, src])
        wts = np.concatenate([weight, weight])
        seq = np.concatenate([2 * order, 2 * order + 1])
        loops = src == dst
        if loops.any():  # a self loop is a single adjacency entry
            keep = np.concatenate([np.ones(m, bool), ~loops])
            rows, cols, wts, seq = rows[keep], cols[keep], wts[keep], seq[keep]
        perm = np.lexsort((seq, rows))
        rows, cols, wts = rows[perm], cols[perm], wts[perm]
        counts = np.bincount(rows, minlength=num_nodes + 2)
        indptr = np.zeros(num_nodes + 3, dtype=np.int64)
        np.cumsum(counts, out=indptr[1:])
        return cls(num_nodes, indptr, cols.astype(np.int32), wts)

    @classmethod
    def from_edge_list_device(cls, num_nodes: int, src, dst, weight, device=None) -> "ItemGraph":
        """``from_edge_list`` on the GPU: the doubled edge list is sorted by (row, insertion sequence) with one stable
        device sort, and the softmax CDF plus the sampler's lookup tables are built by kernels
        (``pmgt_graph_create_device``).  Same CSR, same neighbour order; the CDF agrees with the host builder to fp32
        rounding (the host copy kept in ``self.cdf`` is the device result, so CPU replays see the same values).
        Edge lists with self loops or repeated pairs take the host path (networkx de-duplication semantics)."""
        import torch

        from . import _lib

        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        src = torch.as_tensor(src, dtype=torch.int64, device=dev)
        dst = torch.as_tensor(dst, dtype=torch.int64, device=dev)
        w = torch.as_tensor(weight, dtype=torch.float64, device=dev)
        m = src.numel()
        code = torch.minimum(src, dst) << 32 | torch.maximum(src, dst)
        if bool((src == dst).any()) or torch.unique(code).numel() != m:
            return cls.from_edge_list(num_nodes, src.cpu().numpy(), dst.cpu().numpy(), w.cpu().numpy())
        del code
        rows = torch.cat([src, dst])
        # key = (row, 2 * edge + side): edge i appends dst to src's row and then src to dst's row
        seq = torch.cat([2 * torch.arange(m, device=dev), 2 * torch.arange(m, device=dev) + 1])
        key = rows * (2 * m) + seq
        del seq
        perm = torch.argsort(key)
        del key
        cols = torch.cat([dst, src])[perm].to(torch.int32)
        wts = torch.cat([w, w])[perm]
        counts = torch.bincount(rows, minlength=num_nodes + 2)
        del rows, perm
        indptr = torch.zeros(num_nodes + 3, dtype=torch.int64, device=dev)
        torch.cumsum(counts, 0, out=indptr[1:])
        h = _lib.graph_create_device(dev.index, num_nodes, indptr, cols, wts)
        cdf = torch.empty(cols.numel(), dtype=torch.float32, device=dev)
        if cols.numel():
            cdf = _device_array_as_tensor(_lib.lib().pmgt_graph_cdf(h), cols.numel(), torch.float32, dev).clone()
        g = cls(num_nodes, indptr.cpu().numpy(), cols.cpu().numpy(), wts.cpu().numpy(), cdf=cdf.cpu().numpy())
        g._handles[dev.index] = h
        return g

    # -- device residency ---------------------------------------------------------
    def device_handle(self, device_index: int):
        """Opaque ``pmgt_graph*`` for ``cuda:device_index`` (created on first use)."""
        from . import _lib

        h = self._handles.get(device_index)
        if h is None:
            h = _lib.graph_create(device_index, self.num_nodes, self.indptr, self.indices, self.cdf)
            self._handles[device_index] = h
        return h

    def isolated_nodes(self) -> np.ndarray:
        deg = np.diff(self.indptr)[2:]
        return np.nonzero(deg == 0)[0] + 2

    def __del__(self):
        try:
            from . import _lib

            for h in self._handles.values():
                _lib.graph_destroy(h)
        except Exception:
            pass
        self._handles = {}
