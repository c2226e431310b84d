"""Host-side mirror of ``pmgt/pmgt/datasets.py``: ``PMGTDataset``,
``pmgt_collate_fn``, ``get_input_tensor`` -- MCNSampling on the GPU.

The reference samples each context in Python inside DataLoader worker
processes (145 softmaxes + 656 weighted draws per context).  Here the item
graph lives on the device as CSR + per-row softmax CDF and a whole batch of
contexts is one kernel launch driven by counter-based Philox4x32-10; draws are
keyed on (seed, epoch, target node, context slot, draw index), so results do
not depend on batch composition, worker count or world size.

``__getitem__`` / ``pmgt_collate_fn`` keep the reference's item-at-a-time
contract (CPU tensors, same shapes and dtypes); the fast path used by the
trainer is ``sample_batch`` which returns the collated batch on the device.
"""
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import ops
from ._lib import PMGTError
from .graph import ItemGraph

_SLOT_BITS = 8
_NODE_BITS = 32


def _as_item_graph(graph) -> ItemGraph:
    if isinstance(graph, ItemGraph):
        return graph
    cached = getattr(graph, "_pmgt_item_graph", None)
    if cached is None:
        cached = ItemGraph.from_networkx(graph)
        try:
            graph._pmgt_item_graph = cached
        except Exception:
            pass
    return cached


def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise PMGTError("MCNSampling runs on the GPU and no CUDA device is available (pmgt_b200 has no CPU fallback)")
    return torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())


def context_keys(epoch: int, nodes: torch.Tensor, slot) -> torch.Tensor:
    """64-bit Philox sub-stream id of a context: (epoch, target node, slot)."""
    return (int(epoch) << (_NODE_BITS + _SLOT_BITS)) | (nodes << _SLOT_BITS) | slot


def sample_contexts(graph: ItemGraph, roots: torch.Tensor, keys: torch.Tensor, hop_sampling_sizes: Sequence[int],
                    max_num_ctx_neigh: int, seed: int, want_visited_deg: bool = False):
    """Batch form of ``get_input_tensor``: (ids int64 [n, L], mask float32 [n, L])."""
    dev = roots.device
    n = roots.numel()
    L = max_num_ctx_neigh + 1
    ids = torch.empty((n, L), dtype=torch.int64, device=dev)
    mask = torch.empty((n, L), dtype=torch.float32, device=dev)
    vdeg = torch.empty(n, dtype=torch.int64, device=dev) if want_visited_deg else None
    ops.sample_contexts(graph.device_handle(dev.index), roots.contiguous(), keys.contiguous(),
                        [int(h) for h in hop_sampling_sizes], max_num_ctx_neigh, seed, ids, mask, vdeg)
    return (ids, mask, vdeg) if want_visited_deg else (ids, mask)


def get_input_tensor(graph, target_node: int, hop_sampling_sizes: List[int], max_num_ctx_neigh: int,
                     seed: int = 0, epoch: int = 0) -> Tuple[torch.LongTensor, torch.FloatTensor]:
    """datasets.py:64-79 for one node; returns CPU tensors like the reference."""
    g = _as_item_graph(graph)
    dev = _device()
    roots = torch.tensor([int(target_node)], dtype=torch.int64, device=dev)
    ids, mask = sample_contexts(g, roots, context_keys(epoch, roots, 0), hop_sampling_sizes, max_num_ctx_neigh, seed)
    assert ids.shape[1] == max_num_ctx_neigh + 1, f"# of context nodes must be {max_num_ctx_neigh}"
    return ids[0].cpu(), mask[0].cpu()


class PMGTDataset(torch.utils.data.Dataset):
    def __init__(
        self,
        graph,
        node_ids: Optional[np.ndarray] = None,
        max_ctx_neigh: int = 5,
        hop_sampling_sizes: List[int] = [16, 8, 4],
        max_total_samples: int = 10,
        min_neg_samples: int = 5,
        is_training: bool = True,
        is_inference: bool = False,
        seed: int = 0,
        device=None,
    ) -> None:
        super().__init__()
        self.graph = graph
        self.item_graph = _as_item_graph(graph)
        # 0 is <pad>, 1 is <mask>
        self.node_ids = node_ids if node_ids is not None else np.arange(start=2, stop=len(self.item_graph) + 2)
        self.node_ids = np.asarray(self.node_ids, dtype=np.int64)
        self.max_num_ctx_neigh = max_ctx_neigh
        self.hop_sampling_sizes = list(hop_sampling_sizes)
        self.max_total_samples = max_total_samples
        self.min_neg_samples = min_neg_samples
        self.is_training = is_training
        self.is_inference = is_inference
        self.seed = int(seed)
        self._device_arg = device
        self._node_ids_dev = None
        self._calls = 0
        if not is_inference:
            deg = np.diff(self.item_graph.indptr)[self.node_ids]
            if (deg == 0).any():
                # the reference fails inside np.random.choice on an empty neighbour list
                raise ValueError(f"{int((deg == 0).sum())} target nodes have no neighbours; "
                                 "'a' cannot be empty unless no samples are taken")

    def __len__(self) -> int:
        return len(self.node_ids)  # num of nodes

    # -- pair layout --------------------------------------------------------------
    @property
    def max_pos(self) -> int:
        return (self.max_total_samples - self.min_neg_samples) if self.is_training else 1

    @property
    def pairs_per_target(self) -> int:
        """n_pos + n_neg is constant for targets with >= 1 neighbour."""
        if self.is_inference:
            return 0
        return max(self.max_total_samples, self.min_neg_samples) if self.is_training else 2

    def _dev(self) -> torch.device:
        return _device(self._device_arg)

    def sample_batch(self, indices, epoch: int = 0):
        """Collated batch for dataset positions ``indices`` on the device.

        Returns what ``pmgt_collate_fn`` returns: ``target_inputs`` only for an
        inference dataset, else ``(target_inputs, pair_inputs, num_pairs, labels)``.
        """
        dev = self._dev()
        if self._node_ids_dev is None or self._node_ids_dev.device != dev:
            self._node_ids_dev = torch.from_numpy(self.node_ids).to(dev)
        idx = torch.as_tensor(indices, dtype=torch.int64, device=dev)
        targets = self._node_ids_dev[idx]
        B = targets.numel()
        g = self.item_graph
        if self.is_inference:
            ids, mask = sample_contexts(g, targets, context_keys(epoch, targets, 0), self.hop_sampling_sizes,
                                        self.max_num_ctx_neigh, self.seed)
            return {"node_ids": ids, "attention_mask": mask}
        if self.is_training:
            max_pos, min_neg, max_total = self.max_pos, self.min_neg_samples, self.max_total_samples
        else:
            max_pos, min_neg, max_total = 1, 1, 2
        P = max(max_pos + min_neg, max_total)
        tkeys = context_keys(epoch, targets, 0)
        pairs = torch.empty((B, P), dtype=torch.int64, device=dev)
        labels = torch.empty((B, P), dtype=torch.float32, device=dev)
        num_pairs = torch.empty(B, dtype=torch.int64, device=dev)
        ops.sample_pairs(g.device_handle(dev.index), targets.contiguous(), tkeys, max_pos, min_neg, max_total, P,
                         self.seed, pairs, labels, num_pairs)
        # roots of every context of the batch: targets first, then the pair nodes row by row
        slots = torch.arange(1, P + 1, device=dev, dtype=torch.int64)
        pkeys = context_keys(epoch, targets[:, None].expand(B, P), slots[None, :]).reshape(-1)
        roots = torch.cat([targets, pairs.reshape(-1)])
        keys = torch.cat([tkeys, pkeys])
        ids, mask = sample_contexts(g, roots, keys, self.hop_sampling_sizes, self.max_num_ctx_neigh, self.seed)
        target_inputs = {"node_ids": ids[:B], "attention_mask": mask[:B]}
        pair_inputs = {"node_ids": ids[B:], "attention_mask": mask[B:]}
        return target_inputs, pair_inputs, num_pairs, labels.reshape(-1)

    def __getitem__(self, idx: int):
        """One item in the reference's layout (CPU tensors): ``(target_inputs,)`` or
        ``(target_inputs, pair_inputs, labels)`` with ``*_inputs = (ids, mask)``."""
        self._calls += 1
        out = self.sample_batch([int(idx)], epoch=self._calls)
        if self.is_inference:
            return ((out["node_ids"][0].cpu(), out["attention_mask"][0].cpu()),)
        t, p, n, lab = out
        n = int(n[0])
        return ((t["node_ids"][0].cpu(), t["attention_mask"][0].cpu()),
                (p["node_ids"][:n].cpu(), p["attention_mask"][:n].cpu()), lab[:n].cpu())


def pmgt_collate_fn(batch: Iterable[Tuple[torch.Tensor, ...]]) -> Union[Dict[str, torch.Tensor], tuple]:
    """datasets.py:186-208."""
    batch = list(batch)
    target_inputs = {
        "node_ids": torch.stack([item[0][0] for item in batch]),
        "attention_mask": torch.stack([item[0][1] for item in batch]),
    }
    if len(batch[0]) == 1:
        return target_inputs
    pair_inputs = {
        "node_ids": torch.cat([item[1][0] for item in batch]),
        "attention_mask": torch.cat([item[1][1] for item in batch]),
    }
    num_pairs = torch.LongTensor([len(item[1][0]) for item in batch])
    labels = torch.cat([item[2] for item in batch])
    return target_inputs, pair_inputs, num_pairs, labels
