"""Mirror of ``pmgt/pmgt/utils.py`` (+ the seed helper of ``pmgt/utils/base.py``)."""
import random
from typing import List, Tuple

import numpy as np
import torch
import torch.nn as nn


def get_input_feat_embeds(node_ids: torch.LongTensor, feat_embeddings_list: nn.ModuleList) -> List[torch.Tensor]:
    """pmgt/pmgt/utils.py:43-50 -- one table lookup per modality.  Kept for API
    parity (``PMGTModel.forward`` accepts its output); ``PMGT.forward`` itself
    gathers inside the projection GEMM and never materialises these tensors."""
    return [torch.nn.functional.embedding(node_ids, e.weight) for e in feat_embeddings_list]


def load_node_init_emb(item_encoder_path: str, node_encoder_path: str, node_init_emb_path: str,
                       normalize: bool = True) -> np.ndarray:
    """pmgt/pmgt/utils.py:15-40 -- map node-order PMGT embeddings to item order for the
    downstream NCF / DCN item tables; items missing from the graph get N(0, 1) rows;
    optional row-wise L2 normalisation (sklearn ``normalize`` semantics)."""
    import joblib

    item_encoder = joblib.load(item_encoder_path)
    node_encoder = joblib.load(node_encoder_path)
    node_init_emb = np.load(node_init_emb_path)
    return remap_node_embeddings(item_encoder.classes_, node_encoder.classes_, node_init_emb, normalize)


def remap_node_embeddings(item_classes, node_classes, node_init_emb: np.ndarray, normalize: bool = True) -> np.ndarray:
    item2idx = {item: i for i, item in enumerate(node_classes)}
    out = np.empty((len(item_classes), node_init_emb.shape[1]), dtype=node_init_emb.dtype)
    for i, item in enumerate(item_classes):
        j = item2idx.get(item)
        out[i] = node_init_emb[j] if j is not None else np.random.normal(size=node_init_emb.shape[1])
    if normalize:
        norms = np.sqrt((out.astype(np.float64) ** 2).sum(axis=1))
        norms[norms == 0.0] = 1.0
        out = (out / norms[:, None]).astype(node_init_emb.dtype)
    return out


def set_seed(seed: int) -> None:
    """pmgt/utils/base.py:35-39."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
