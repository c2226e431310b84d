"""world_size-2 gloo tests (CPU) of the host-side data-parallel logic: rank sharding is a partition of the
global batch, the epoch permutation is identical on every rank, and 'allreduce(sum) x 1/W' of per-rank
mean-of-per-target-means gradients equals the single-process gradient of the concatenated batch.
The model math here is the fp32 oracle (CPU) -- the test is about the sharding/reduction algebra."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import model_ref
from pmgt_b200 import trainer


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    torch.set_num_threads(1)
    try:
        assert trainer.world() == (rank, ws)
        B, P, L, node_size, n_train = 4, 3, 6, 40, 64
        perm = trainer.epoch_permutation(n_train, seed=5, epoch=2)
        gathered = [None] * ws
        dist.all_gather_object(gathered, perm.tolist())
        assert all(g == gathered[0] for g in gathered)                      # same permutation on every rank
        mine = trainer.shard_indices(perm, step=1, batch_per_rank=B, rank=rank, world_size=ws)
        dist.all_gather_object(gathered, mine.tolist())
        glob = perm[1 * B * ws: 2 * B * ws].tolist()
        assert sum(gathered, []) == glob                                    # shards partition the global batch in order

        cfg = model_ref.default_cfg(hidden_size=32, intermediate_size=32, num_hidden_layers=1, feat_hidden_sizes=[16, 8])
        sd = model_ref.init_state_dict(cfg, node_size, seed=1, perturb=0.05)
        params = {k: v.requires_grad_(True) for k, v in sd.items() if not k.startswith("feat_embeddings")}

        def batch_for(targets):
            g = torch.Generator().manual_seed(int(sum(targets)))
            t = {"node_ids": torch.tensor([[2 + (x % node_size)] + [2 + ((x * 7 + j) % node_size) for j in range(L - 1)]
                                           for x in targets]), "attention_mask": torch.ones(len(targets), L)}
            p = {"node_ids": torch.randint(2, node_size + 2, (len(targets) * P, L), generator=g),
                 "attention_mask": torch.ones(len(targets) * P, L)}
            return t, p, torch.full((len(targets),), P), (torch.rand(len(targets) * P, generator=g) < 0.5).float()

        def grads_of(batches):
            for v in params.values():
                v.grad = None
            loss = sum(model_ref.pretrain_forward(sd, cfg, node_size, *b, training=False)["loss"] for b in batches) / len(batches)
            loss.backward()
            return torch.cat([(v.grad if v.grad is not None else torch.zeros_like(v)).reshape(-1) for v in params.values()])

        # per-target batches keep the pair generator independent of how targets are grouped
        local = grads_of([batch_for([int(x)]) for x in mine])
        dist.all_reduce(local)                                              # ONE allreduce of the flat gradient
        local /= ws
        full = grads_of([batch_for([int(x)]) for x in glob])
        assert torch.allclose(local, full, rtol=1e-4, atol=1e-6)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_data_parallel_sharding_and_gradient_average_gloo():
    ws = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_weak_scaling_indices_do_not_depend_on_world_size():
    perm = trainer.epoch_permutation(1000, seed=0, epoch=3)
    one = trainer.shard_indices(perm, 0, 64, 0, 1)
    two = np.concatenate([trainer.shard_indices(perm, 0, 32, r, 2) for r in range(2)])
    assert np.array_equal(one, two)
