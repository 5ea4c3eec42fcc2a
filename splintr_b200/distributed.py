"""Document-sharded encode_batch over the GPUs of one box: one process per GPU.

The reference parallelises `encode_batch` over documents with Rayon
(/root/reference/src/core/tokenizer.rs:932-934).  Documents are independent, so the
multi-GPU form is the same partition: rank r encodes a contiguous document range balanced
by cumulative input bytes, and the only exchange step is one all-gather of the per-rank
`(n_docs, n_tokens)` pair, from which every rank derives the global output offsets
(NCCL over NVLink on the GPU box; gloo in the CPU tests).  Token ids never cross ranks.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def shard_bounds(offsets: np.ndarray, world_size: int) -> np.ndarray:
    """Document boundaries dlo[0..world]: rank r owns documents [dlo[r], dlo[r+1]).
    Split points are the first document starting at or after k * N / world bytes (the same
    rule spl_encode_batch applies across the devices of one handle)."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    n_docs = len(offsets) - 1
    total = int(offsets[-1])
    dlo = np.zeros(world_size + 1, dtype=np.int64)
    for g in range(1, world_size):
        target = total // world_size * g
        d = int(np.searchsorted(offsets, np.uint64(target), side="left"))
        dlo[g] = max(min(d, n_docs), dlo[g - 1])
    dlo[world_size] = n_docs
    return dlo


def local_shard(data: np.ndarray, offsets: np.ndarray, rank: int, world_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """(bytes, offsets rebased to 0, first document index) of rank's shard."""
    dlo = shard_bounds(offsets, world_size)
    d0, d1 = int(dlo[rank]), int(dlo[rank + 1])
    b0, b1 = int(offsets[d0]), int(offsets[d1])
    return data[b0:b1], (np.asarray(offsets[d0:d1 + 1], dtype=np.uint64) - np.uint64(b0)), d0


def exchange_counts(n_docs_local: int, n_tokens_local: int, device=None) -> np.ndarray:
    """All-gather of (n_docs, n_tokens): returns int64[world, 2].  The path's only collective."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.array([[n_docs_local, n_tokens_local]], dtype=np.int64)
    mine = torch.tensor([n_docs_local, n_tokens_local], dtype=torch.int64, device=device)
    out = torch.empty(dist.get_world_size() * 2, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, mine)
    return out.cpu().numpy().reshape(-1, 2)


def encode_sharded(local_encode: Callable[[np.ndarray, np.ndarray], Tuple[np.ndarray, np.ndarray]],
                   data: np.ndarray, offsets: np.ndarray, rank: int, world_size: int,
                   device=None) -> Tuple[np.ndarray, np.ndarray, int, int, int]:
    """Encode this rank's shard of a batch every rank holds.

    `local_encode(bytes, offsets) -> (ids, out_offsets)` is the per-GPU encoder
    (`Tokenizer.encode_packed`).  Returns (ids_local, out_offsets_global for the local
    documents [n_local + 1], first_doc, token_base, total_tokens): document d of this shard
    owns global ids [out_offsets_global[i], out_offsets_global[i+1]), i = d - first_doc."""
    sb, so, d0 = local_shard(data, offsets, rank, world_size)
    ids, out_off = local_encode(sb, so)
    counts = exchange_counts(len(so) - 1, int(len(ids)), device)
    if counts.shape[0] != world_size:
        raise RuntimeError("process group size does not match world_size")
    base = int(counts[:rank, 1].sum())
    return ids, np.asarray(out_off, dtype=np.uint64) + np.uint64(base), d0, base, int(counts[:, 1].sum())
