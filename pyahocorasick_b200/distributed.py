"""Multi-GPU batch search: the haystack batch shards embarrassingly, the flattened automaton is
replicated on every GPU, and the only exchange is one all-gather of the per-rank match counts
(SURVEY.md section 8(e)).  One process per GPU, torch.distributed for the plumbing (NCCL over
NVLink on GPUs, gloo in the CPU tests).  No collective touches the haystack bytes or the records
unless the caller asks for a gathered record list.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_bounds(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n_items for `rank` (first n_items % world ranks get one extra)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class ShardedMatches:
    """This rank's matches with GLOBAL haystack ids, plus the all-gathered counts.

    counts[r]  = number of matches found by rank r
    offsets[r] = where rank r's records start in the concatenated global list (exclusive scan)
    """

    def __init__(self, hay_id, end_index, key_id, values, counts, rank):
        self.hay_id, self.end_index, self.key_id = hay_id, end_index, key_id
        self._values = values
        self.counts = counts
        self.offsets = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
        self.total = int(counts.sum())
        self.rank = rank

    def __len__(self):
        return len(self.hay_id)

    def values(self):
        v = self._values
        return [v[k] for k in self.key_id.tolist()]

    def records(self):
        return list(zip(self.hay_id.tolist(), self.end_index.tolist(), self.values()))


def scan_sharded(A, haystacks: np.ndarray, group=None, device: Optional[int] = None, algo: str = "auto",
                 already_local: bool = False, n_global: Optional[int] = None) -> ShardedMatches:
    """Search `haystacks` (uint8 [n, stride]) across the ranks of `group`.

    Every rank passes the same global array (each scans only its contiguous shard), or, with
    already_local=True, just its own shard (then n_global = total number of haystacks).
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if already_local:
        if n_global is None:
            raise ValueError("n_global is required with already_local=True")
        lo, hi = shard_bounds(n_global, world, rank)
        local = haystacks
        if local.shape[0] != hi - lo:
            raise ValueError(f"rank {rank}: shard has {local.shape[0]} haystacks, expected {hi - lo}")
    else:
        lo, hi = shard_bounds(haystacks.shape[0], world, rank)
        local = haystacks[lo:hi]
    m = A.find_all_batch(np.ascontiguousarray(local), algo=algo, device=device)
    backend = dist.get_backend(group) if dist.is_initialized() else "none"
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.tensor([len(m)], dtype=torch.int64, device=dev)
    if world > 1:
        allc = torch.zeros(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, mine, group=group)       # the one collective of the path
        counts = allc.cpu().numpy()
    else:
        counts = mine.cpu().numpy()
    return ShardedMatches(m.hay_id.astype(np.int64) + lo, m.end_index, m.key_id, m._values, counts, rank)


def gather_records(sm: ShardedMatches, group=None) -> Optional[np.ndarray]:
    """Optional payload gather: every rank receives all records as an int64 [total, 3] array
    (hay_id, end_index, key_id) in rank order = global haystack order."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return np.stack([sm.hay_id, sm.end_index.astype(np.int64), sm.key_id.astype(np.int64)], axis=1)
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    cap = int(sm.counts.max())
    buf = torch.zeros((cap, 3), dtype=torch.int64, device=dev)
    if len(sm):
        rec = np.stack([sm.hay_id, sm.end_index.astype(np.int64), sm.key_id.astype(np.int64)], axis=1)
        buf[:len(sm)] = torch.from_numpy(rec).to(dev)
    out = torch.zeros((world * cap, 3), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.cpu().numpy().reshape(world, cap, 3)
    return np.concatenate([out[r, :int(sm.counts[r])] for r in range(world)], axis=0)
