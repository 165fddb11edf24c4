"""One-process-per-GPU driver for the pairwise screen (torch.distributed): the store replicated on every rank, the
tile-pair schedule of the screen cut into one shard per rank inside the library (gwasdev_pairwise_scan's shard
arguments: the tensor-core engine deals runs of 64 consecutive 128x128-SNP tiles of its banded schedule, the AND+POPC
engine single 64x64 tiles round-robin; `shard_tiles` asks the library for either), no data-path collective, hit lists
combined with one all_gather at the end -- NCCL over NVLink on GPUs, gloo on CPU for the host-logic tests.

The same split inside ONE process (a host thread per device, ncclAllGather of the hit records) is
gwasdev_pairwise_scan_multi, which the C++ mirror and `gwas_b200 --devices N` use; this module is the torchrun form.
The reference has no counterpart (single process, single thread); the work split follows SURVEY.md 8(e).
"""
from __future__ import annotations

import numpy as np

HIT_DTYPE = np.dtype([("i", "<u4"), ("j", "<u4"), ("stat", "<f8")])


def hits_to_tensor(hits: np.ndarray, capacity: int, device):
    """Pack a HIT_DTYPE array into a fixed-size (capacity, 2) int64 tensor (16-byte records) for all_gather."""
    import torch
    buf = np.zeros((capacity, 2), np.int64)
    buf.view(np.uint8).reshape(-1)[: hits.nbytes] = hits.view(np.uint8).reshape(-1)
    return torch.from_numpy(buf).to(device)


def tensor_to_hits(t, n: int) -> np.ndarray:
    raw = t.cpu().numpy().view(np.uint8).reshape(-1)[: n * HIT_DTYPE.itemsize]
    return np.frombuffer(raw.tobytes(), HIT_DTYPE).copy()


def merge_hits(parts) -> np.ndarray:
    """Union of per-rank hit lists in the reference's (i, j) emission order. Shards are disjoint by
    construction; duplicates would mean a sharding bug, so they are rejected."""
    parts = [p for p in parts if len(p)]
    if not parts:
        return np.zeros(0, HIT_DTYPE)
    allh = np.concatenate(parts)
    order = np.lexsort((allh["j"], allh["i"]))
    allh = allh[order]
    key = allh["i"].astype(np.uint64) << np.uint64(32) | allh["j"].astype(np.uint64)
    if len(key) > 1 and np.any(key[1:] == key[:-1]):
        raise ValueError("duplicate SNP pair across shards")
    return allh


def top_k(hits: np.ndarray, k: int) -> np.ndarray:
    """The k strongest interactions (ties broken by (i, j))."""
    order = np.lexsort((hits["j"], hits["i"], -hits["stat"]))
    return hits[order[:k]]


def gather_hits(local_hits: np.ndarray, group=None, device=None) -> np.ndarray:
    """all_gather the ranks' hit lists (counts first, then buffers padded to the largest count) and merge."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return merge_hits([local_hits])
    world = dist.get_world_size(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    counts = gather_counts(len(local_hits), world, device, group)
    cap = max(1, max(counts))
    mine = hits_to_tensor(local_hits, cap, device)
    allb = gather_records(mine, world, group)
    return merge_hits([tensor_to_hits(allb[r], n) for r, n in enumerate(counts)])


def gather_counts(n_local: int, world: int, device, group=None) -> list:
    """Every rank's hit count: one collective and ONE device-to-host read (a .item() per rank is a synchronisation each)."""
    import torch
    import torch.distributed as dist
    cnt = torch.tensor([n_local], dtype=torch.int64, device=device)
    if dist.get_backend(group) != "nccl":          # gloo (host-logic tests) has no all_gather_into_tensor
        parts = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(parts, cnt, group=group)
        return [int(c.item()) for c in parts]
    out = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, cnt, group=group)
    return out.tolist()


def gather_records(mine, world: int, group=None):
    """(cap, 2) int64 hit records of every rank -> (world, cap, 2), one collective into one buffer."""
    import torch
    import torch.distributed as dist
    mine = mine.contiguous()
    if dist.get_backend(group) != "nccl":
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        return torch.stack(parts)
    out = torch.empty((world,) + tuple(mine.shape), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(out, mine, group=group)
    return out


def snp_range(n_snps: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous SNP range of `rank` for the marginal scan: balanced to one SNP, in rank order, so that the
    concatenation of the ranks' outputs is the whole-table output."""
    return n_snps * rank // world, n_snps * (rank + 1) // world


def marginal_scan_distributed(store, group=None, gather: bool = False, **kw):
    """This rank's SNP range of the marginal scan on a store that holds the whole table (replicated, as for the pairwise
    screen). Outputs stay sharded -- no data-path collective -- unless gather=True, in which case every rank receives the
    whole-table arrays (one all_gather_object of the numpy outputs: 96 to 288 bytes per SNP)."""
    import torch.distributed as dist
    on = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if on else 0
    world = dist.get_world_size(group) if on else 1
    b, e = snp_range(store.n_snps, rank, world)
    out = store.marginal_scan(b, e, **kw)
    if not gather or world == 1:
        return (b, e), out
    parts = [None] * world
    dist.all_gather_object(parts, out, group=group)
    return (0, store.n_snps), {k: np.concatenate([p[k] for p in parts]) for k in out}


def shard_tiles(n_snps: int, shard: int, n_shards: int, engine: int = 2, n_samples: int = 10_000):
    """Tile pairs (I <= J, SNP-block indices) `shard` handles and the pairs they cover, from the library's own schedule
    (gwasdev_shard_schedule; host arithmetic, no GPU needed). engine 2 (default, tensor cores): blocks of 128 SNPs, banded
    for a table of n_samples individuals; engine 1 (AND+POPC): blocks of 64. Returns (list of (I, J), pairs covered, SNPs per block)."""
    import libgwaspp_b200 as gw
    tiles, pairs = gw.shard_schedule(n_snps, shard, n_shards, engine, n_samples)
    return [(int(I), int(J)) for I, J in tiles], pairs, 128 if engine == 2 else 64


def pairwise_scan_distributed(store, threshold: float = 30.0, group=None):
    """Run this rank's shard of the exhaustive screen and return the merged, (i, j)-ordered hit list on
    every rank together with this rank's PairStats."""
    import torch.distributed as dist
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    hits, stats = store.pairwise_scan(threshold, shard=rank, n_shards=world)
    return gather_hits(hits, group), stats
