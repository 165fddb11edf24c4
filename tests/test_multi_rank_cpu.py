"""N > 1 host logic on CPU: world_size-2 gloo all_gather of hit lists, shard bookkeeping, merge order.
(The GPU side of the same path is covered by tests/test_gpu_parity.py through the shard arguments.)"""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libgwaspp_b200 import multi_gpu as mg  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_hits(rank, world, n_snps=900):
    """Deterministic pseudo-hits living in this rank's tile shard."""
    tiles, _, blk = mg.shard_tiles(n_snps, rank, world)          # the default (tensor-core) engine's schedule, from the library
    rng = np.random.default_rng(100 + rank)
    out = []
    for I, J in tiles[:: max(1, len(tiles) // 7)]:
        i = I * blk + int(rng.integers(0, blk // 2))
        j = J * blk + blk // 2 + int(rng.integers(0, blk // 2))
        if i < j < n_snps:
            out.append((i, j, 30.0 + float(rng.random()) * 10))
    return np.array(out, mg.HIT_DTYPE)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    merged = mg.gather_hits(_fake_hits(rank, world))
    q.put((rank, merged.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_gather_of_hits(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = mg.merge_hits([_fake_hits(r, world) for r in range(world)])
    assert len(want) > 5
    for r in range(world):
        assert np.array_equal(np.frombuffer(got[r], mg.HIT_DTYPE), want)     # every rank holds the same merged list
    key = want["i"].astype(np.int64) * 10_000 + want["j"]
    assert np.all(np.diff(key) > 0)                                          # the reference's (i, j) emission order


@pytest.mark.parametrize("n_snps,world,engine,n_samples", [(130, 2, 1, 4000), (130, 2, 2, 4000), (1000, 3, 1, 10_000), (1000, 3, 2, 10_000), (64, 4, 2, 500),
                                                           (50_000, 8, 1, 4000), (50_000, 8, 2, 4000), (4097, 5, 1, 10_000), (4097, 5, 2, 10_000),
                                                           (150_000, 8, 2, 10_000), (20_000, 7, 2, 40_000), (20_000, 3, 2, 200_000)])
def test_shards_partition_the_pair_space(n_snps, world, engine, n_samples):
    """The library's own shard enumeration (both engines' schedules): disjoint, complete, every pair counted once,
    and -- tensor-core engine -- balanced to one run of 64 tiles."""
    seen, total, sizes = set(), 0, []
    for r in range(world):
        tiles, pairs, blk = mg.shard_tiles(n_snps, r, world, engine, n_samples)    # band heights 16, 13 and 8 among the cases
        assert all(I <= J for I, J in tiles)
        assert not (seen & set(tiles))
        seen |= set(tiles)
        total += pairs
        sizes.append(len(tiles))
    T = (n_snps + blk - 1) // blk
    assert len(seen) == T * (T + 1) // 2
    assert total == n_snps * (n_snps - 1) // 2
    assert max(sizes) - min(sizes) <= (64 if engine == 2 else 1)


def test_merge_rejects_duplicates_and_orders_top_k():
    a = np.array([(1, 5, 31.0), (0, 9, 40.0)], mg.HIT_DTYPE)
    b = np.array([(0, 3, 35.0)], mg.HIT_DTYPE)
    m = mg.merge_hits([a, b])
    assert [(int(x["i"]), int(x["j"])) for x in m] == [(0, 3), (0, 9), (1, 5)]
    assert [float(x["stat"]) for x in mg.top_k(m, 2)] == [40.0, 35.0]
    with pytest.raises(ValueError):
        mg.merge_hits([a, a])
    assert len(mg.merge_hits([])) == 0


class _FakeStore:
    """Stands in for GenoStore on the CPU: marginal_scan(b, e) returns recognisable per-SNP rows."""
    def __init__(self, n_snps):
        self.n_snps = n_snps

    def marginal_scan(self, b, e, **kw):
        idx = np.arange(b, e, dtype=np.uint32)
        return {"counts": np.stack([idx * 8 + k for k in range(8)], 1), "stats": idx.astype(np.float64) * 0.5}


def _marginal_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    (b, e), part = mg.marginal_scan_distributed(_FakeStore(1003))
    (b2, e2), whole = mg.marginal_scan_distributed(_FakeStore(1003), gather=True)
    q.put((rank, (b, e, len(part["counts"]), b2, e2, whole["counts"].tobytes(), whole["stats"].tobytes())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_marginal_scan_ranges_and_gather(world):
    """Marginal scan over ranks: contiguous balanced SNP ranges, outputs sharded by default, whole table after a gather."""
    for n in (1, 7, 1003, 500_000):
        r = [mg.snp_range(n, k, world) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == n and all(r[k][1] == r[k + 1][0] for k in range(world - 1))
        assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_marginal_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _FakeStore(1003).marginal_scan(0, 1003)
    for r in range(world):
        b, e, n_part, b2, e2, counts, stats = got[r]
        assert (b, e) == mg.snp_range(1003, r, world) and n_part == e - b and (b2, e2) == (0, 1003)
        assert counts == want["counts"].tobytes() and stats == want["stats"].tobytes()
