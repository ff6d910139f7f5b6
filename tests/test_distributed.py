"""World-size-2 gloo test of the multi-GPU host logic: pair sharding + the match-count gather (CPU)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, n_pairs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from imcui_b200.shard import gather_match_counts, shard_range
    lo, hi = shard_range(n_pairs, rank, world)
    local = (torch.arange(lo, hi, dtype=torch.int32) * 7 + 3) % 1000  # stand-in for per-pair match counts
    full = gather_match_counts(local, n_pairs)
    q.put((rank, lo, hi, full.tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [10, 7])
def test_shard_and_gather_gloo(n_pairs):
    import random
    port = 29500 + random.randint(0, 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    expect = [(i * 7 + 3) % 1000 for i in range(n_pairs)]
    ranges = sorted((lo, hi) for _, lo, hi, _ in res)
    assert ranges[0][0] == 0 and ranges[-1][1] == n_pairs and ranges[0][1] == ranges[1][0]
    for _, _, _, full in res:
        assert full == expect


def test_shard_range_covers_everything():
    from imcui_b200.shard import shard_range
    for n in (0, 1, 5, 64, 10000):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
