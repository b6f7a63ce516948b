"""N>1 host logic on CPU: two processes over gloo (127.0.0.1).  The device work is replaced by the
oracle here (tests may use it); what is under test is the sharding rule, the one all-gather of block
roots and the fold -- the same code shape bench.py runs over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vainplex_openclaw_b200.sharding import blocks_of, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, leaf, k, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle as O
    from vainplex_openclaw_b200 import workload as W
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = W.make_leaves(n, leaf).numpy()                       # every rank regenerates the same log
    lo, hi = shard_range(n, rank, world, align=1 << k)
    # this rank's block roots (on the GPU box: cg_merkle_block_roots_device)
    roots = []
    for s in range(lo, hi, 1 << k):
        e = min(hi, s + (1 << k))
        roots.append(np.frombuffer(O.merkle_root_fixed(data[s * leaf:], leaf, e - s), dtype=np.uint8))
    mine = torch.from_numpy(np.stack(roots).copy()) if roots else torch.zeros((0, 32), dtype=torch.uint8)
    # ragged all-gather: counts first, then padded payloads (bench.py's shards are equal-sized)
    cnt = torch.tensor([mine.shape[0]], dtype=torch.int64)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, cnt)
    mx = int(max(c.item() for c in counts))
    pad = torch.zeros((mx, 32), dtype=torch.uint8)
    pad[:mine.shape[0]] = mine
    bufs = [torch.zeros((mx, 32), dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(bufs, pad)
    allroots = np.concatenate([bufs[r][:int(counts[r].item())].numpy() for r in range(world)])
    root = O.merkle_fold(allroots)                               # on the GPU box: cg_merkle_fold_device
    # scan shards: no collective, each rank owns a slice of the result
    mlo, mhi = shard_range(1000, rank, world)
    if rank == 0:
        q.put((root, O.merkle_root_fixed(data, leaf, n), (mlo, mhi)))
    else:
        q.put((root, None, (mlo, mhi)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,k", [(5000, 6), (4096, 8), (777, 4)])
def test_sharded_merkle_root_equals_single_tree(n, k):
    world, leaf = 2, 64
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, leaf, k, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    roots = {o[0] for o in outs}
    single = [o[1] for o in outs if o[1] is not None][0]
    assert roots == {single}
    ranges = sorted(o[2] for o in outs)
    assert ranges[0][0] == 0 and ranges[-1][1] == 1000 and ranges[0][1] == ranges[1][0]


def test_shard_range_properties():
    for n in (0, 1, 7, 1000, 10_000_019):
        for world in (1, 2, 4, 8):
            for align in (1, 16, 1 << 16):
                prev = 0
                for r in range(world):
                    lo, hi = shard_range(n, r, world, align)
                    assert lo == prev and lo <= hi <= n
                    assert lo % align == 0 or lo == n
                    prev = hi
                assert prev == n
    assert blocks_of(0, 100, 4) == 7 and blocks_of(16, 16, 4) == 0


def test_library_split_rule_equals_the_host_one():
    """cg_shard_range (the split rule the C ABI exports for one-process-per-GPU callers) == sharding.shard_range for every
    rank: contiguous, disjoint, covering, every inner boundary a multiple of `align` (needs no device)."""
    from vainplex_openclaw_b200 import _native as N, sharding as S
    import random
    rnd = random.Random(5)
    for _ in range(300):
        n = rnd.choice([0, 1, 7, 1000, 65536, 10_000_000, (1 << 40) + 12345])
        world = rnd.choice([1, 2, 3, 4, 8, 16]); align = rnd.choice([1, 32, 1 << 16])
        prev = 0
        for r in range(world):
            lo, hi = N.shard_range(n, r, world, align)
            assert (lo, hi) == S.shard_range(n, r, world, align) and lo == prev and (hi == n or hi % align == 0)
            prev = hi
        assert prev == n
