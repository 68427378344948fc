"""CPU tier, world_size 2 over gloo: the host-side sharding logic of the multi-GPU paths
(independent-problem blocks for C4, residual-row partition for C5, result gathering)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from saddlepoint_b200 import dist as sd
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # C4: 11 independent problems over 2 ranks; each rank "solves" its block, results are merged
        lo, hi = sd.shard_range(11, world, rank)
        mine = [(i, {"problem": i, "iters": 100 + i}) for i in range(lo, hi)]
        merged = sd.merge_problem_results(sd.gather_objects(mine, dist))
        assert [r["problem"] for r in merged] == list(range(11))
        # C5: row partition of a 2n-row Jacobian, pair-aligned
        m = 2 * 37
        rb, re = sd.row_partition(m, world, rank, align=2)
        blocks = sd.gather_objects((rb, re), dist)
        assert blocks[0][0] == 0 and blocks[-1][1] == m
        assert all(blocks[k][1] == blocks[k + 1][0] for k in range(world - 1))
        assert all(b[0] % 2 == 0 for b in blocks)
        # the owned value subsequences tile the global CSC value array
        rng = np.random.default_rng(0)
        rows = np.sort(rng.integers(0, m, 200)).astype(np.int32)
        vals = rng.standard_normal(200)
        colptr = np.array([0, 200], np.int32)
        loc = sd.local_values(colptr, rows, vals, rb, re)
        tot = sd.gather_objects(float(loc.sum()), dist)
        assert abs(sum(tot) - vals.sum()) < 1e-12
        # the unique-id broadcast plumbing (object broadcast) works over the process group
        box = [b"x" * 128 if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        assert box[0] == b"x" * 128
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_sharding_logic_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_helpers_single_process():
    sys.path.insert(0, ROOT)
    from saddlepoint_b200 import dist as sd
    for n, w in [(4096, 8), (10, 3), (2, 4), (0, 2)]:
        spans = [sd.shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[k][1] == spans[k + 1][0] for k in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sd.merge_problem_results([[(0, "a")], [(2, "b")]])


def test_column_partition_and_ring_for_the_halo_mode():
    from saddlepoint_b200 import dist as sd
    n, halo = 1000, 40
    for world in (2, 3, 8):
        blocks = [sd.column_partition(n, world, r, halo) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
        assert all(hi - lo >= 2 * halo for lo, hi in blocks)
        for r in range(world):
            left, right = sd.ring_neighbours(world, r)
            assert sd.ring_neighbours(world, left)[1] == r and sd.ring_neighbours(world, right)[0] == r
    import pytest
    with pytest.raises(ValueError):
        sd.column_partition(n, 16, 0, halo)  # 62 unknowns per rank < 2*halo
