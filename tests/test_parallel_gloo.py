"""CPU, world_size 2, gloo: the N>1 host logic of the hot path (sharding, barrier, max/sum over ranks)."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from monoflex_b200 import parallel
    r, w, _ = parallel.init("gloo")
    assert (r, w) == (rank, world)
    lo, hi = parallel.shard_range(17, r, w)
    parallel.barrier()
    mx = parallel.max_over_ranks([float(rank + 1), 10.0 - rank])
    sm = parallel.sum_over_ranks([hi - lo])
    torch.save({"shard": (lo, hi), "max": mx, "sum": sm}, os.path.join(out, "r%d.pt" % rank))
    torch.distributed.destroy_process_group()


def test_shard_barrier_and_reductions(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(str(tmp_path), "r%d.pt" % r)) for r in range(world)]
    assert res[0]["shard"] == (0, 9) and res[1]["shard"] == (9, 17)
    for r in res:
        assert r["max"] == [2.0, 10.0] and r["sum"] == [17.0]


def test_shard_range_partitions_everything():
    from monoflex_b200.parallel import shard_range
    for n in (0, 1, 7, 8, 64, 3769):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
