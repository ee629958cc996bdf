"""CPU, world_size 2, gloo: the N > 1 bookkeeping of bench.py -- independent worlds per rank, MAX over
ranks for the time, SUM for the units, value = units / max time ("weak" scaling, no collective on the
data path)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ferox_b200 import scenes, sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    mine = sharding.owned_worlds(7, rank, world_size)
    # every rank builds ITS worlds only (different seeds): nothing is exchanged about bodies
    bodies = sum(scenes.small_world(sharding.world_seed(0, w), n_dynamic=12).n for w in mine)
    elapsed_ms = 10.0 * (rank + 1)                     # rank 1 is the slow one
    t, n, extra = sharding.reduce_report(dist, torch.device("cpu"), elapsed_ms, bodies, extra_max=(rank,))
    out[rank] = (mine, bodies, t, n, extra, sharding.body_steps_per_second(n, 100, t))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in ps:
        p.start()
    for p in ps:
        p.join(120)
        assert p.exitcode == 0
    r0, r1 = out[0], out[1]
    assert r0[0] == [0, 2, 4, 6] and r1[0] == [1, 3, 5]                  # disjoint, complete
    assert sorted(r0[0] + r1[0]) == list(range(7))
    assert r0[2] == r1[2] == 20.0                                        # MAX over ranks
    assert r0[3] == r1[3] == r0[1] + r1[1] == 7 * 16                     # SUM of units
    assert r0[4] == r1[4] == [1.0]
    assert r0[5] == 7 * 16 * 100 / 0.020


def test_single_rank_needs_no_process_group():
    t, n, extra = sharding.reduce_report(None, torch.device("cpu"), 5.0, 123, extra_max=(2.0,))
    assert (t, n, extra) == (5.0, 123, [2.0])
    assert sharding.owned_worlds(5, 0, 1) == [0, 1, 2, 3, 4]
