"""Multi-GPU bookkeeping: the world step shards by INDEPENDENT WORLDS -- one
process per GPU, no data-path collective.  These helpers only decide which worlds a rank owns and
reduce the timings (MAX over ranks) and the unit counts (SUM over ranks) for the report; they work on
any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests)."""


def owned_worlds(n_worlds, rank, world_size):
    """Round-robin assignment of world indices to ranks (world w -> rank w mod world_size)."""
    return list(range(rank, n_worlds, world_size))


def world_seed(base_seed, world_index):
    return base_seed + world_index


def reduce_report(dist, device, elapsed_ms, bodies, extra_max=()):
    """MAX-reduce the elapsed time (and any extra per-rank maxima), SUM-reduce the body count."""
    import torch
    t = torch.tensor([float(elapsed_ms)] + [float(x) for x in extra_max], dtype=torch.float64, device=device)
    n = torch.tensor([float(bodies)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
    vals = [float(x) for x in t.tolist()]
    return vals[0], int(n.item()), vals[1:]


def body_steps_per_second(total_bodies, steps, elapsed_ms):
    return total_bodies * steps / (elapsed_ms * 1e-3)
