"""One world cut into vertical strips, ONE PROCESS PER GPU, neighbours linked with NCCL send/recv
(frxStripInitNccl).  Launched by tests/test_gpu_strips.py (and by hand) as

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29541 tests/strip_ranks.py [--bodies 3000] [--steps 600]

Every rank builds its strip of the same seeded pile, steps it, and the ranks reduce the pile
statistics; rank 0 also runs the undivided world on its own GPU and prints one JSON line with both
(the caller asserts on it).  torch.distributed is only the rendezvous (the 128-byte NCCL id and the
final reductions); the per-stage exchange is the library's own NCCL communicator."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bodies", type=int, default=3000)
    ap.add_argument("--width", type=float, default=100.0)
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--margin", type=float, default=5.0)
    ap.add_argument("--ghost-cap", type=int, default=1024)
    ap.add_argument("--migrate", type=int, default=0, help="frxStripMigrate every this many steps (0: static ownership)")
    ap.add_argument("--push", type=float, default=0.0, help="initial x velocity of every dynamic body")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import ferox_b200 as fb
    from ferox_b200 import capi, scenes

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    L = fb.lib()
    L.frxSetDevice(local)

    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_ubyte * 128)()
        assert L.frxStripGetNcclUniqueId(buf) == 0, L.frxGetLastErrorString()
        uid = torch.frombuffer(bytearray(buf), dtype=torch.uint8).clone()
    uid = uid.to(dev)
    dist.broadcast(uid, 0)
    uid_bytes = bytes(uid.cpu().numpy().tobytes())

    spec = scenes.mixed_pile(args.bodies, width=args.width)
    if args.push:
        spec.vel[spec.body_type != capi.BODY_STATIC, 0] = args.push
    cuts = [args.width * (k + 1) / world for k in range(world - 1)]
    ss = scenes.StripScene(L, spec, cuts=cuts, margin=args.margin, ghost_cap=args.ghost_cap, rank=rank, devices={rank: local})
    assert L.frxStripInitNccl(ss.worlds[0], rank, world, uid_bytes) == 0, L.frxGetLastErrorString()
    transport = L.frxStripGetTransport(ss.worlds[0])
    ss.step(1)
    L.frxSynchronizeWorld(ss.worlds[0])
    dist.barrier()
    t0 = time.perf_counter()
    moved = 0
    if args.migrate > 0:
        for block in range(args.steps // args.migrate):
            ss.step(args.migrate)
            m = L.frxStripMigrate(ss._arr, 1)
            assert m >= 0, L.frxGetLastErrorString()
            moved += m
            if (block + 1) * args.migrate == 200:
                L.frxResetStepCounters(ss.worlds[0])     # the landing is over: from here on no body may be left behind
    else:
        ss.step(args.steps)
    L.frxSynchronizeWorld(ss.worlds[0])
    dist.barrier()
    ms = (time.perf_counter() - t0) * 1e3 / max(args.steps, 1)
    fb.check_error(L)
    # whatever this rank owns NOW (ownership may have migrated): every body that can move
    views = fb.body_state_views(L, ss.worlds[0])
    ctr = fb.step_counters(L, ss.worlds[0])
    dyn = spec.body_type != capi.BODY_STATIC
    movable = views["vel"][:, 3] > 0
    y = views["posrot"][movable, 1].astype(np.float64)
    v2 = (views["vel"][movable, :2].astype(np.float64) ** 2).sum(1)
    stray = float((ctr["overflow"] & 128) != 0)
    sums = torch.tensor([y.sum(), float(movable.sum()), v2.sum(), float(ctr["manifolds"]), float((ctr["overflow"] & ~128) != 0),
                         float(np.isfinite(views["posrot"][movable]).all()), float(moved), stray], dtype=torch.float64, device=dev)
    ymax = torch.tensor([y.max() if len(y) else -1e30], dtype=torch.float64, device=dev)
    dist.all_reduce(sums)
    dist.all_reduce(ymax, op=dist.ReduceOp.MAX)
    ss.release()
    if rank == 0:
        sc = scenes.Scene(L, spec, prime=False)
        L.frStepWorld(sc.world, scenes.DT)
        L.frxStepWorldN(sc.world, scenes.DT, args.steps)
        want = fb.body_states(L, sc.world)
        wc = fb.step_counters(L, sc.world)
        sc.release()
        s = sums.tolist()
        print(json.dumps({"ranks": world, "bodies": int(s[1]), "dynamic": int(dyn.sum()), "steps": args.steps, "ms_per_step": ms,
                          "mean_y": s[0] / s[1], "mean_y_single": float(want["pos"][dyn, 1].astype(np.float64).mean()),
                          "max_y": float(ymax.item()), "ke": s[2] / s[1],
                          "ke_single": float((want["vel"][dyn].astype(np.float64) ** 2).sum(1).mean()),
                          "manifolds": int(s[3]), "manifolds_single": int(wc["manifolds"]), "overflow": int(s[4]),
                          "finite": int(s[5]) == world, "migrated": int(s[6]), "stray_ranks": int(s[7]), "transport": transport}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
