"""GPU: strip decomposition of one world (frxStrip*, SURVEY.md section 8e / BASELINE config 5) with
in-process neighbours on ONE GPU: the same pack / ghost / per-stage reconciliation code path that
the NCCL transport drives across GPUs (the last test launches that one, two processes, when the box
has two GPUs).  Cross-cut pairs are solved on the left strip only; velocity changes of ghosts are
returned to their owner after every solver stage (Jacobi coupling across a cut, mass-split), so the
comparison with the undivided world is statistical, like the coloured-vs-serial comparison."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import ferox_b200 as fb
from ferox_b200 import capi, scenes

pytestmark = pytest.mark.gpu
DT = scenes.DT


def undivided(L, spec, steps):
    sc = scenes.Scene(L, spec, prime=False)
    L.frStepWorld(sc.world, DT)
    L.frxStepWorldN(sc.world, DT, steps)
    st = fb.body_states(L, sc.world)
    c = fb.step_counters(L, sc.world)
    sc.release()
    return st, c


def test_strips_with_no_interaction_across_the_cut_match_the_single_world(gpu):
    """Two piles far from the cut: no ghost ever interacts, every strip must evolve exactly like the
    same bodies in one world as far as the pair set and contact geometry go, and end within the
    ordering tolerance (colours differ because manifold indices differ)."""
    L = gpu
    a = scenes.mixed_pile(400, width=30.0, seed=5)
    spec = scenes.mixed_pile(400, width=30.0, seed=5)
    # second pile shifted 200 units to the right, merged into one spec
    b = scenes.mixed_pile(400, width=30.0, seed=6)
    b.pos = b.pos + np.array([200.0, 0.0], np.float32)
    spec.shapes = a.shapes + b.shapes
    spec.body_type = np.concatenate([a.body_type, b.body_type])
    spec.body_shape = np.concatenate([a.body_shape, b.body_shape + len(a.shapes)])
    spec.pos = np.concatenate([a.pos, b.pos]); spec.angle = np.concatenate([a.angle, b.angle])
    spec.vel = np.concatenate([a.vel, b.vel]); spec.omega = np.concatenate([a.omega, b.omega])
    ss = scenes.StripScene(L, spec, cuts=[100.0], margin=4.0, ghost_cap=256)
    ss.step(1)                                  # priming
    ss.step(600)
    fb.check_error(L)
    got = ss.gather_states()
    want, _ = undivided(L, spec, 600)
    dyn = spec.body_type != capi.BODY_STATIC
    assert np.isfinite(got["pos"][dyn]).all()
    my, wy = got["pos"][dyn, 1].mean(), want["pos"][dyn, 1].mean()
    assert abs(my - wy) < 0.02 * abs(wy) + 0.05, (my, wy)
    assert got["pos"][dyn, 1].max() < 0.6
    for w in ss.worlds:
        assert fb.step_counters(L, w)["overflow"] == 0
    ss.release()


def test_two_and_three_strips_through_a_pile_match_the_single_world_statistically(gpu):
    L = gpu
    spec = scenes.mixed_pile(3000, width=100.0)
    want, wc = undivided(L, spec, 600)
    dyn = spec.body_type != capi.BODY_STATIC
    masses = None
    for cuts in ([50.0], [33.0, 66.0]):
        ss = scenes.StripScene(L, spec, cuts=cuts, margin=5.0, ghost_cap=1024)
        ss.step(1)
        ss.step(600)
        fb.check_error(L)
        got = ss.gather_states()
        assert np.isfinite(got["pos"][dyn]).all(), cuts
        ctrs = [fb.step_counters(L, w) for w in ss.worlds]
        # bit 128 = "a body lies outside its strip + margin": ownership is static and a landing pile throws a few
        # bodies that far; everything else (capacity overflows) must stay clear
        assert all(c["overflow"] & ~128 == 0 for c in ctrs)
        # every dynamic body is owned exactly once
        assert sum(len(ss.owned[k]) for k in ss.local) - (~dyn).sum() * len(ss.worlds) == dyn.sum()
        # pile height / potential energy and kinetic energy level agree with the undivided world
        my, wy = got["pos"][dyn, 1].mean(), want["pos"][dyn, 1].mean()
        assert abs(my - wy) < 0.02 * abs(wy) + 0.05, (cuts, my, wy)
        assert got["pos"][dyn, 1].max() < 0.6                      # nothing fell through the floor (top at y = 0)
        ke_s = (got["vel"][dyn] ** 2).sum(1).mean()
        ke_w = (want["vel"][dyn] ** 2).sum(1).mean()
        assert ke_s < 4.0 * ke_w + 1.0, (cuts, ke_s, ke_w)
        # manifolds: a cross-cut manifold exists on the left strip only; stale ghost entries add a few
        m = sum(c["manifolds"] for c in ctrs)
        assert wc["manifolds"] * 0.97 < m < wc["manifolds"] * 1.25, (cuts, m, wc["manifolds"])
        ss.release()


@pytest.mark.parametrize("peer", ["1", "0"], ids=["peer_direct", "nccl_send_recv"])
def test_two_strips_on_two_gpus_over_nccl(gpu, peer):
    """One process per GPU (tests/strip_ranks.py), both transports: peer-direct (messages stored straight into the
    neighbour's IPC-mapped receive buffer over NVLink and announced with a flag) and the library's NCCL send/recv."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(root, "tests", "strip_ranks.py"), "--bodies", "3000", "--steps", "600"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root, env=dict(os.environ, FEROX_STRIP_PEER=peer))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    r = json.loads(line)
    assert r["transport"] == (2 if peer == "1" else 1), r
    assert r["finite"] and r["overflow"] == 0 and r["bodies"] == r["dynamic"], r
    assert abs(r["mean_y"] - r["mean_y_single"]) < 0.02 * abs(r["mean_y_single"]) + 0.05, r
    assert r["max_y"] < 0.6 and r["ke"] < 4.0 * r["ke_single"] + 1.0, r
    assert r["manifolds_single"] * 0.97 < r["manifolds"] < r["manifolds_single"] * 1.25, r


def test_migration_over_nccl_on_two_gpus(gpu):
    """The same sloshing pile with one process per GPU: migrants travel as rows over the library's NCCL communicator and
    are re-created on the receiving rank; every body stays owned by exactly one rank and none is left behind."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29543", os.path.join(root, "tests", "strip_ranks.py"), "--bodies", "3000", "--steps", "600", "--migrate", "10",
           "--push", "8.0", "--margin", "6.0"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    r = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert r["finite"] and r["overflow"] == 0 and r["stray_ranks"] == 0 and r["bodies"] == r["dynamic"] and r["migrated"] > 20, r
    # (a pile thrown against a wall is still sloshing after 600 steps, and the Jacobi coupling across a cut damps that
    # more slowly than one world does: the kinetic energy is only bounded loosely here)
    assert r["max_y"] < 0.6 and r["ke"] < 10.0 * r["ke_single"] + 2.0, r


def test_eight_strips_of_the_65k_pile_match_the_single_world_statistically(gpu):
    """'parity shown by running the same code path at 65 k bodies on 2-8
    strips vs 1 GPU'.  The BASELINE config-2 pile (65,536 mixed bodies, 1,160 units wide) cut into 8
    strips, all on this GPU with in-process links, 400 steps."""
    L = gpu
    spec = scenes.mixed_pile(65536)
    steps = 400
    want, wc = undivided(L, spec, steps)
    dyn = spec.body_type != capi.BODY_STATIC
    width = float(spec.pos[dyn, 0].max() - spec.pos[dyn, 0].min())
    x0 = float(spec.pos[dyn, 0].min())
    cuts = [x0 + width * k / 8.0 for k in range(1, 8)]
    ss = scenes.StripScene(L, spec, cuts=cuts, margin=5.0, ghost_cap=2048)
    ss.step(1)
    ss.step(steps)
    fb.check_error(L)
    got = ss.gather_states()
    ctrs = [fb.step_counters(L, w) for w in ss.worlds]
    assert all(c["overflow"] & ~128 == 0 for c in ctrs)
    assert np.isfinite(got["pos"][dyn]).all()
    my, wy = got["pos"][dyn, 1].mean(), want["pos"][dyn, 1].mean()
    assert abs(my - wy) < 0.01 * abs(wy) + 0.05, (my, wy)
    assert got["pos"][dyn, 1].max() < 0.6
    ke_s = (got["vel"][dyn] ** 2).sum(1).mean()
    ke_w = (want["vel"][dyn] ** 2).sum(1).mean()
    assert ke_s < 2.0 * ke_w + 0.5, (ke_s, ke_w)
    m = sum(c["manifolds"] for c in ctrs)
    assert wc["manifolds"] * 0.97 < m < wc["manifolds"] * 1.1, (m, wc["manifolds"])
    # no hot spot at any cut: kinetic energy of the bodies within 10 units of a cut vs the rest
    x = got["pos"][dyn, 0]
    near = np.zeros(x.shape, bool)
    for c in cuts:
        near |= np.abs(x - c) < 10.0
    v2 = (got["vel"][dyn] ** 2).sum(1)
    assert v2[near].mean() < 3.0 * v2[~near].mean() + 0.5, (v2[near].mean(), v2[~near].mean())
    ss.release()


def test_ghost_overflow_and_stray_bodies_are_reported(gpu):
    """Static ownership has limits: too few ghost rows (bit 32) and a body that left its strip by more
    than the margin (bit 128) must be reported, not silently wrong."""
    L = gpu
    spec = scenes.mixed_pile(3000, width=100.0)
    ss = scenes.StripScene(L, spec, cuts=[50.0], margin=5.0, ghost_cap=16)      # a 51-row pile needs ~150 ghost rows
    ss.step(3)
    for w in ss.worlds:
        L.frxSynchronizeWorld(w)
    assert any(fb.step_counters(L, w)["overflow"] & 32 for w in ss.worlds)
    assert L.frxGetLastError() == 2                       # lost work is an error, sticky until cleared
    L.frxClearLastError()
    ss.release()
    ss = scenes.StripScene(L, spec, cuts=[50.0], margin=5.0, ghost_cap=1024)
    ss.step(2)
    for w in ss.worlds:
        L.frxSynchronizeWorld(w)
    assert all(fb.step_counters(L, w)["overflow"] == 0 for w in ss.worlds)
    i = int(ss.owned[0][-1])                                                      # a dynamic body of the left strip ...
    L.frSetBodyPosition(ss.bodies[(0, i)], capi.Vector2(90.0, -30.0))             # ... teleported deep into the right one
    ss.step(2)
    for w in ss.worlds:
        L.frxSynchronizeWorld(w)
    assert fb.step_counters(L, ss.worlds[0])["overflow"] & 128
    ss.release()


def test_migration_keeps_a_sloshing_pile_whole(gpu):
    """frxStripMigrate: a pile that starts out moving sideways carries hundreds of bodies across both cuts.  With the
    ownership re-assigned every 10 steps every body is owned by exactly one strip at the end, the pile ends where the
    undivided world puts it, and once the landing is over (a landing pile throws single bodies further than the margin
    within 10 steps) no body is ever outside its strip + margin again (bit 128 stays clear for the last 400 steps)."""
    L = gpu
    spec = scenes.mixed_pile(3000, width=100.0)
    dyn = spec.body_type != capi.BODY_STATIC
    spec.vel[dyn, 0] = 8.0
    ss = scenes.StripScene(L, spec, cuts=[100.0 / 3.0, 200.0 / 3.0], margin=6.0, ghost_cap=1024)
    index_of = {h: i for (k, i), h in ss.bodies.items() if dyn[i]}
    ss.step(1)
    moved = 0
    for block in range(60):
        ss.step(10)
        m = L.frxStripMigrate(ss._arr, len(ss.worlds))
        assert m >= 0, L.frxGetLastErrorString()
        moved += m
        if block == 19:
            for w in ss.worlds:
                L.frxResetStepCounters(w)
    fb.check_error(L)
    pos = np.full((spec.n, 2), np.nan, np.float32)
    vel = np.full((spec.n, 2), np.nan, np.float32)
    owners = np.zeros(spec.n, np.int32)
    for w in ss.worlds:
        st = fb.body_states(L, w)
        for j in range(L.frGetBodyCountInWorld(w)):
            i = index_of.get(L.frGetBodyInWorld(w, j))
            if i is not None:
                pos[i], vel[i] = st["pos"][j], st["vel"][j]
                owners[i] += 1
        c = fb.step_counters(L, w)
        assert c["overflow"] == 0, c                 # neither a dropped ghost (32) nor a body left behind (128) since step 200
    assert (owners[dyn] == 1).all() and moved > 50, (moved, np.bincount(owners[dyn]))
    assert np.isfinite(pos[dyn]).all()
    want, wc = undivided(L, spec, 600)
    for axis in (0, 1):
        a, b = pos[dyn, axis].astype(np.float64).mean(), want["pos"][dyn, axis].astype(np.float64).mean()
        assert abs(a - b) < 0.02 * abs(b) + 0.25, (axis, a, b)
    assert pos[dyn, 1].max() < 0.6                    # nothing fell through the floor
    ke = lambda v: float((v.astype(np.float64) ** 2).sum(1).mean())
    assert ke(vel[dyn]) < 4.0 * ke(want["vel"][dyn]) + 0.5
    ss.release()
