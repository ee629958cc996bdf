"""Where a coloured solve spends its time: stage stamps of k_solve_colored (FEROX_SOLVE_TRACE=1, frxGetSolveTrace).

    FEROX_SOLVE_TRACE=1 python tools/solve_trace.py [config2|config3] [settle_steps]
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("FEROX_SOLVE_TRACE", "1")
import ferox_b200 as fb                      # noqa: E402
from ferox_b200 import scenes                 # noqa: E402
import bench                                  # noqa: E402

NAMES = ["body state cleared", "inherited colours validated", "colouring rounds", "colours counted", "ordered by colour",
         "records + warm start"] + ["sweep %d" % i for i in range(10)]


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "config2"
    settle = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    L = fb.lib()
    L.frxSetDevice(0)
    if which == "config3":
        spec, _, _ = scenes.strip_pile(1 << 20, 0, 1, seed=1, storeys=8, max_rows=16, circle_fraction=1.0)   # bench.py config3_shallow
        sc = bench.build_bulk_world(L, spec)
    else:
        sc = scenes.Scene(L, scenes.mixed_pile(65536), prime=False)
    L.frxStepWorldN(sc.world, scenes.DT, settle)
    fb.check_error(L)
    rows = []
    out = (C.c_ulonglong * 32)()
    for _ in range(5):
        L.frStepWorld(sc.world, scenes.DT)
        n = L.frxGetSolveTrace(sc.world, out, 32)
        assert n == 32, "tracing is off (FEROX_SOLVE_TRACE=1 must be set before the world is created)"
        t = np.array(out[:17], dtype=np.float64)
        rows.append(np.diff(t) / 1000.0)
    med = np.median(np.array(rows), axis=0)
    ctr = fb.step_counters(L, sc.world)
    print("%s: %d bodies, %d manifolds, %d colours, %d colouring rounds" % (which, ctr["bodies"], ctr["manifolds"], ctr["colors"], ctr.get("colorRounds", -1)))
    for name, us in zip(NAMES, med):
        print("  %-30s %9.1f us" % (name, us))
    print("  %-30s %9.1f us (write-back not included)" % ("total", med.sum()))


if __name__ == "__main__":
    main()
