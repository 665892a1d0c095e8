"""Full-size golden vectors: BASELINE.json configs[1] (the basin and storm bench.py times:
build_basin(1000, 1000, seed=1), 1M Voronoi cells, 20 mm/h for 6 h) advanced 128 steps by the C
oracle (oracle/tribs_oracle.c, pinned bit for bit to the reference on the small decks). The storm
ends at step 96, so the window contains the storm -> interstorm regime change.

Kept per snapshot (steps 64 and 128, after Reset): a fixed sample of cells (every stream cell + a
strided sample of the hillslopes) of the state fields, sums over blocks of 1000 consecutive cells
of every field (full coverage at block resolution), the hydrograph arrays and the basin sums.

    python tests/golden/make_fullsize.py            # ~5 min, writes tests/golden/fullsize_1m.npz"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

SIDE, SEED, SNAP_STEPS, RUNTIME_H = 1000, 1, (64, 128), 24.0   # RUNTIME only sizes the hydrograph arrays
STATE = ["NwtOld", "MuOld", "MiOld", "NtOld", "NfOld", "RuOld", "RiOld", "Qpout", "cumsrf", "intstorm"]
RESULTS = ["mhydro", "HsrfRout", "SbsrfRout", "SatsrfRout"]
BLOCK = 1000


def sample_cells(basin):
    n = int(basin["n_active"][0])
    stream = np.nonzero(np.asarray(basin["boundary"]) == 3)[0]
    strided = np.arange(0, n, 251)
    return np.unique(np.concatenate([stream, strided])).astype(np.int64)


def block_sums(a):
    n = (a.size // BLOCK) * BLOCK
    return a[:n].reshape(-1, BLOCK).sum(axis=1)


def main(side=SIDE, out=None):
    from bench import storm_rain
    from common import OracleStepper
    from tribs_b200.synthbasin import build_basin
    t0 = time.time()
    basin = build_basin(side, side, seed=SEED, runtime_h=RUNTIME_H)
    o = OracleStepper(basin, storm_rain)
    idx = sample_cells(basin)
    z = {"side": np.array([side]), "seed": np.array([SEED]), "snap_steps": np.array(SNAP_STEPS), "cells": idx,
         "n_active": np.array([o.n])}
    for st in range(1, max(SNAP_STEPS) + 1):
        o.advance(); o.unsat()
        if o.gw_due():
            o.sat()
        o.route(); o.balance(); o.reset()
        if st in SNAP_STEPS:
            got = o.get(STATE)
            for f in STATE:
                z[f"s{st}/{f}"] = got[f][idx].copy()
                z[f"s{st}/blk/{f}"] = block_sums(got[f])
            for r in RESULTS:
                z[f"s{st}/res/{r}"] = np.array(o.result(r)).copy()
            g = o.globals()
            z[f"s{st}/glob_names"] = np.array(sorted(g))
            z[f"s{st}/glob"] = np.array([g[k] for k in sorted(g)], float)
            print(f"step {st}: {time.time() - t0:.0f} s", flush=True)
    out = out or os.path.join(HERE, f"fullsize_{side // 1000}m.npz" if side >= 1000 else f"fullsize_{side}.npz")
    np.savez_compressed(out, **z)
    print("wrote", out, os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else SIDE)
