"""Full-size parity (BASELINE.json configs[1]: the 1M-cell basin and storm that bench.py times).

tests/golden/fullsize_1m.npz holds what the UNMODIFIED REFERENCE leaves after 64 and 128 steps of that
basin (it reads the basin as OPTMESHINPUT 9 files written by synthbasin.build_basin(mesh_dir=...),
tests/golden/make_fullsize.py; the C oracle reproduces those runs bit for bit, checked by the
generator on the full arrays). The window contains the end of the storm, when most columns change
state. Kept: a sample of cells (all stream cells + a strided hillslope sample), 1000-cell block sums
of every state field, the hydrograph arrays, and the sparse patches that turn build_basin()'s dict
into the reference's own flattened inputs. The device runs the same 128 steps in pipelined batches
of 64 (the configuration bench.py measures).

Bar (north_star): per-cell Nwt, Nf, Mu ... and the hydrograph within 1e-6 relative."""
import os

import numpy as np
import pytest

from common import GpuStepper
from refrun import ROOT

pytestmark = pytest.mark.gpu

REL_TOL = 1e-6
FIXTURE = os.path.join(ROOT, "tests", "golden", "fullsize_1m.npz")


def test_gpu_full_size_basin_matches_oracle_fixture(built):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from bench import storm_rain
    from make_fullsize import BLOCK, RESULTS, RUNTIME_H, STATE, apply_patches, block_sums
    from tribs_b200.synthbasin import build_basin
    z = np.load(FIXTURE)
    side, seed = int(z["side"][0]), int(z["seed"][0])
    basin = apply_patches(build_basin(side, side, seed=seed, runtime_h=RUNTIME_H), z)
    assert int(basin["n_active"][0]) == int(z["n_active"][0]) == side * side and int(z["oracle_bitexact"][0]) == 1
    idx = z["cells"]
    g = GpuStepper(basin, storm_rain)
    done, worst = 0, {}
    for st in [int(s) for s in z["snap_steps"]]:
        while done < st:
            k = min(64, st - done)
            g.sim.run_batch(k)
            done += k
        got = g.get(list(STATE))
        for f in STATE:
            ref = z[f"s{st}/{f}"]
            err = np.abs(ref - got[f][idx]) / np.maximum(np.abs(ref), 1e-6)
            worst[(st, f)] = float(err.max())
            ref_b, got_b = z[f"s{st}/blk/{f}"], block_sums(got[f])
            scale = np.maximum(np.abs(ref_b), 1e-6 * BLOCK)
            worst[(st, "blk", f)] = float((np.abs(ref_b - got_b) / scale).max())
        for r in RESULTS:
            ref = z[f"s{st}/res/{r}"]
            dev = np.asarray(g.result(r))[: ref.size]
            worst[(st, "res", r)] = float(np.abs(ref - dev).max() / max(np.abs(ref).max(), 1e-6))
    rows = sorted(worst.items(), key=lambda kv: -kv[1])
    print("\n[full size] worst relative differences:\n" + "\n".join(f"  {k}: {v:.3e}" for k, v in rows[:10]))
    assert rows[0][1] <= REL_TOL, rows[:10]
    # the sample must have seen the physics: wetting fronts, lateral outflow, runoff, discharge
    assert np.abs(z["s128/Qpout"]).max() > 0 and np.abs(z["s128/cumsrf"]).max() > 0 and z["s128/res/mhydro"].max() > 0
