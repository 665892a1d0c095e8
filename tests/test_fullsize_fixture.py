"""CPU checks of the full-size golden fixture (tests/golden/fullsize_1m.npz, made by
tests/golden/make_fullsize.py from a run of the unmodified reference on the 1M-cell bench basin): it
belongs to the basin bench.py times, the oracle reproduced it bit for bit, and it contains the
physics the GPU test is supposed to compare (tests/test_zz_fullsize_gpu.py)."""
import os
import sys

import numpy as np

from refrun import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def test_fullsize_fixture_belongs_to_the_bench_basin():
    from make_fullsize import BLOCK, RUNTIME_H, STATE, apply_patches, sample_cells
    from tribs_b200.synthbasin import build_basin
    z = np.load(os.path.join(ROOT, "tests", "golden", "fullsize_1m.npz"))
    side, seed = int(z["side"][0]), int(z["seed"][0])
    assert (side, seed) == (1000, 1)                     # bench.py: build_basin(args.side = 1000, ..., seed = 1 + rank)
    raw = build_basin(side, side, seed=seed, runtime_h=RUNTIME_H)
    basin = apply_patches(raw, z)
    n = int(basin["n_active"][0])
    assert int(z["oracle_bitexact"][0]) == 1             # the generator compared the oracle with the reference's full arrays
    assert 1e5 < float(z["ref_cell_steps_per_s"][0]) < 1e7 and float(z["ref_setup_s"][0]) > 0
    # the patches are what the reference derives itself: hull-cell areas, stream-node convention, array length
    patched = sorted(k.split("/", 2)[2] for k in z.files if k.startswith("patch/idx/") or k.startswith("patch/full/"))
    assert set(patched) <= {"varea", "stream_node", "BasArea", "res_limit"}, patched
    assert z["patch/idx/varea"].size <= 4 * side and np.abs(basin["varea"] - raw["varea"]).max() < 0.05 * raw["varea"].max()
    assert n == int(z["n_active"][0]) == 1_000_000
    assert np.array_equal(z["cells"], sample_cells(basin))
    stream = np.nonzero(np.asarray(basin["boundary"]) == 3)[0]
    assert np.isin(stream, z["cells"]).all() and z["cells"].size > 4000
    for st in (64, 128):
        for f in STATE:
            assert z[f"s{st}/{f}"].shape == z["cells"].shape and np.isfinite(z[f"s{st}/{f}"]).all()
            assert z[f"s{st}/blk/{f}"].shape == (n // BLOCK,)
    # storm until step 96, then dry: fronts moved, lateral flow and runoff happened, the river carries water
    assert (z["s128/NfOld"] > 0).mean() > 0.5 and z["s128/Qpout"].max() > 0 and z["s128/cumsrf"].max() > 0
    assert z["s128/res/mhydro"].max() > z["s64/res/mhydro"].max() > 0
    assert not np.array_equal(z["s64/NwtOld"], z["s128/NwtOld"])
