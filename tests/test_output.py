"""tribs_b200/output.py against the files the reference writes itself: the harness dumps the raw
FP64 node fields of the last step (what tribs_gpu_*sync hand to the host) and the reference writes
its `.HHHH_MMd` spatial output from the same nodes; the writer must reproduce that text exactly."""
import glob
import os
import subprocess
import tempfile

import numpy as np
import pytest

from refrun import HARNESS, have_harness, spec_for
from tribs_b200.output import dynamic_vars_name, write_dynamic_vars
from tribs_b200.synth import write_basin
from tribs_b200.trb import read_trb

pytestmark = pytest.mark.skipif(not have_harness(), reason="oracle/_ref/tribs_ref_harness not built")


@pytest.mark.parametrize("name,hours", [("met", 6.0), ("snow", 54.0), ("deep", 7.0)])
def test_dynamic_vars_file_matches_the_reference_text(name, hours, built):
    """An output written inside the run (every SPOPINTRVL hours, between SurfaceFlow and Reset)."""
    d = tempfile.mkdtemp(prefix="tribs_out_")
    write_basin(spec_for(name, 12, runtime_h=hours + 2.0, seed=2, opt_spatial=1, extra={"SPOPINTRVL": hours}), d)
    last = int(round(hours * 16))
    r = subprocess.run([HARNESS, "basin.in", "-K", "--out", d, "--snap-from", str(last), "--snap-to", str(last),
                        "--phases", "mur"], cwd=d, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    basin, snaps = read_trb(os.path.join(d, "basin.trb")), read_trb(os.path.join(d, "steps.trb"))
    state = {}
    for ph in ("m", "u", "r"):                                     # later phases overwrite earlier ones
        pre = f"s{last}_{ph}_"
        state.update({k[len(pre):]: v for k, v in snaps.items() if k.startswith(pre)})
    ref_files = glob.glob(os.path.join(d, "out", "sp.*d"))
    want = os.path.join(d, "out", os.path.basename(dynamic_vars_name("sp", hours)))
    assert want in ref_files, (want, ref_files)
    mine = os.path.join(d, "mine.d")
    write_dynamic_vars(mine, basin, state)
    a, b = open(want).read().splitlines(), open(mine).read().splitlines()
    assert len(a) == len(b) == int(basin["n_active"][0]) + 1
    assert a[0] == b[0]
    bad = [(i, x, y) for i, (x, y) in enumerate(zip(a, b)) if x != y]
    if bad:
        i, x, y = bad[0]
        cols = [(k, p, q) for k, (p, q) in enumerate(zip(x.split(","), y.split(","))) if p != q]
        raise AssertionError(f"{len(bad)} rows differ; row {i}: columns {cols[:6]}")
    if name == "snow":
        assert any(float(row.split(",")[11]) > 0 for row in a[1:])   # IWE column carries a pack
