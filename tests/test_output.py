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


@pytest.mark.parametrize("name,hours", [("met", 3.0), ("snow", 52.0)])
def test_pixel_files_match_the_reference_text(name, hours, built):
    """The per-cell `.pixel` time series (tCOutput::WritePixelInfo): one hillslope and one stream cell,
    a row per OPINTRVL, fixed notation, the precision quirk of the first row included."""
    from tribs_b200.output import PixelWriter
    d = tempfile.mkdtemp(prefix="tribs_pix_")
    write_basin(spec_for(name, 12, runtime_h=hours, seed=2), d)
    steps = [int(round(h * 16)) for h in range(int(hours) - 2, int(hours) + 1)]       # the last three output hours
    r0 = subprocess.run([HARNESS, "basin.in", "-K", "--out", d], cwd=d, capture_output=True, text=True, timeout=900)
    assert r0.returncode == 0
    basin = read_trb(os.path.join(d, "basin.trb"))
    cells = [3, int(np.flatnonzero(basin["boundary"] == 3)[2])]
    ids = [int(basin["id"][c]) for c in cells]
    open(os.path.join(d, "nodes.pix"), "w").write(f"{len(ids)}\n" + " ".join(map(str, ids)) + "\n")
    r = subprocess.run([HARNESS, "basin.in", "-K", "--out", d, "--snap-from", str(steps[0]), "--snap-to", str(steps[-1]),
                        "--snap-every", "16", "--phases", "mur"], cwd=d, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    snaps = read_trb(os.path.join(d, "steps.trb"))
    for c, node in zip(cells, ids):
        ref = open(os.path.join(d, "out", f"sp{node}.pixel")).read().splitlines()
        assert len(ref) == int(hours) + 1
        mine = os.path.join(d, f"mine{node}.pixel")
        w = PixelWriter(mine, basin, c)
        w.first = (len(ref) - 1) == len(steps)      # False when the reference wrote earlier rows before these three
        for st in steps:
            state = {}
            for ph in ("m", "u", "r"):
                pre = f"s{st}_{ph}_"
                state.update({k[len(pre):]: v for k, v in snaps.items() if k.startswith(pre)})
            w.write_row(st / 16.0, state)
        w.close()
        got = open(mine).read().splitlines()
        assert got[0] == ref[0]
        for a, b in zip(ref[-3:], got[1:]):
            if a != b:
                cols = [(k, p, q) for k, (p, q) in enumerate(zip(a.split(), b.split())) if p != q]
                raise AssertionError(f"node {node}: {cols[:6]}\n{a[:200]}\n{b[:200]}")
    # the first row of a file has 6 decimals in its third column
    w = PixelWriter(os.path.join(d, "first.pixel"), basin, cells[0])
    st = steps[0]
    w.write_row(1.0, {k[len(f"s{st}_r_"):]: v for k, v in snaps.items() if k.startswith(f"s{st}_r_")})
    w.close()
    row = open(os.path.join(d, "first.pixel")).read().splitlines()[1].split()
    assert len(row[2].split(".")[1]) == 6 and len(row[3].split(".")[1]) == 7
