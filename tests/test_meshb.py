"""OPTMESHINPUT 9 ("MeshBuilder") files (tribs_b200/meshb.py, synthbasin.build_basin(mesh_dir=...))
against the reference itself (oracle/_ref/tribs_ref_harness = the unmodified sources)."""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from refrun import HARNESS, ROOT, have_harness, spec_for
from tribs_b200.meshb import EDGE_DTYPE, NODE_DTYPE, read_meshb, write_meshb
from tribs_b200.synth import write_basin
from tribs_b200.trb import read_trb

pytestmark = pytest.mark.skipif(not have_harness(), reason="oracle/_ref/tribs_ref_harness not built")
SNAP = ["--snap-from", "1", "--snap-to", "64", "--snap-every", "9", "--phases", "ugr"]


def _run(d, deck, out, *extra):
    os.makedirs(os.path.join(out, "out"), exist_ok=True)
    r = subprocess.run([HARNESS, deck, "-K", "--out", out, *SNAP, *extra], cwd=d, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    return read_trb(os.path.join(out, "basin.trb")), read_trb(os.path.join(out, "steps.trb"))


def test_reference_reproduces_itself_from_its_own_meshb_files(built):
    """Format check: the mesh + flow network the reference built from points (OPTMESHINPUT 8), dumped
    in the MeshBuilder formats, read back by meshb.py, written again and fed to the reference with
    OPTMESHINPUT 9 must give the same flattened basin and the same run, bit for bit."""
    d = tempfile.mkdtemp(prefix="tribs_meshb_")
    write_basin(spec_for("deep", 14, runtime_h=4.0, seed=3), d)
    b8, s8 = _run(d, "basin.in", d, "--dump-meshb", d)
    m = read_meshb(d)
    assert m["nodes"].dtype == NODE_DTYPE and m["edges"].dtype == EDGE_DTYPE and m["nodes"].size == 14 * 14
    assert np.array_equal(m["edges"]["orig"][0::2], m["edges"]["dest"][1::2])          # complementary pairs are adjacent
    raw = {f: open(os.path.join(d, f + ".meshb"), "rb").read() for f in ("nodes", "edges", "flow")}
    write_meshb(d, m["nodes"], m["edges"], m["flow"])                                    # reader -> writer round trip
    for f, blob in raw.items():
        assert open(os.path.join(d, f + ".meshb"), "rb").read() == blob, f
    txt = re.sub(r"(OPTMESHINPUT:\s*\n)\s*\d+", r"\g<1>9", open(os.path.join(d, "basin.in")).read())
    open(os.path.join(d, "basin9.in"), "w").write(txt)
    b9, s9 = _run(d, "basin9.in", os.path.join(d, "o9"))
    assert set(b8) == set(b9) and set(s8) == set(s9) and len(s8) > 300
    for k in b8:
        assert np.array_equal(b8[k], b9[k], equal_nan=True), k
    for k in s8:
        assert np.array_equal(s8[k], s9[k], equal_nan=True), k


def test_reference_runs_the_synthetic_bench_basin_and_the_oracle_matches_bitwise(built):
    """The chain behind tests/golden/fullsize_1m.npz at a size the CPU suite affords: build_basin()
    writes its basin as MeshBuilder files, the reference runs them, its flattened inputs equal
    build_basin()'s dict up to the sparse patches, and the oracle on those inputs reproduces the
    reference's arrays bit for bit (storm end included)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_fullsize as mf
    out = os.path.join(tempfile.mkdtemp(prefix="tribs_fs_"), "fs.npz")
    mf.main(side=40, out=out)                      # asserts oracle == reference on the full arrays
    z = np.load(out)
    assert int(z["oracle_bitexact"][0]) == 1 and int(z["n_active"][0]) == 1600
    patched = {k.split("/", 2)[2] for k in z.files if k.startswith("patch/")}
    assert patched <= {"varea", "stream_node", "BasArea", "res_limit"}, patched
    assert z["s128/res/mhydro"].max() > 0 and (z["s128/NfOld"] > 0).any()


def test_bench_cpu_legs_run_the_reference_on_the_bench_generator(built):
    """bench.py's two CPU legs (cpu_baseline of the GPU arm, and --impl reference) at a size the CPU
    suite affords: both must come from the unmodified reference reading MeshBuilder decks."""
    import json
    import bench
    cb = bench.cpu_baseline_sample(side=40)
    assert cb["kind"] == "reference" and cb["cores"] == 1 and cb["value"] > 1e4 and "1600 cells" in cb["sample"]
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "2",
                        "--side", "48"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-1500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["steps"] == 3 and line["cpu_baseline"]["kind"] == "reference"
    assert line["value"] == line["e2e"]["value"] > 1e4 and line["cpu_baseline"]["cores"] >= 1
    assert "4096-cell" in line["cpu_baseline"]["sample"]        # the arm never goes below 64 x 64 cells per process
