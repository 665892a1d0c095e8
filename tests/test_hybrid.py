"""GPU test of the drop-in boundary: the reference's own objects and time loop
(oracle/_ref/tribs_ref_harness --gpu) drive the B200 path through the binding in
tribs_b200/host/tribs_ref_binding.h; tKinemat's channel routing, fed from the synced tCNode
fields, must give the same outlet hydrograph as the pure reference run (north_star: outlet
hydrograph within 1e-6 relative), and the final per-cell Nwt / Mu / Nf must agree too."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from refrun import HARNESS, ROOT, have_harness, spec_for
from tribs_b200.synth import write_basin
from tribs_b200.trb import read_trb

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not have_harness(), reason="oracle/_ref/tribs_ref_harness not built")
@pytest.mark.parametrize("name,n,hours", [("deep", 24, 30.0), ("shallow", 20, 30.0), ("met", 20, 30.0)])
def test_reference_simulator_drives_gpu_path(name, n, hours, built):
    d = tempfile.mkdtemp(prefix="tribshyb_")
    over = dict(gw_depth_mm=300.0, pmean=30.0) if name == "met" else {}   # wet enough for outlet discharge
    write_basin(spec_for(name, n, runtime_h=hours, seed=2, **over), d)
    lib = os.path.join(ROOT, "tribs_b200", "libtribs_gpu.so")
    for extra in ([], ["--gpu", lib]):
        r = subprocess.run([HARNESS, "basin.in", "-K", "--out", d, "--no-basin", *extra], cwd=d, capture_output=True,
                           text=True, timeout=1200)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    ref, got = read_trb(os.path.join(d, "qout_ref.trb")), read_trb(os.path.join(d, "qout_gpu.trb"))
    q0, q1 = ref["qout"], got["qout"]
    assert q0.shape == q1.shape and q0.max() > 0.0
    scale = max(q0.max(), 1e-12)
    assert np.abs(q0 - q1).max() <= 1e-6 * scale, ("outlet hydrograph", np.abs(q0 - q1).max(), scale)
    for f in ("NwtNew", "MuNew", "NfNew"):
        err = np.abs(ref[f] - got[f]) / np.maximum(np.abs(ref[f]), 1e-6)
        assert err.max() <= 1e-6, (f, err.max())
    if name == "met":   # kernel (1) driven by the reference's tEvapoTrans / tIntercept objects
        assert ref["CumTotEvap"].max() > 0.0
        for f, floor in (("PotEvap", 1e-6), ("SurfTemp", 1.0), ("CanStorage", 1e-6), ("CumTotEvap", 1e-6), ("CumIntercept", 1e-6)):
            err = np.abs(ref[f] - got[f]) / np.maximum(np.abs(ref[f]), floor)
            assert err.max() <= 1e-6, (f, err.max())


@pytest.mark.skipif(not have_harness(), reason="oracle/_ref/tribs_ref_harness not built")
@pytest.mark.parametrize("ranks", [2, 3])
def test_reference_host_runs_partitioned_through_the_binding(ranks, built):
    """The reference's C++ objects drive a PARTITIONED run (what its MPI build does with tGraph + tParallel): the
    harness partitions its own reach lists with tGraph::createDefaultPartition, the binding creates one handle per
    partition (tribs_part_t) and runs 16-step pipelined batches on all of them (tribs_gpu_run, peer-memory halos);
    the tCNodes, refilled from the ranks that own them, must match the pure serial reference (its serial == MPI
    contract, test_big_spring.py:5-22) within 1e-6."""
    d = tempfile.mkdtemp(prefix="tribshybp_")
    write_basin(spec_for("deep", 40, runtime_h=8.0, seed=4, tributary_every=8, pmean=30.0), d)
    lib = os.path.join(ROOT, "tribs_b200", "libtribs_gpu.so")
    common = [HARNESS, "basin.in", "-K", "--no-basin", "--snap-from", "16", "--snap-to", "96", "--snap-every", "16",
              "--phases", "e", "--max-steps", "96"]
    outs = []
    for tag, extra in (("ref", []), ("gpu", ["--gpu", lib, "--gpu-ranks", str(ranks), "--gpu-batch", "16"])):
        out = os.path.join(d, tag)
        os.makedirs(out)
        r = subprocess.run([*common, "--out", out, *extra], cwd=d, capture_output=True, text=True, timeout=1200)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        outs.append(read_trb(os.path.join(out, "steps.trb")))
    ref, got = outs
    fields = ["NwtNew", "MuNew", "MiNew", "NtNew", "NfNew", "RuNew", "RiNew", "Qpout", "intstorm", "srf_hr",
              "SoilMoisture", "RootMoisture"]
    worst = 0.0
    for st in (16, 48, 96):
        for f in fields:
            a, b = ref[f"s{st}_e_{f}"], got[f"s{st}_e_{f}"]
            worst = max(worst, float((np.abs(a - b) / np.maximum(np.abs(a), 1e-6)).max()))
    assert worst <= 1e-6, worst
    moved = {f: float(np.abs(ref[f"s96_e_{f}"] - ref[f"s16_e_{f}"]).max()) for f in ("NwtNew", "MuNew", "NfNew", "srf_hr")}
    assert moved["MuNew"] > 0 and moved["NfNew"] > 0, moved


@pytest.mark.skipif(not have_harness(), reason="oracle/_ref/tribs_ref_harness not built")
def test_reference_host_routes_its_channels_on_the_device(built):
    """SURVEY §8 row f2 through the boundary: the reference's own tKinemat hands its reach lists, widths, slopes and levels
    to tribs_gpu_channel_init (TribsGpuBinding::attachChannels), every 16-step batch is followed by ONE
    tribs_gpu_channel_route call, and Hlevel / Qstrm of the tCNodes and the outlet discharge of every step must match the
    pure serial reference (its own RunRoutingModel on the host) within 1e-6. Dendritic net, two reach partitions."""
    d = tempfile.mkdtemp(prefix="tribshybc_")
    write_basin(spec_for("deep", 40, runtime_h=8.0, seed=4, tributary_every=8, pmean=30.0), d)
    lib = os.path.join(ROOT, "tribs_b200", "libtribs_gpu.so")
    common = [HARNESS, "basin.in", "-K", "--no-basin", "--snap-from", "16", "--snap-to", "96", "--snap-every", "16",
              "--phases", "e", "--max-steps", "96"]
    outs = {}
    for tag, extra in (("ref", []), ("gpu", ["--gpu", lib, "--gpu-ranks", "2", "--gpu-batch", "16", "--gpu-channels"])):
        out = os.path.join(d, tag)
        os.makedirs(out)
        r = subprocess.run([*common, "--out", out, *extra], cwd=d, capture_output=True, text=True, timeout=1200)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        outs[tag] = read_trb(os.path.join(out, "steps.trb"))
    q_ref = read_trb(os.path.join(d, "ref", "qout_ref.trb"))["qout"][:96]
    q_gpu = outs["gpu"]["qout"]
    assert q_gpu.size == 96 and q_ref.max() > 0.05, (q_gpu.size, q_ref.max())
    worst = float((np.abs(q_ref - q_gpu) / np.maximum(np.abs(q_ref), 1e-6)).max())
    for st in (16, 48, 96):
        for f in ("Hlevel", "Qstrm"):
            a, b = outs["ref"][f"s{st}_e_{f}"], outs["gpu"][f"s{st}_e_{f}"]
            worst = max(worst, float((np.abs(a - b) / np.maximum(np.abs(a), 1e-6)).max()))
    assert worst <= 1e-6, worst
    assert outs["ref"]["s96_e_Hlevel"].max() > 0.01
