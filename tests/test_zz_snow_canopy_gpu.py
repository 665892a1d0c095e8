"""GPU parity of tSnowPack::callSnowPack with a canopy (OPTSNOW 1 + OPTINTERCEPT 1/2), through the
C ABI, against live runs of the reference binary. The reference carries two class members
(`albedo`, `snCanWE`) from cell to cell in sweep order; the device reproduces that with store-free
probe rounds + a last-writer scan + one commit (snow_step_with_canopy, tribs_gpu.cu)."""
import numpy as np
import pytest

from common import GpuStepper, run_and_compare
from refrun import have_harness, run_reference, schedule_for, spec_for

pytestmark = pytest.mark.gpu

REL_TOL = 1e-6
ABS_FLOOR = 1e-6


@pytest.mark.skipif(not have_harness(), reason="oracle/_ref/tribs_ref_harness not built")
@pytest.mark.parametrize("extra", [{}, {"HILLALBOPT": 2}, {"OPTINTERCEPT": 1}, {"OPTINTERCEPT": 0, "HILLALBOPT": 2}])
def test_gpu_snow_pack_with_canopy_matches_live_reference(extra, built):
    """(The last variant has no canopy: HILLALBOPT 2 then mixes land and snow albedo, snCanWE staying 0.)"""
    spec = spec_for("snow", 20, runtime_h=96.0, seed=9, pmean=8.0, extra=extra)
    n_steps = 96 * 16
    basin, snaps, _, _ = run_reference(spec, snap_from=1, snap_to=n_steps, snap_every=16, phases="murb")
    assert (int(basin["et_Ioption"][0]) != 0) == (extra.get("OPTINTERCEPT", 2) != 0)
    for k in [k for k in snaps if k.endswith("_m_Uerror") or k.endswith("_m_CumUerror")]:
        del snaps[k]      # energy-balance closure error: 1e-14 round-off noise, see include/tribs_gpu.h
    g = GpuStepper(basin, schedule_for(basin))
    w = run_and_compare(g, snaps, n_steps, floor=ABS_FLOOR)
    print("\n[snow+canopy] worst:\n" + w.report(8))
    assert w.max() <= REL_TOL, w.report()
    assert ("m", "IntSWE") in w.w and ("m", "CumIntSub") in w.w and ("m", "CumMelt") in w.w
    assert max(np.abs(snaps[k]).max() for k in snaps if k.endswith("_m_CumMelt")) > 0.0
    members, rounds = g.sim.dev.snow_members()
    if extra.get("OPTINTERCEPT", 2) != 0:
        assert max(np.abs(snaps[k]).max() for k in snaps if k.endswith("_m_CumIntSub")) > 0.0
        assert 0.45 <= members[0] <= 0.88 and rounds >= 1
    else:
        assert rounds == 0


def test_gpu_snow_golden_fixture(built):
    """The committed snow fixture (tests/golden/snow.npz: winter met, Rutter canopy, 1 152 steps of the
    reference with melt and rain on snow) through the device path."""
    from refrun import load_golden
    basin, snaps, _ = load_golden("snow")
    snaps = {k: v for k, v in snaps.items() if not (k.endswith("_m_Uerror") or k.endswith("_m_CumUerror"))}
    n_steps = int(snaps["n_steps"][0])
    w = run_and_compare(GpuStepper(basin, schedule_for(basin)), snaps, n_steps, floor=ABS_FLOOR)
    print("\n[snow golden] worst:\n" + w.report(8))
    assert w.max() <= REL_TOL, w.report()
    assert ("m", "IntSWE") in w.w and ("m", "IceWE") in w.w


@pytest.mark.skipif(not have_harness(), reason="oracle/_ref/tribs_ref_harness not built")
@pytest.mark.parametrize("name,option", [("met", 2), ("met", 3), ("snow", 1)])
def test_gpu_radiation_sheltering_matches_live_reference(name, option, built):
    """OPTRADSHELT 1-3: tShelter's sky-view factor (init SheltFact) and horizon angle at azimuth 0 reach
    kernel (1) / the snow kernels through tribs_et_params_t; a DEM with a 300 m wall on its east
    margin makes the factor vary from cell to cell."""
    hours = 40.0
    spec = spec_for(name, 16, runtime_h=hours, seed=5, dem_ridge=300.0, extra={"OPTRADSHELT": option})
    n_steps = int(hours * 16)
    basin, snaps, _, _ = run_reference(spec, snap_from=1, snap_to=n_steps, snap_every=16, phases="murb")
    assert int(basin["et_shelterOption"][0]) == option
    for k in [k for k in snaps if k.endswith("_m_Uerror") or k.endswith("_m_CumUerror")]:
        del snaps[k]
    w = run_and_compare(GpuStepper(basin, schedule_for(basin)), snaps, n_steps, floor=ABS_FLOOR)
    print(f"\n[shelter {option} {name}] worst:\n" + w.report(6))
    assert w.max() <= REL_TOL, w.report()
    assert ("m", "LongRadIn") in w.w and ("m", "ShortRadIn") in w.w


@pytest.mark.skipif(not have_harness(), reason="oracle/_ref/tribs_ref_harness not built")
@pytest.mark.parametrize("name,hours,extra", [("snow", 48.0, dict(pmean=8.0)), ("met", 30.0, {})])
def test_gpu_kernel1_and_snow_on_a_10k_cell_basin_match_live_reference(name, hours, extra, built):
    """Kernel (1) and the snow kernels beyond a few thread blocks: a 100 x 100-cell basin that the
    reference READS (OPTMESHINPUT 9 files from synthbasin, refrun.run_reference_meshb) - 79 blocks per
    launch, and a last-writer scan of the carried snow members that spans several scan tiles."""
    from refrun import run_reference_meshb
    n_steps = int(hours * 16)
    basin, snaps, timing = run_reference_meshb(100, name, hours, seed=9, snap_from=16, snap_to=n_steps, snap_every=64,
                                               phases="mur", **extra)
    assert int(basin["n_active"][0]) == 10000
    for k in [k for k in snaps if k.endswith("_m_Uerror") or k.endswith("_m_CumUerror")]:
        del snaps[k]
    g = GpuStepper(basin, schedule_for(basin))
    w = run_and_compare(g, snaps, n_steps, floor=ABS_FLOOR)
    print(f"\n[{name}, 10k cells] worst:\n" + w.report(6))
    assert w.max() <= REL_TOL, w.report()
    if name == "snow":
        assert max(np.abs(snaps[k]).max() for k in snaps if k.endswith("_m_CumIntSub")) > 0.0
        assert g.sim.dev.snow_members()[1] >= 1
    else:
        assert max(np.abs(snaps[k]).max() for k in snaps if k.endswith("_m_PotEvap")) > 0.0


@pytest.mark.skipif(not have_harness(), reason="oracle/_ref/tribs_ref_harness not built")
def test_reference_simulator_drives_gpu_snow_path(built):
    """The drop-in boundary for OPTSNOW 1: the reference's own objects and time loop
    (tribs_ref_harness --gpu) call TribsGpuBinding::callSnowPack (tribs_b200/host/tribs_ref_binding.h)
    where Simulator::SurfaceHydroProcesses calls tSnowPack::callSnowPack; pack, canopy snow and the
    water balance underneath must end where the pure reference run ends."""
    import os
    import subprocess
    import tempfile
    from refrun import HARNESS, ROOT
    from tribs_b200.synth import write_basin
    from tribs_b200.trb import read_trb
    d = tempfile.mkdtemp(prefix="tribshyb_snow_")
    write_basin(spec_for("snow", 16, runtime_h=72.0, seed=4, pmean=8.0), d)
    lib = os.path.join(ROOT, "tribs_b200", "libtribs_gpu.so")
    for extra in ([], ["--gpu", lib]):
        r = subprocess.run([HARNESS, "basin.in", "-K", "--out", d, "--no-basin", *extra], cwd=d, capture_output=True,
                           text=True, timeout=1200)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    ref, got = read_trb(os.path.join(d, "qout_ref.trb")), read_trb(os.path.join(d, "qout_gpu.trb"))
    assert ref["IceWE"].max() > 0 and np.abs(ref["CumIntSub"]).max() > 0
    for f, floor in (("NwtNew", 1e-6), ("MuNew", 1e-6), ("NfNew", 1e-6), ("IceWE", 1e-6), ("LiqWE", 1e-6), ("IntSWE", 1e-6),
                     ("SnTempC", 1.0), ("CumMelt", 1e-6), ("CumIntSub", 1e-6), ("PeakSWE", 1e-6), ("SurfTemp", 1.0)):
        err = np.abs(ref[f] - got[f]) / np.maximum(np.abs(ref[f]), floor)
        assert err.max() <= REL_TOL, (f, float(err.max()))


def test_gpu_snow_state_round_trips_through_sync_and_push(built):
    """What a restart needs from the snow path: every snow field and the two carried members can be
    read (tribs_gpu_snow_sync / _members), written back (tribs_gpu_snow_push / _set_members) into a
    second instance, and both then advance identically."""
    from refrun import load_golden
    from tribs_b200.gpu import ET_FIELDS, SNOW_FIELDS
    basin, _, _ = load_golden("snow")
    a, b = GpuStepper(basin, schedule_for(basin)), GpuStepper(basin, schedule_for(basin))

    def step(g, k):
        for _ in range(k):
            g.advance(); g.unsat()
            if g.gw_due():
                g.sat()
            g.route(); g.reset()

    step(a, 160)
    step(b, 160)
    sn = a.sim.dev.snow_sync(SNOW_FIELDS)
    members, _ = a.sim.dev.snow_members()
    assert np.abs(sn["IntSWE"]).max() > 0
    b.sim.dev.snow_push({k: np.zeros_like(v) for k, v in sn.items()}, members=[0.5, 0.0])     # scramble b ...
    assert np.array_equal(b.sim.dev.snow_sync(["IntSWE"])["IntSWE"], np.zeros_like(sn["IntSWE"]))
    b.sim.dev.snow_push(sn, members=members)                                                   # ... and restore it
    back = b.sim.dev.snow_sync(SNOW_FIELDS)
    for k in SNOW_FIELDS:
        assert np.array_equal(back[k], sn[k]), k
    assert np.array_equal(b.sim.dev.snow_members()[0], members)
    step(a, 48)
    step(b, 48)
    fa, fb = a.get(SNOW_FIELDS + ["LiqWE", "IceWE", "NwtNew"]), b.get(SNOW_FIELDS + ["LiqWE", "IceWE", "NwtNew"])
    for k in fa:
        assert np.array_equal(fa[k], fb[k]), k
