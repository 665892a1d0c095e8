"""GPU parity tests (run on the B200 box through the C ABI).

Tolerance: north_star asks for per-cell Nwt, Nf, Mu and the hydrograph within 1e-6 relative of
the serial reference after the full run. The device libm differs from glibc in the last ulp, so
bit equality with the CPU is not the contract; REL_TOL below is applied to *every* compared
field, with an absolute floor for quantities that are differences of O(1e3) values.
Bit equality *is* required run-to-run on the GPU (test_gpu_run_to_run_bitwise)."""
import numpy as np
import pytest

from common import STATE_FIELDS, GpuStepper, OracleStepper, run_and_compare
from refrun import GPU_SCENARIOS, have_harness, load_golden, run_reference, schedule_for, spec_for, storm_of

pytestmark = pytest.mark.gpu

REL_TOL = 1e-6
ABS_FLOOR = 1e-6   # fields are mm, mm/h, m3/s ...: 1e-6 relative to max(|ref|, 1e-6)


def _sched(basin):
    return schedule_for(basin)


@pytest.mark.parametrize("name", GPU_SCENARIOS)
def test_gpu_matches_reference_golden(name, built):
    basin, snaps, _ = load_golden(name)
    n_steps = int(snaps["n_steps"][0])
    w = run_and_compare(GpuStepper(basin, _sched(basin)), snaps, n_steps, floor=ABS_FLOOR)
    print(f"\n[{name}] worst relative differences vs reference:\n" + w.report(8))
    assert w.max() <= REL_TOL, w.report()


def test_gpu_two_level_segment_reduce_matches_reference_stacks(built, monkeypatch):
    """Large partitions stitch the block-boundary runs of the hillslope routing in two levels
    (route_seg_level2_kernel); forced on here, so that the reference's (TimeInd, Qeff) stacks check that path too."""
    monkeypatch.setenv("TRIBS_SEG_LEVEL2_MIN", "0")
    basin, snaps, _ = load_golden("deep")
    n_steps = int(snaps["n_steps"][0])
    w = run_and_compare(GpuStepper(basin, _sched(basin)), snaps, n_steps, floor=ABS_FLOOR)
    assert ("r", "QeffBins") in w.w and w.max() <= REL_TOL, w.report()


@pytest.mark.skipif(not have_harness(), reason="oracle/_ref/tribs_ref_harness not built")
@pytest.mark.parametrize("name,n,hours", [("deep", 40, 30.0), ("shallow", 32, 30.0), ("redistribute", 32, 40.0), ("evap", 32, 40.0), ("met", 32, 40.0)])
def test_gpu_matches_live_reference(name, n, hours, built):
    """Larger basins: the reference binary itself runs on the box's CPU and is compared with."""
    spec = spec_for(name, n, runtime_h=hours, seed=3)
    n_steps = int(round(hours * 16))
    basin, snaps, timing, _ = run_reference(spec, snap_from=1, snap_to=n_steps, snap_every=16,
                                           phases="murb" if spec.forcing == "met" else "urb")
    w = run_and_compare(GpuStepper(basin, _sched(basin)), snaps, n_steps, floor=ABS_FLOOR)
    print(f"\n[{name} n={n}] cells={basin['n_active'][0]} worst:\n" + w.report(8))
    assert w.max() <= REL_TOL, w.report()


@pytest.mark.skipif(not have_harness(), reason="oracle/_ref/tribs_ref_harness not built")
@pytest.mark.parametrize("variant", ["deardorff", "priestley-taylor", "gray", "gradient-gflux", "lapse-rate",
                                     "no-interception", "no-shelter", "sky-cover-nodata"])
def test_gpu_et_option_variants_match_live_reference(variant, built):
    """Kernel (1) under every option value it accepts (tribs_gpu_et_init refuses the others)."""
    from test_oracle import et_variant_spec
    hours = 30.0
    spec = et_variant_spec(variant, 20, hours, 6)
    n_steps = int(hours * 16)
    basin, snaps, _, _ = run_reference(spec, snap_from=1, snap_to=n_steps, snap_every=8, phases="murb")
    w = run_and_compare(GpuStepper(basin, _sched(basin)), snaps, n_steps, floor=ABS_FLOOR)
    print(f"\n[{variant}] worst:\n" + w.report(6))
    assert w.max() <= REL_TOL, w.report()
    assert any(k[0] == "m" for k in w.w)


@pytest.mark.skipif(not have_harness(), reason="oracle/_ref/tribs_ref_harness not built")
@pytest.mark.parametrize("extra", [{"OPTINTERCEPT": 0}, {"OPTINTERCEPT": 0, "GFLUXOPTION": 1}])
def test_gpu_snow_pack_matches_live_reference(extra, built):
    """tSnowPack::callSnowPack without a canopy on the device (winter run with accumulation, melt and
    rain on snow). The energy-balance closure error (1e-14 round-off noise that the reference also
    drags from cell to cell) is not compared."""
    spec = spec_for("snow", 20, runtime_h=96.0, seed=9, pmean=8.0, extra=extra)
    n_steps = 96 * 16
    basin, snaps, _, _ = run_reference(spec, snap_from=1, snap_to=n_steps, snap_every=16, phases="murb")
    for k in [k for k in snaps if k.endswith("_m_Uerror") or k.endswith("_m_CumUerror")]:
        del snaps[k]
    w = run_and_compare(GpuStepper(basin, _sched(basin)), snaps, n_steps, floor=ABS_FLOOR)
    print("\n[snow] worst:\n" + w.report(8))
    assert w.max() <= REL_TOL, w.report()
    assert ("m", "IceWE") in w.w and ("m", "CumMelt") in w.w
    assert max(np.abs(snaps[k]).max() for k in snaps if k.endswith("_m_CumMelt")) > 0.0


@pytest.mark.skipif(not have_harness(), reason="oracle/_ref/tribs_ref_harness not built")
def test_gpu_snow_pack_inside_batches_is_bitwise_identical_to_single_steps(built):
    """OPTSNOW 1 without a canopy: the snow pack is a phase of the pipelined batch (tribs_gpu_run_snow) and must give
    exactly what tribs_gpu_snow_step between single steps gives - every water-balance, ET and snow field."""
    from tribs_b200.gpu import ET_FIELDS, SNOW_FIELDS
    spec = spec_for("snow", 20, runtime_h=60.0, seed=9, pmean=8.0, extra={"OPTINTERCEPT": 0})
    basin, _, _, _ = run_reference(spec, snap_from=1, snap_to=0)
    n_steps = 40 * 16
    a = GpuStepper(basin, _sched(basin))
    for _ in range(n_steps):
        a.advance(); a.unsat()
        if a.gw_due():
            a.sat()
        a.route(); a.reset()
    b = GpuStepper(basin, _sched(basin))
    done = 0
    while done < n_steps:
        k = min(48, n_steps - done)       # three snow phases inside every batch
        b.sim.run_batch(k)
        done += k
    fa, fb = a.get(ALL_CHECK + ["LiqWE", "IceWE", "LiqRouted"]), b.get(ALL_CHECK + ["LiqWE", "IceWE", "LiqRouted"])
    for f in fa:
        assert np.array_equal(fa[f], fb[f]), f
    ea, eb = a.get(ET_FIELDS), b.get(ET_FIELDS)
    for f in ET_FIELDS:
        assert np.array_equal(ea[f], eb[f]), f
    sa, sb = a.get(SNOW_FIELDS), b.get(SNOW_FIELDS)
    for f in SNOW_FIELDS:
        assert np.array_equal(sa[f], sb[f]), f
    assert np.abs(fa["IceWE"]).max() > 0 and np.abs(sa["CumMelt"]).max() > 0
    for nm in ["mhydro", "crr", "msm", "mgw"]:
        assert np.array_equal(a.result(nm), b.result(nm)), nm


def test_gpu_et_per_cell_met_records_equal_station_records(built):
    """Gridded meteorology (METDATAOPTION 2) reaches kernel (1) as one record per cell: the same
    values expanded per cell must give bit-identical results to the station form."""
    from tribs_b200.gpu import ET_FIELDS
    basin, _, _ = load_golden("met")
    a, b = GpuStepper(basin, _sched(basin)), GpuStepper(basin, _sched(basin))
    h = b.sim.evapotrans.host
    st_idx = h.cell_record.copy()
    station_record = h.record
    h.record = lambda t: {k: np.ascontiguousarray(v[st_idx]) for k, v in station_record(t).items()}
    h.cell_record = np.arange(b.n, dtype=np.int32)
    h.elev, h.lat, h.long, h.gmt_st = h.elev[st_idx], h.lat[st_idx], h.long[st_idx], h.gmt_st[st_idx]
    for g in (a, b):
        for _ in range(120):
            g.advance(); g.unsat()
            if g.gw_due():
                g.sat()
            g.route(); g.reset()
    ea, eb = a.get(ET_FIELDS), b.get(ET_FIELDS)
    for f in ET_FIELDS:
        assert np.array_equal(ea[f], eb[f]), f
    assert np.abs(ea["LFlux"]).max() > 0


def test_gpu_et_init_refuses_uncovered_options(built):
    from tribs_b200 import gpu
    basin, _, _ = load_golden("met")
    for key, val in (("et_shelterOption", 2), ("et_luOption", 1), ("et_option", 4)):
        b = dict(basin)
        b[key] = np.array([val], np.int32)
        dev = gpu.GpuBasin(b)
        with pytest.raises(gpu.TribsGpuError):
            dev.et_init(b)
        dev.close()


def test_gpu_run_to_run_bitwise(built):
    basin, snaps, _ = load_golden("deep")
    outs = []
    for _ in range(2):
        g = GpuStepper(basin, _sched(basin))
        for _ in range(200):
            g.advance(); g.unsat()
            if g.gw_due():
                g.sat()
            g.route(); g.reset()
        outs.append((g.get(["NwtOld", "MuOld", "NfOld", "Qpout", "cumsrf"]), g.result("mhydro"), g.globals()))
    for k in outs[0][0]:
        assert np.array_equal(outs[0][0][k], outs[1][0][k]), k
    assert np.array_equal(outs[0][1], outs[1][1])
    assert outs[0][2] == outs[1][2]


def test_gpu_sync_push_roundtrip_and_old_new_views(built):
    basin, _, _ = load_golden("shallow")
    g = GpuStepper(basin, _sched(basin))
    for _ in range(9):
        g.advance(); g.unsat()
        if g.gw_due():
            g.sat()
        g.route(); g.reset()
    # after Reset() the reference has New == Old
    a = g.get(["NwtNew", "NwtOld", "MuNew", "MuOld"])
    assert np.array_equal(a["NwtNew"], a["NwtOld"]) and np.array_equal(a["MuNew"], a["MuOld"])
    x = np.linspace(1.0, 2.0, g.n)
    g.sim.dev.push({"cumsrf": x})
    assert np.array_equal(g.get(["cumsrf"])["cumsrf"], x)


def test_gpu_water_balance_closes(built):
    """Reference-style invariant (testing/black_box: water-balance closure): rain in = runoff +
    storage change, here per step from the device accumulators at machine precision scale."""
    basin, snaps, _ = load_golden("deep")
    g = GpuStepper(basin, _sched(basin))
    for _ in range(int(snaps["n_steps"][0])):
        g.advance(); g.unsat()
        if g.gw_due():
            g.sat()
        g.route(); g.reset()
    gl = g.globals()
    resid = gl["TotRain"] - gl["Stok"] - gl["TotMoist"] + gl["TotGWchange"]
    # lateral exports to the outlet node and the Newton tolerances leave a small residual
    assert abs(resid) < 0.02 * max(gl["TotRain"], 1.0), (gl, resid)


# ------------------------------------------------------------------ pipelined batches (tribs_gpu_run)
ALL_CHECK = ["NwtOld", "MuOld", "MiOld", "NtOld", "NfOld", "RuOld", "RiOld", "Qpout", "intstorm", "srf_hr", "cumsrf",
             "SoilMoisture", "RootMoisture", "AvSoilMoisture", "Recharge", "hsrfOccur", "sbsrfOccur", "psrfOccur",
             "satsrfOccur", "satOccur", "RechDisch", "esrf", "Transmissivity", "TTime", "FlowVelocity", "UnSatFlowIn",
             "UnSaturatedStorage", "SaturatedStorage"]


@pytest.mark.parametrize("name,batch", [("deep", 8), ("shallow", 16), ("redistribute", 5), ("deep", 64),
                                        ("met", 16), ("met", 40)])
def test_gpu_batch_is_bitwise_identical_to_single_steps(name, batch, built):
    """tribs_gpu_run (space-time pipelined) must reproduce tribs_gpu_step exactly."""
    basin, snaps, _ = load_golden(name)
    n_steps = 320
    a = GpuStepper(basin, _sched(basin))
    qa = []
    for _ in range(n_steps):
        a.advance(); a.unsat()
        if a.gw_due():
            a.sat()
        qa.append(a.route().copy()); a.reset()
    b = GpuStepper(basin, _sched(basin))
    qb = []
    done = 0
    while done < n_steps:
        k = min(batch, n_steps - done)
        b.sim.run_batch(k, collect_qeff=True)
        qb.append(b.sim.batch_qeff)
        done += k
    qb = np.concatenate(qb)
    assert np.array_equal(np.array(qa), qb), "per-step lateral inflow differs"
    fa, fb = a.get(ALL_CHECK), b.get(ALL_CHECK)
    for f in ALL_CHECK:
        assert np.array_equal(fa[f], fb[f]), f
    if name == "met":   # batches are cut at the ET / interception sweeps, which must see the same state
        from tribs_b200.gpu import ET_FIELDS
        ea, eb = a.get(ET_FIELDS), b.get(ET_FIELDS)
        for f in ET_FIELDS:
            assert np.array_equal(ea[f], eb[f]), f
        assert np.abs(ea["PotEvap"]).max() > 0 and np.abs(ea["CanStorage"]).max() >= 0
    for nm in ["mhydro", "HsrfRout", "SbsrfRout", "SatsrfRout", "crr", "msm", "mgw", "qunsat", "rmax", "rmin"]:
        assert np.array_equal(a.result(nm), b.result(nm)), nm
    assert a.globals() == b.globals()


def test_gpu_batch_matches_live_reference_after_full_run(built):
    if not have_harness():
        pytest.skip("oracle/_ref/tribs_ref_harness not built")
    spec = spec_for("deep", 48, runtime_h=36.0, seed=5)
    n_steps = 36 * 16
    basin, snaps, timing, _ = run_reference(spec, snap_from=n_steps, snap_to=n_steps, phases="r")
    g = GpuStepper(basin, _sched(basin))
    done = 0
    while done < n_steps:
        k = min(64, n_steps - done)
        g.sim.run_batch(k)
        done += k
    got = g.get(["NwtOld", "MuOld", "NfOld", "NtOld", "cumsrf", "Qpout"])
    pre = f"s{n_steps}_r_"
    for f, r in (("NwtOld", "NwtNew"), ("MuOld", "MuNew"), ("NfOld", "NfNew"), ("NtOld", "NtNew"), ("cumsrf", "cumsrf"),
                 ("Qpout", "Qpout")):
        ref = snaps[pre + r]
        err = np.abs(ref - got[f]) / np.maximum(np.abs(ref), ABS_FLOOR)
        assert err.max() <= REL_TOL, (f, err.max())
    assert np.abs(snaps[pre + "mhydro"] - g.result("mhydro")).max() <= REL_TOL * max(1e-6, np.abs(snaps[pre + "mhydro"]).max())


# ------------------------------------------------------------------ one reach partition per GPU
@pytest.mark.parametrize("side,steps", [(60, 32), (150, 48)])
def test_ranks_on_separate_gpus_bitwise_equal_to_one(side, steps, built):
    """The reference's own multi-rank contract (serial == MPI, test_big_spring.py:5-22), one process per GPU over
    cudaIpc + NVLink. On a one-GPU box the same protocol is covered by tests/test_gpu_peers.py (ranks as handles of
    one process), so nothing N>1-specific goes untested there."""
    import subprocess
    import sys
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("one GPU on the box: see tests/test_gpu_peers.py for the same protocol with in-process ranks")
    from refrun import ROOT
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", f"{ROOT}/tests/mgpu_check.py", str(side), str(steps)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "BITWISE EQUAL" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
