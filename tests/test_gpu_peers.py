"""Reach partitions in pipelined batches (tribs_part_t): the partition layout on one GPU, and several
ranks exchanging boundary values through peer memory. The ranks live in one process here (LocalPeers:
several handles on cuda:0, each with its share of the SMs), which exercises the same kernels, remote
stores and system-scope atomics as one process per GPU (bench.py --gpus N checks that form on its own
basin before it times anything). Contract: the reference's serial == MPI (test_big_spring.py:5-22), here
bit for bit - state, hydrograph arrays, basin accumulators and per-step lateral inflows."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

STATE = ["NwtOld", "MuOld", "MiOld", "NtOld", "NfOld", "RuOld", "RiOld", "Qpout", "intstorm", "srf_hr", "cumsrf",
         "SoilMoisture", "RootMoisture", "AvSoilMoisture", "Recharge", "hsrfOccur", "sbsrfOccur", "psrfOccur",
         "satsrfOccur", "satOccur", "RechDisch", "esrf", "Transmissivity", "TTime", "FlowVelocity", "UnSatFlowIn",
         "UnSaturatedStorage", "SaturatedStorage", "gwchange"]
RESULTS = ["mhydro", "HsrfRout", "SbsrfRout", "PsrfRout", "SatsrfRout", "crr", "rmax", "rmin", "frac", "msm", "msmRt",
           "msmU", "mgw", "sat", "qunsat"]


def _basin(nx, ny, **kw):
    from tribs_b200.synthbasin import build_basin
    return build_basin(nx, ny, seed=3, runtime_h=48.0, gw_depth_mm=kw.pop("gw_depth_mm", 900.0), **kw)


def _rain(t):
    return 25.0 if t <= 3.0 else 0.0


def _drive(sim, n_steps, batch):
    q, done = [], 0
    while done < n_steps:
        k = min(batch, n_steps - done)
        sim.run_batch(k, rain_of_time=_rain, collect_qeff=True)
        q.append(sim.batch_qeff)
        done += k
    return np.concatenate(q)


def _one_gpu(basin, owner, n_steps, batch):
    from tribs_b200.model import Simulator
    part = None if owner is None else dict(rank=0, nranks=1, cell_owner=owner)
    sim = Simulator(basin, part=part)
    q = _drive(sim, n_steps, batch)
    lim = int(basin["res_limit"][0])
    out = dict(state=sim.dev.sync(STATE), res={nm: sim.dev.results(nm, lim) for nm in RESULTS}, glob=sim.dev.globals(), q=q)
    sim.dev.close()
    return out


def test_partition_layout_changes_no_cell_value(built):
    """Laying the basin out by partition regroups the basin sums and nothing else."""
    from tribs_b200.partition import cell_owner
    basin = _basin(36, 108)
    a = _one_gpu(basin, None, 96, 24)
    b = _one_gpu(basin, cell_owner(basin, 3), 96, 24)
    for f in STATE:
        assert np.array_equal(a["state"][f], b["state"][f]), f
    # sums regroup with the layout: the hillslope inflows of a stream cell and the basin arrays
    assert np.allclose(a["q"], b["q"], rtol=1e-12, atol=1e-300) and np.abs(a["q"]).max() > 0
    assert np.abs(a["res"]["mhydro"]).max() > 0
    for nm in RESULTS:
        assert np.allclose(a["res"][nm], b["res"][nm], rtol=1e-12, atol=1e-300), nm
    for k in a["glob"]:
        assert np.isclose(a["glob"][k], b["glob"][k], rtol=1e-12, atol=1e-9), k


@pytest.mark.parametrize("world,nx,ny,stream_every", [(2, 36, 72, 0), (3, 36, 108, 0), (4, 30, 120, 0), (4, 64, 64, 16)])
def test_peer_ranks_are_bitwise_equal_to_one_gpu(world, nx, ny, stream_every, built):
    from tribs_b200.partition import cell_owner
    from tribs_b200.peers import LocalPeers
    basin = _basin(nx, ny, stream_every=stream_every)
    owner = cell_owner(basin, world)
    assert np.unique(owner).size == world
    n_steps, batch = 128, 32
    ref = _one_gpu(basin, owner, n_steps, batch)
    lp = LocalPeers(basin, world, owner=owner, max_batch_steps=batch)
    q, done = [], 0
    while done < n_steps:
        lp.run_batch(batch, rain_of_time=_rain)
        q.append([d.retrieve_qeff_steps(batch) for d in lp.devs])
        done += batch
    got = lp.gather(STATE)
    for f in STATE:
        assert np.array_equal(ref["state"][f], got[f]), (f, int(np.argmax(ref["state"][f] != got[f])))
    lim = int(basin["res_limit"][0])
    for r, d in enumerate(lp.devs):      # every rank ends up with the whole hydrograph
        for nm in RESULTS:
            assert np.array_equal(ref["res"][nm], d.results(nm, lim)), (r, nm)
        g = d.globals()
        for k in ("Stok", "TotRain", "TotGWchange", "TotMoist", "psrf_last", "BasinRain", "BasinEvap", "BasinRunoff"):
            assert ref["glob"][k] == g[k], (r, k, ref["glob"][k], g[k])
        assert np.array_equal(ref["q"], np.concatenate([qq[r] for qq in q])), r
    assert np.abs(ref["res"]["mhydro"]).max() > 0 and np.abs(ref["q"]).max() > 0
    lp.close()


def _gauge_map(basin, k=3):
    x, y = np.asarray(basin["x"]), np.asarray(basin["y"])
    gx = np.minimum((k * (x - x.min()) / max(np.ptp(x), 1e-9)).astype(np.int32), k - 1)
    gy = np.minimum((k * (y - y.min()) / max(np.ptp(y), 1e-9)).astype(np.int32), k - 1)
    return (gy * k + gx).astype(np.int32)


def _gauge_values(t, n=9):
    base = 18.0 if 5.0 < t <= 8.0 else 0.0       # a morning storm: the canopy fills, then evaporates in daylight
    return base * (0.5 + np.arange(n) / n)


ET_CHECK = ["PotEvap", "ActEvap", "SurfTemp", "SoilTemp", "CanStorage", "CumIntercept", "EvapWetCanopy", "EvapoTrans",
            "CumTotEvap", "CanopyStorVol", "NetRad", "LFlux"]


def test_config3_forcing_on_peer_ranks_is_bitwise_equal_to_one_gpu(built):
    """configs[2] in small: gauge rain (per-gauge values + the static Thiessen map on the device), a weather
    station, Penman-Monteith + Rutter interception as phases of the batch (tribs_gpu_run_et), three ranks."""
    from tribs_b200 import gpu
    from tribs_b200.model import Simulator
    from tribs_b200.partition import cell_owner
    from tribs_b200.peers import LocalPeers
    from tribs_b200.synthbasin import add_met_forcing
    basin = add_met_forcing(_basin(36, 108), hours=48)
    owner = cell_owner(basin, 3)
    cg = _gauge_map(basin)
    rain = lambda t: gpu.GaugeRain(_gauge_values(t))
    n_steps, batch = 192, 48     # midnight to noon; hourly ET: three ET phases inside every batch
    ref = Simulator(basin, part=dict(rank=0, nranks=1, cell_owner=owner))
    ref.dev.set_rain_gauges(cg)
    for _ in range(n_steps // batch):
        ref.run_batch(batch, rain_of_time=rain)
    lp = LocalPeers(basin, 3, owner=owner, max_batch_steps=batch, cell_gauge=cg)
    for _ in range(n_steps // batch):
        lp.run_batch(batch, rain_of_time=rain)
    want = ref.dev.sync(STATE + ["Rain", "NetPrecipitation", "EvapSoil"])
    want.update(ref.dev.et_sync(ET_CHECK))
    got = lp.gather(STATE + ["Rain", "NetPrecipitation", "EvapSoil"] + ET_CHECK)
    for f in want:
        assert np.array_equal(want[f], got[f]), f
    assert np.abs(want["PotEvap"]).max() > 0 and np.abs(want["CumTotEvap"]).max() > 0 and np.abs(want["cumsrf"]).max() >= 0
    lim = int(basin["res_limit"][0])
    for nm in RESULTS:
        assert np.array_equal(ref.dev.results(nm, lim), lp.devs[1].results(nm, lim)), nm
    # and the per-gauge forcing equals the same rain handed over cell by cell
    pc = Simulator(basin, part=dict(rank=0, nranks=1, cell_owner=owner))
    for _ in range(n_steps // batch):
        pc.run_batch(batch, rain_of_time=lambda t: _gauge_values(t)[cg])
    other = pc.dev.sync(STATE)
    for f in STATE:
        assert np.array_equal(want[f], other[f]), f
    lp.close(); ref.dev.close(); pc.dev.close()


def test_snapshot_rollback_and_checkpoint_file(built, tmp_path):
    """tribs_gpu_snapshot / rollback, and the binary checkpoint loaded into a handle with ANOTHER partition layout
    (per-cell arrays travel in sweep order): the continued runs agree bit for bit on every cell value."""
    from tribs_b200.model import Simulator
    from tribs_b200.partition import cell_owner
    basin = _basin(30, 60)
    a = Simulator(basin)
    a.run_batch(40, rain_of_time=_rain)
    a.dev.snapshot()
    t40 = a.timer.currentTime
    a.dev.save_state(str(tmp_path / "ck.bin"))
    a.run_batch(56, rain_of_time=_rain)
    first = a.dev.sync(STATE)
    res_first = a.dev.results("mhydro", int(basin["res_limit"][0]))
    a.dev.rollback()
    a.timer.currentTime = t40
    a.run_batch(56, rain_of_time=_rain)
    again = a.dev.sync(STATE)
    for f in STATE:
        assert np.array_equal(first[f], again[f]), f
    assert np.array_equal(res_first, a.dev.results("mhydro", int(basin["res_limit"][0])))
    b = Simulator(basin, part=dict(rank=0, nranks=1, cell_owner=cell_owner(basin, 2)))
    b.dev.load_state(str(tmp_path / "ck.bin"))
    b.timer.currentTime = t40
    b.run_batch(56, rain_of_time=_rain)
    other = b.dev.sync(STATE)
    for f in STATE:
        assert np.array_equal(first[f], other[f]), f
    assert np.allclose(res_first, b.dev.results("mhydro", int(basin["res_limit"][0])), rtol=1e-12, atol=1e-300)
    a.dev.close(); b.dev.close()
