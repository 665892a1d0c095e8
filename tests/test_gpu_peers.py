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
