"""Device channel router (tribs_gpu_channel_route) against the unmodified reference:
tests/golden/channel_*.npz hold what tKinemat::RunRoutingModel left after EVERY step (outlet discharge, level / discharge /
velocity of every stream node, OutletHlev per reach). The hillslope inflows come from the device's own routing, so the
comparison covers the whole chain rain -> soil column -> hillslope routing -> channels.

Bar: 1e-6 relative (north_star); levels and discharges that are still at zero compare with an absolute floor."""
import numpy as np
import pytest

from common import GpuStepper
from refrun import schedule_for
from test_channel import channel_golden

pytestmark = pytest.mark.gpu

REL_TOL = 1e-6
FLOOR = {"qout": 1e-6, "Hlevel": 1e-6, "Qstrm": 1e-6, "FlowVelocity": 1e-6, "OutletHlev": 1e-6}


def worst_of(ref, got, floor):
    ref, got = np.ravel(np.asarray(ref, float)), np.ravel(np.asarray(got, float))
    return float((np.abs(ref - got) / np.maximum(np.abs(ref), floor)).max())


def compare_state(dev, ref, s, slot_of_stream, worst):
    got = dev.channel_sync()
    for f in ("Hlevel", "Qstrm", "FlowVelocity"):
        worst[f] = max(worst.get(f, 0.0), worst_of(ref[f][s], got[f][slot_of_stream], FLOOR[f]))
    worst["OutletHlev"] = max(worst.get("OutletHlev", 0.0), worst_of(ref["OutletHlev"][s], got["OutletHlev"], FLOOR["OutletHlev"]))
    worst["outlet"] = max(worst.get("outlet", 0.0), worst_of([ref["outlet_Hlevel"][s], ref["outlet_Qstrm"][s]],
                                                             [got["Hlevel"][-1], got["Qstrm"][-1]], 1e-6))
    return got


def slots_of(dev, ref):
    """fixture order of the stream cells (ascending cell index) -> the device's stream slots"""
    pos = {int(c): k for k, c in enumerate(dev.stream_cells())}
    return np.array([pos[int(c)] for c in ref["stream_cells"]])


@pytest.mark.parametrize("name", ["channel_single", "channel_dendritic"])
def test_channel_router_single_steps_match_reference(built, name):
    basin, ref = channel_golden(name)
    g = GpuStepper(basin, schedule_for(basin))
    g.sim.flow.route_channels(basin)
    sl = slots_of(g.sim.dev, ref)
    worst = {}
    for s in range(int(ref["n_steps"][0])):
        g.advance(); g.unsat()
        if g.gw_due():
            g.sat()
        g.route()
        worst["qout"] = max(worst.get("qout", 0.0), worst_of(ref["qout"][s], g.sim.flow.Qout, FLOOR["qout"]))
        if s % 7 == 0 or s == int(ref["n_steps"][0]) - 1:
            got = compare_state(g.sim.dev, ref, s, sl, worst)
        g.reset()
    print(f"\n[{name}] worst relative differences: {worst}; Newton iterations {got['newton_iterations']}, "
          f"backtracks {got['backtracks']}")
    assert max(worst.values()) <= REL_TOL, worst
    assert got["newton_iterations"] > 0 and got["maxits_exceeded"] == 0
    g.sim.dev.close()


@pytest.mark.parametrize("name,ranks", [("channel_single", 1), ("channel_dendritic", 1), ("channel_dendritic", 2)])
def test_channel_router_after_batches_matches_reference(built, name, ranks):
    """One launch routes all steps of a batch from the per-step inflows the batch left on the device; with two reach
    partitions every rank routes the whole net from the combined inflows."""
    from tribs_b200.model import Simulator
    from tribs_b200.peers import LocalPeers
    basin, ref = channel_golden(name)
    n_steps, batch = int(ref["n_steps"][0]), 40
    if ranks == 1:
        sims = [Simulator(basin, rain_schedule=schedule_for(basin))]
        run = lambda k: sims[0].run_batch(k)
    else:
        lp = LocalPeers(basin, ranks, rain_schedule=schedule_for(basin), max_batch_steps=batch)
        sims = lp.sims
        run = lp.run_batch
    for sim in sims:
        sim.flow.route_channels(basin)
    sl = slots_of(sims[0].dev, ref)
    worst, done = {}, 0
    while done < n_steps:
        k = min(batch, n_steps - done)
        run(k)
        done += k
        for sim in sims:
            compare_state(sim.dev, ref, done - 1, sl, worst)
    for sim in sims:
        q = np.array(sim.flow.qout_series)
        assert q.size == n_steps
        worst["qout"] = max(worst.get("qout", 0.0), worst_of(ref["qout"], q, FLOOR["qout"]))
    if ranks > 1:
        assert np.array_equal(np.array(sims[0].flow.qout_series), np.array(sims[1].flow.qout_series))
    print(f"\n[{name}, {ranks} rank(s)] worst relative differences: {worst}")
    assert max(worst.values()) <= REL_TOL, worst
    (lp.close() if ranks > 1 else sims[0].dev.close())


def test_two_node_reaches_and_uniform_width_match_the_oracle(built):
    """tKinemat::SolveForTwoNodeReach (tKinemat.cpp:1736-1765: a reach of a head and its outlet, solved by rtsafe on one
    polynomial) and CHANNELWIDTH as a uniform width (1227-1228). The decks of this repo never produce such reaches, so the
    single reach of the golden basin is cut into a two-node reach, a three-node reach and the rest (confluences are the
    heads of the next reaches), and the device is compared with the C restatement on the same lists."""
    from oracle.pyoracle import OracleChannel
    from common import OracleStepper
    from tribs_b200.channel import channel_network
    from tribs_b200.model import Simulator
    basin, _ = channel_golden("channel_single")
    net = channel_network(basin)
    cuts = [(0, 2), (1, 4), (3, len(net["node"]))]             # entry ranges; a cut repeats the confluence node
    sel = np.concatenate([np.arange(a, b) for a, b in cuts])
    net2 = dict(net, off=np.cumsum([0] + [b - a for a, b in cuts]).astype(np.int32), node=net["node"][sel].copy(),
                width=net["width"][sel].copy(), length=net["length"][sel].copy(), slope=net["slope"][sel].copy(),
                uniform_width=1.5)
    sim = Simulator(basin, rain_schedule=schedule_for(basin))
    sim.flow.route_channels(basin, net=net2)
    o, ch = OracleStepper(basin, schedule_for(basin)), OracleChannel(basin, net=net2)
    q_ref = []
    for _ in range(5):
        sim.run_batch(32)
        for _ in range(32):
            o.advance(); o.unsat()
            if o.gw_due():
                o.sat()
            o.m.route()
            q_ref.append(ch.step(o.m.retrieve_qeff()))
            o.balance(); o.reset()
        got, stream = sim.dev.channel_sync(), sim.dev.stream_cells()
        for f, ref in (("Hlevel", ch.hlevel), ("Qstrm", ch.qstrm), ("FlowVelocity", ch.velocity)):
            assert worst_of(ref[stream], got[f][:-1], 1e-6) <= REL_TOL, f
        assert worst_of(ch.outlet_hlev, got["OutletHlev"], 1e-6) <= REL_TOL
    q_ref = np.array(q_ref)
    assert q_ref.max() > 0.3 and worst_of(q_ref, np.array(sim.flow.qout_series), 1e-6) <= REL_TOL
    sim.dev.close()


def test_channel_router_with_its_vectors_in_global_memory(built, monkeypatch):
    """Reaches longer than about 4 000 nodes do not fit the shared-memory layout of the linear solve; the kernel then works
    on its global workspace. Same results (forced here on the small basin)."""
    monkeypatch.setenv("TRIBS_CHAN_GLOBAL_VECTORS", "1")
    from tribs_b200.model import Simulator
    basin, ref = channel_golden("channel_single")
    sim = Simulator(basin, rain_schedule=schedule_for(basin))
    sim.flow.route_channels(basin)
    sl = slots_of(sim.dev, ref)
    worst = {}
    for done in range(40, 201, 40):
        sim.run_batch(40)
        compare_state(sim.dev, ref, done - 1, sl, worst)
    worst["qout"] = worst_of(ref["qout"][:200], np.array(sim.flow.qout_series), FLOOR["qout"])
    assert max(worst.values()) <= REL_TOL and ref["qout"][:200].max() > 0.5, worst
    sim.dev.close()


def test_channel_init_refuses_what_is_not_built(built):
    from tribs_b200.channel import channel_network
    from tribs_b200.model import Simulator
    basin, _ = channel_golden("channel_single")
    sim = Simulator(basin, rain_schedule=schedule_for(basin))
    with pytest.raises(RuntimeError, match="OPTPERCOLATION"):
        sim.dev.channel_init(channel_network(basin), 225.0, percolation_option=1)
    with pytest.raises(RuntimeError, match="channel_init has not been called"):
        sim.dev.channel_route(0)
    net = channel_network(basin)
    net["node"] = net["node"].copy(); net["node"][1] = 0          # a hillslope cell inside a reach
    with pytest.raises(RuntimeError, match="not a stream cell"):
        sim.dev.channel_init(net, 225.0)
    sim.dev.close()


def test_channel_state_in_snapshot_and_checkpoint(built, tmp_path):
    """tribs_gpu_snapshot / rollback and the checkpoint file carry the channel router's state: a run rolled back and
    repeated, and a run continued in a fresh handle from a file, give the same outlet series bit for bit."""
    from tribs_b200.model import Simulator
    basin, ref = channel_golden("channel_dendritic")

    def fresh():
        sim = Simulator(basin, rain_schedule=schedule_for(basin))
        sim.flow.route_channels(basin)
        return sim

    a = fresh()
    a.run_batch(48)
    t48 = a.timer.currentTime
    a.dev.snapshot(); a.host_snapshot()
    a.dev.save_state(str(tmp_path / "ck.bin"))
    a.run_batch(48)
    first = list(a.flow.qout_series[48:])
    state_first = a.dev.channel_sync()
    a.dev.rollback(); a.host_restore()
    a.run_batch(48)
    assert a.flow.qout_series[96:] == first and max(first) > 0
    b = fresh()
    b.dev.load_state(str(tmp_path / "ck.bin"))
    b.timer.currentTime, b.step_no = t48, 48
    b.run_batch(48)
    assert list(b.flow.qout_series) == first
    state_b = b.dev.channel_sync()
    for f in ("Hlevel", "Qstrm", "FlowVelocity", "OutletHlev"):
        assert np.array_equal(state_first[f], state_b[f]), f
    c = Simulator(basin, rain_schedule=schedule_for(basin))       # no channels on the device: the section is skipped
    c.dev.load_state(str(tmp_path / "ck.bin"))
    for s in (a, b, c):
        s.dev.close()
