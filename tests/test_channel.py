"""The kinematic-wave channel router (tKinemat::RunRoutingModel, src/tFlowNet/tKinemat.cpp:803-897).

CPU: the C restatement (oracle/tribs_oracle_channel.c), fed by the oracle's own hillslope routing, against what the
unmodified reference left after EVERY step (tests/golden/channel_*.npz, tests/golden/make_channel_golden.py) - bit for
bit: outlet discharge, level / discharge / velocity of every stream node, the per-reach outlet levels."""
import numpy as np
import pytest

from common import OracleStepper
from refrun import load_golden, schedule_for


def channel_golden(name):
    basin, _, _ = load_golden(name)
    import os
    from refrun import ROOT
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    return basin, {k[2:]: z[k] for k in z.files if k.startswith("c/")}


def oracle_step(o, ch):
    """one reference time step up to and including tKinemat::SurfaceFlow; returns Qout"""
    o.advance(); o.unsat()
    if o.gw_due():
        o.sat()
    o.m.route()
    qout = ch.step(o.m.retrieve_qeff())
    o.balance(); o.reset()
    return qout


@pytest.mark.parametrize("name", ["channel_single", "channel_dendritic"])
def test_oracle_channel_router_is_bit_exact(name):
    from oracle.pyoracle import OracleChannel
    basin, ref = channel_golden(name)
    o, ch = OracleStepper(basin, schedule_for(basin)), OracleChannel(basin)
    stream, n = ref["stream_cells"], int(basin["n_active"][0])
    assert ref["qout"].max() > 0.5 and (ref["Hlevel"][-1] > 0).any()          # water did reach the channels
    for s in range(int(ref["n_steps"][0])):
        qout = oracle_step(o, ch)
        assert qout == ref["qout"][s], (s, qout, ref["qout"][s])
        assert np.array_equal(ch.hlevel[stream], ref["Hlevel"][s]), s
        assert np.array_equal(ch.qstrm[stream], ref["Qstrm"][s]), s
        assert np.array_equal(ch.velocity[stream], ref["FlowVelocity"][s]), s
        assert np.array_equal(ch.outlet_hlev, ref["OutletHlev"][s]), s
        assert ch.hlevel[n] == ref["outlet_Hlevel"][s] and ch.qstrm[n] == ref["outlet_Qstrm"][s], s
    assert ch.iters > 0


def test_partition_solver_matches_the_reference_order_solver(built):
    """channel_solve.cuh compiled with g++: the partition form of the Newton step's linear solve (chunks eliminated side by
    side, a small interface system, assembly) against the reference-order pivoted band solve and against LAPACK, on
    Jacobians of the kinematic-wave system (diagonal ~ b/dt, off-diagonals +-K h^(2/3): pivoting in almost every row) and
    on random tridiagonal systems, sizes from 2 rows to the 10M-cell basin's reach."""
    import ctypes as C
    import os
    from refrun import ROOT
    L = C.CDLL(os.path.join(ROOT, "tests", "hostcheck", "libhostcheck.so"))
    PD = C.POINTER(C.c_double)
    L.hc_chan_solve.argtypes = [C.c_int, PD, PD, PD, PD, PD, PD, C.POINTER(C.c_int)]
    rng = np.random.default_rng(7)
    worst, seen_parts = 0.0, set()
    for n in (2, 3, 11, 12, 23, 24, 25, 100, 383, 384, 385, 1000, 3163):
        for kind in ("jacobian", "random", "dominant"):
            if kind == "jacobian":
                h = rng.uniform(0.0, 3.0, n + 2) ** (2.0 / 3.0)
                K = rng.uniform(0.5, 30.0)
                dia = rng.uniform(0.5, 1.5, n)
                sup = (5.0 / 6.0) * K * h[2:] + 0.2
                sub = -(5.0 / 6.0) * K * h[:n] + 0.2
            elif kind == "random":
                dia, sup, sub = (rng.normal(size=n) for _ in range(3))
                dia = dia + np.sign(dia) * 0.3
            else:
                sup, sub = rng.normal(size=n), rng.normal(size=n)
                dia = 4.0 + np.abs(sup) + np.abs(sub)
            f = rng.normal(size=n)
            from scipy.linalg import solve_banded
            ab = np.zeros((3, n))
            ab[0, 1:], ab[1], ab[2, :-1] = sup[:-1], dia, sub[1:]
            x = solve_banded((1, 1), ab, f)                  # LAPACK dgbsv: partial pivoting
            resid = np.abs(dia * x + np.r_[sup[:-1] * x[1:], 0.0] + np.r_[0.0, sub[1:] * x[:-1]] - f).max()
            if resid > 1e-9 * max(1.0, np.abs(x).max()):     # too ill-conditioned to serve as a reference
                continue
            xr, xp, parts = np.empty(n), np.empty(n), C.c_int(0)
            arrs = [np.ascontiguousarray(a, np.float64) for a in (sub, dia, sup, f)]
            L.hc_chan_solve(n, *[a.ctypes.data_as(PD) for a in arrs], xr.ctypes.data_as(PD), xp.ctypes.data_as(PD), C.byref(parts))
            scale = np.abs(x).max()
            e_ref, e_part = np.abs(xr - x).max() / scale, np.abs(xp - x).max() / scale
            tol = 1e-9 if kind != "random" else 1e-6        # random systems can be badly conditioned
            assert e_ref <= tol and e_part <= tol, (n, kind, parts.value, e_ref, e_part)
            assert 1 <= parts.value <= 32 and (parts.value == 1 or n // parts.value >= 12)
            seen_parts.add(parts.value)
            worst = max(worst, e_part)
    assert worst > 0.0 and max(seen_parts) >= 13 and 1 in seen_parts and 2 in seen_parts


def test_oracle_channel_router_uniform_width_against_a_live_reference_run():
    """CHANNELWIDTHCOEFF <= 0 makes tKinemat use CHANNELWIDTH for every node (tKinemat.cpp:44-47, 1227-1228): the branch the
    golden basins (geomorphic widths) do not take. tInputFile::ReadItem matches keywords by prefix
    (tInputFile.cpp:85), so CHANNELWIDTH has to come first in the deck for the reference to read it at all. Live run of
    the reference, every step, bit for bit."""
    import os
    import re
    import subprocess
    import tempfile
    from refrun import HARNESS, have_harness, spec_for
    if not have_harness():
        pytest.skip("oracle/_ref/tribs_ref_harness not built")
    from oracle.pyoracle import OracleChannel
    from tribs_b200.synth import write_basin
    from tribs_b200.trb import read_trb
    steps = 240
    spec = spec_for("shallow", 14, runtime_h=steps * 0.0625 + 1, tributary_every=5, seed=3,
                    extra={"CHANNELWIDTHCOEFF": 0.0, "CHANNELWIDTH": 4.0})
    d = tempfile.mkdtemp(prefix="tribs_chanw_")
    write_basin(spec, d)
    inp = os.path.join(d, "basin.in")
    txt = open(inp).read()
    blocks = {k: re.search(rf"^{k}:[^\n]*\n[^\n]*\n", txt, re.M).group(0) for k in ("CHANNELWIDTHCOEFF", "CHANNELWIDTH")}
    txt = txt.replace(blocks["CHANNELWIDTH"], "").replace(blocks["CHANNELWIDTHCOEFF"], blocks["CHANNELWIDTH"] + blocks["CHANNELWIDTHCOEFF"])
    open(inp, "w").write(txt)
    r = subprocess.run([HARNESS, "basin.in", "-K", "--out", d, "--snap-from", "1", "--snap-to", str(steps), "--snap-every", "1",
                        "--phases", "r"], cwd=d, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    basin, snaps = read_trb(os.path.join(d, "basin.trb")), read_trb(os.path.join(d, "steps.trb"))
    basin["_spec_pmean"], basin["_spec_stdur"] = np.array([spec.pmean]), np.array([spec.stdur])
    basin["_spec_istdur"], basin["_spec_evap"] = np.array([spec.istdur]), np.array([spec.inject_evap])
    assert float(basin["chan_width_uniform"][0]) == 4.0 and len(basin["reach_head"]) > 1
    o, ch = OracleStepper(basin, schedule_for(basin)), OracleChannel(basin)
    stream = np.nonzero(np.asarray(basin["boundary"]) == 3)[0]
    peak = 0.0
    for s in range(1, steps + 1):
        qout = oracle_step(o, ch)
        assert qout == snaps[f"s{s}_r_globals"][6], s
        assert np.array_equal(ch.hlevel[stream], snaps[f"s{s}_r_Hlevel"][stream]), s
        assert np.array_equal(ch.qstrm[stream], snaps[f"s{s}_r_Qstrm"][stream]), s
        assert np.array_equal(ch.outlet_hlev, snaps[f"s{s}_r_OutletHlev"]), s
        peak = max(peak, qout)
    assert peak > 0.1


def test_synthetic_basin_carries_the_reference_channel_network_and_aspect():
    """synthbasin.build_basin's channel keys (one reach from the head of the river to the outlet, geomorphic widths) and
    cell aspects against what the reference itself sets up when it reads the same basin (OPTMESHINPUT 9): the inputs of
    tribs_gpu_channel_init and of kernel (1) on the bench basins."""
    from refrun import have_harness, run_reference_meshb
    if not have_harness():
        pytest.skip("oracle/_ref/tribs_ref_harness not built")
    from tribs_b200.channel import channel_network
    from tribs_b200.synthbasin import build_basin
    ref, _, _ = run_reference_meshb(60, "metday", 1.0)
    mine = build_basin(60, 60, seed=1, runtime_h=1.0, gw_depth_mm=1500.0)
    for k in ("reach_head", "reach_outlet", "reach_nnodes", "outlet_chan_width", "chan_roughness", "chan_width_uniform",
              "chan_percolation", "chan_optres"):
        assert np.array_equal(np.asarray(mine[k]), np.asarray(ref[k])), k
    stream = np.asarray(ref["boundary"]) == 3
    assert np.array_equal(np.asarray(mine["chan_width"])[stream], np.asarray(ref["chan_width"])[stream])
    a, b = channel_network(mine), channel_network(ref)
    for k in ("off", "node", "width", "length", "slope"):
        assert np.array_equal(a[k], b[k]), k
    assert np.abs(np.asarray(mine["aspect"]) - np.asarray(ref["aspect"])).max() <= 4e-15      # numpy's arccos against libm's
