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
