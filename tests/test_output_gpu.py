"""The `.HHHH_MMd` spatial output written from DEVICE state (tCOutput::WriteDynamicVars, src/tInOut/tOutput.cpp:1238-1400,
through tribs_b200.model.tCOutput): fields synced from the GPU after the last SurfaceFlow, channel columns (Qstrm, Hlev,
FlwVlc of the stream nodes) from the device's channel router, against the file the unmodified reference writes itself in
the same run. Numbers are compared as printed: within one unit of the last printed digit or 1e-6 relative."""
import glob
import os
import subprocess
import tempfile

import numpy as np
import pytest

from common import GpuStepper
from refrun import HARNESS, have_harness, schedule_for, spec_for
from tribs_b200.output import dynamic_vars_name
from tribs_b200.synth import write_basin
from tribs_b200.trb import read_trb

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_harness(), reason="oracle/_ref/tribs_ref_harness not built")]


@pytest.mark.parametrize("name,hours", [("met", 6.0), ("shallow", 3.0)])
def test_dynamic_vars_file_from_device_state_matches_the_reference(name, hours, built):
    from tribs_b200.model import tCOutput
    d = tempfile.mkdtemp(prefix="tribs_outg_")
    spec = spec_for(name, 12, runtime_h=hours + 2.0, seed=2, opt_spatial=1, extra={"SPOPINTRVL": hours})
    write_basin(spec, d)
    last = int(round(hours * 16))
    r = subprocess.run([HARNESS, "basin.in", "-K", "--out", d], cwd=d, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    basin = read_trb(os.path.join(d, "basin.trb"))
    basin["_spec_pmean"], basin["_spec_stdur"] = np.array([spec.pmean]), np.array([spec.stdur])
    basin["_spec_istdur"], basin["_spec_evap"] = np.array([spec.istdur]), np.array([spec.inject_evap])
    want = os.path.join(d, "out", os.path.basename(dynamic_vars_name("sp", hours)))
    assert want in glob.glob(os.path.join(d, "out", "sp.*d"))
    g = GpuStepper(basin, schedule_for(basin))
    g.sim.flow.route_channels(basin)
    n = int(basin["n_active"][0])
    for s in range(last):
        g.advance(); g.unsat()
        if g.gw_due():
            g.sat()
        g.route()
        if s < last - 1:
            g.reset()
    ch, stream = g.sim.dev.channel_sync(), g.sim.dev.stream_cells()
    vel = g.sim.dev.sync(["FlowVelocity"])["FlowVelocity"]
    cols = {"Qstrm": np.zeros(n), "Hlevel": np.zeros(n), "FlowVelocity": vel}
    for f in cols:
        cols[f][stream] = ch[f][:-1]
    out = tCOutput(g.sim.dev, basin, os.path.join(d, "dev"), g.has_et, False)
    mine = out.WriteDynamicVars(hours, channel=cols)
    a, b = open(want).read().splitlines(), open(mine).read().splitlines()
    assert len(a) == len(b) == n + 1 and a[0] == b[0]
    head = a[0].split(",")
    worst = {}
    for x, y in zip(a[1:], b[1:]):
        for col, (p, q) in enumerate(zip(x.split(","), y.split(","))):
            if p == q:
                continue
            digits = len(p.split(".")[1]) if "." in p else 0
            err = abs(float(p) - float(q))
            tol = max(1.01 * 10.0 ** -digits, 1e-6 * abs(float(p)))
            if err > tol:
                worst[head[col]] = max(worst.get(head[col], 0.0), err)
    assert not worst, worst
    body = np.array([[float(v) for v in row.split(",")] for row in a[1:]])
    if name == "shallow":     # the channels carry water
        assert body[:, head.index("Qstrm")].max() > 0 and body[:, head.index("Hlev")].max() > 0
    else:                     # kernel (1) has run
        assert body[:, head.index("ActEvp")].max() > 0 or body[:, head.index("CanStorage")].max() > 0
    g.sim.dev.close()
