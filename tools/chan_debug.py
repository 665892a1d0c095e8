"""Debug aid: per-step lateral inflows and channel state, one rank against two in-process ranks (dendritic golden basin)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from refrun import schedule_for
from test_channel import channel_golden
from tribs_b200.model import Simulator
from tribs_b200.peers import LocalPeers

basin, ref = channel_golden("channel_dendritic")
one = Simulator(basin, rain_schedule=schedule_for(basin))
one.flow.route_channels(basin)
lp = LocalPeers(basin, 2, rain_schedule=schedule_for(basin), max_batch_steps=40)
for s in lp.sims:
    s.flow.route_channels(basin)
print("stream cells equal:", np.array_equal(one.dev.stream_cells(), lp.devs[0].stream_cells()), np.array_equal(lp.devs[0].stream_cells(), lp.devs[1].stream_cells()))
print("owner of stream cells:", lp.owner[one.dev.stream_cells()])
for b in range(4):
    one.run_batch(40); lp.run_batch(40)
    q1 = one.dev.retrieve_qeff_steps(40)
    q2 = [d.retrieve_qeff_steps(40) for d in lp.devs]
    print("batch", b, "inflow max", q1.max(), "diff one vs rank0", np.abs(q1 - q2[0]).max(), "rank0 vs rank1", np.abs(q2[0] - q2[1]).max())
    if np.abs(q1 - q2[0]).max() > 0:
        bad = np.argwhere(np.abs(q1 - q2[0]) > 0)
        print("  first differing (step, slot):", bad[:5].tolist(), q1[tuple(bad[0])], q2[0][tuple(bad[0])])
    c1, c2 = one.dev.channel_sync(), lp.devs[0].channel_sync()
    print("  Hlevel diff", np.abs(c1["Hlevel"] - c2["Hlevel"]).max(), "qout diff", np.abs(np.array(one.flow.qout_series) - np.array(lp.sims[0].flow.qout_series)).max())
