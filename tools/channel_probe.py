"""Cost of the device channel router (channel_route_kernel) next to the batch it consumes: a V-valley with a shallow
water table under a uniform storm, so that runoff reaches the river within the first hour.

    python tools/channel_probe.py [side=1000] [batches=8] [steps_per_batch=20]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from tribs_b200.model import Simulator  # noqa: E402
from tribs_b200.synthbasin import build_basin  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
batches = int(sys.argv[2]) if len(sys.argv) > 2 else 8
K = int(sys.argv[3]) if len(sys.argv) > 3 else 20
basin = build_basin(side, side, seed=1, runtime_h=48.0, gw_depth_mm=float(os.environ.get("GW_DEPTH_MM", "300")))
sim = Simulator(basin, rain_schedule=lambda t: 20.0 if t <= 6.0 else 0.0)
sim.flow.route_channels(basin)
sim.flow.defer = True                       # time the two parts separately
sim.dev.reserve_batch(K)
print(f"cells {sim.dev.n}, stream nodes {sim.dev.n_stream + 1}, {K} steps per batch")
prev = 0
for b in range(batches):
    t0 = time.perf_counter()
    sim.run_batch(K)
    sim.dev.wait()
    t1 = time.perf_counter()
    sim.flow.pending_steps.clear()
    sim.flow.channels_after_batch(K, deferred=True)
    t2 = time.perf_counter()
    st = sim.dev.channel_sync()
    it = st["newton_iterations"]
    print(f"batch {b}: hillslopes {1e3 * (t1 - t0):8.2f} ms | channels {1e3 * (t2 - t1):7.2f} ms, {it - prev:4d} Newton iterations "
          f"({1e3 * (t2 - t1) / max(it - prev, 1):.3f} ms each), backtracks so far {st['backtracks']}, Qout {sim.flow.Qout:.4f} m3/s, "
          f"max level {st['Hlevel'].max():.3f} m")
    prev = it
print('thread-0 cycles:', st['cycles'])
sim.dev.close()
