"""Developer aid: hop latency of the dependency chain from a TRIBS_TRACE of one K-step launch on the 1M-cell valley.
    python tools/hop_trace.py [side] [K]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
side = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
path = "gpurun_out/trace.bin"
os.makedirs("gpurun_out", exist_ok=True)
from tribs_b200.model import Simulator  # noqa: E402
from tribs_b200.synthbasin import build_basin  # noqa: E402

basin = build_basin(side, side, seed=1, runtime_h=240.0)
sim = Simulator(basin, rain_schedule=lambda t: 20.0 if (t % 24.0) <= 6.0 else 0.0)
sim.run_batch(5)
os.environ["TRIBS_TRACE"] = path
sim.run_batch(K)
sim.dev.wait()
del os.environ["TRIBS_TRACE"]
raw = np.fromfile(path, dtype=np.int32, count=2)
P, T = int(raw[0]), int(raw[1])
hdr = np.fromfile(path, dtype=np.int32, count=2 + P + 2 * T)
kind, hi, first = hdr[2:2 + P], hdr[2 + P:2 + P + T], hdr[2 + P + T:]
tr = np.fromfile(path, dtype=np.uint64, offset=4 * (2 + P + 2 * T)).reshape(P, T, 2).astype(np.int64)
t0 = tr[..., 0][tr[..., 0] > 0].min()
st, en = (tr[..., 0] - t0) * 1e-3, (tr[..., 1] - t0) * 1e-3      # us
print(f"P={P} tiles={T} launch {en.max() / 1e3:.1f} ms")
bnd = np.asarray(basin["boundary"])
river_tiles = np.flatnonzero((first >= 0) & (bnd[np.maximum(first, 0)] == 3))      # one river cell per tile, in sweep order = downstream
river_tiles = river_tiles[np.argsort(first[river_tiles])]
for p in (0, P // 2, P - 1):
    if kind[p] != 0:
        continue
    s_, e_ = st[p, river_tiles], en[p, river_tiles]
    gap = np.diff(s_)
    dur = e_ - s_
    print(f"phase {p}: river chain {len(river_tiles)} cells: start-to-start hop mean {gap.mean():.1f} us (median {np.median(gap):.1f}), "
          f"task duration mean {dur.mean():.1f} us; first starts at {s_[0] / 1e3:.1f} ms, last ends at {e_[-1] / 1e3:.1f} ms")
# hillslope chain: tiles by depth for one segment: start time of the first task of each depth in phase 0
fd = np.asarray(basin["flow_dest"])
cur = np.where(bnd == 3, 0, -1)
for _ in range(4000):
    todo = cur < 0
    if not todo.any():
        break
    d = cur[fd[todo]]
    cur[np.nonzero(todo)[0][d >= 0]] = d[d >= 0] + 1
tdepth = np.where(first >= 0, cur[np.maximum(first, 0)], -1)
p = 0
starts = np.array([st[p, tdepth == d].min() for d in range(tdepth.max(), 0, -1)])
print(f"phase 0 hillslope front: depth {tdepth.max()} -> 1 in {(starts[-1] - starts[0]) / 1e3:.1f} ms = {np.diff(starts).mean():.1f} us per depth level; "
      f"lo task duration mean {(en[p] - st[p])[(hi == 0) & (first >= 0)].mean():.1f} us")
