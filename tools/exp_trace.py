"""Developer aid (not a test): runs two 64-step batches on the 1M-cell benchmark basin with
TRIBS_TRACE set and summarises the task timeline of the second one."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
side = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
path = "gpurun_out/trace.bin"
os.makedirs("gpurun_out", exist_ok=True)
from tribs_b200.model import Simulator  # noqa: E402
from tribs_b200.synthbasin import build_basin  # noqa: E402

basin = build_basin(side, side, seed=1, runtime_h=240.0)
sim = Simulator(basin, rain_schedule=lambda t: 20.0 if (t % 24.0) <= 6.0 else 0.0)
sim.run_batch(B)
os.environ["TRIBS_TRACE"] = path
sim.run_batch(B)
sim.dev.wait()
del os.environ["TRIBS_TRACE"]
raw = np.fromfile(path, dtype=np.int32, count=2)
P, T = int(raw[0]), int(raw[1])
hdr = np.fromfile(path, dtype=np.int32, count=2 + P + 2 * T)
kind, hi, first = hdr[2:2 + P], hdr[2 + P:2 + P + T], hdr[2 + P + T:]
tr = np.fromfile(path, dtype=np.uint64, offset=4 * (2 + P + 2 * T)).reshape(P, T, 2).astype(np.int64)
t0 = tr[..., 0].min()
st, en = (tr[..., 0] - t0) * 1e-3, (tr[..., 1] - t0) * 1e-3      # us
total = en.max()
dur = en - st
print(f"P={P} tiles={T} total {total/1e3:.1f} ms; task us: lo mean {dur[:, hi == 0].mean():.1f} hi mean {dur[:, hi == 1].mean():.1f}")
# depth of every tile (hops to the stream of its first cell)
fd, bnd = np.asarray(basin["flow_dest"]), np.asarray(basin["boundary"])
depth = np.zeros(len(fd), np.int32)
order = np.argsort(-np.asarray(basin["z"]))   # not needed for correctness: iterate to a fixed point instead
cur = np.where(bnd == 3, 0, -1)
for _ in range(2000):
    todo = cur < 0
    if not todo.any():
        break
    d = cur[fd[todo]]
    cur[np.nonzero(todo)[0][d >= 0]] = d[d >= 0] + 1
tdepth = cur[first]
# concurrency over time
nb = 200
edges = np.linspace(0, total, nb + 1)
busy = np.zeros(nb)
for a, b in ((st.ravel(), en.ravel()),):
    ia = np.clip(np.searchsorted(edges, a) - 1, 0, nb - 1)
    ib = np.clip(np.searchsorted(edges, b) - 1, 0, nb - 1)
    np.add.at(busy, ia, 1); np.add.at(busy, ib, -1)
conc = np.cumsum(busy)
print("running tasks over time (20 slices):", np.round(conc.reshape(20, -1).mean(1)).astype(int).tolist())
# progress of phases: completion time of each phase (max end), of the river tiles, of depth classes
riv = (tdepth == 0)
print("phase end (ms) every 8th:", np.round(en.max(1)[::8] / 1e3, 1).tolist())
print("river-tile phase end (ms) every 8th:", np.round(en[:, riv].max(1)[::8] / 1e3, 1).tolist())
print("phase start (ms) every 8th:", np.round(st.min(1)[::8] / 1e3, 1).tolist())
for dsel in (1, 5, 50, 250, tdepth.max()):
    m = tdepth == dsel
    if m.any():
        print(f"depth {dsel}: tiles {m.sum()} first-phase start {st[0, m].min()/1e3:.1f} ms, last-phase end {en[-1, m].max()/1e3:.1f} ms, "
              f"mean gap between a tile's consecutive phases {np.diff(st[:, m], axis=0).mean():.1f} us")
# along the river: start time of phase 0 and of the last phase per river tile (in river order)
rt = np.nonzero(riv)[0]
rt = rt[np.argsort(first[rt])]
sel = rt[:: max(1, len(rt) // 10)]
print("river tile (every 10%): phase0 start ms", np.round(st[0, sel] / 1e3, 1).tolist())
print("river tile (every 10%): last phase end ms", np.round(en[-1, sel] / 1e3, 1).tolist())
print("river: mean task us", dur[:, riv].mean(), "mean gap between consecutive phases of a river tile (us)", np.diff(st[:, riv], axis=0).mean())

# ---------------------------------------------------------------- critical path walk
import ctypes as C  # noqa: E402
from tribs_b200 import gpu  # noqa: E402
n = sim.dev.n
dix = np.empty(n, np.int32)
gpu._chk(sim.dev.L.tribs_gpu_device_index(sim.dev.h, dix.ctypes.data_as(gpu.PI)))
tile_of = dix // 32
rp, col = np.asarray(basin["spoke_rowptr"]), np.asarray(basin["spoke_nbr"])
src = np.arange(n)
has = fd >= 0
down_pairs = np.unique(np.stack([tile_of[src[has]], tile_of[fd[has]]], 1), axis=0)
down_pairs = down_pairs[down_pairs[:, 0] != down_pairs[:, 1]]
rows = np.repeat(np.arange(n), np.diff(rp))
ok = col >= 0
nbr_pairs = np.unique(np.stack([tile_of[rows[ok]], tile_of[col[ok]]], 1), axis=0)
nbr_pairs = nbr_pairs[nbr_pairs[:, 0] != nbr_pairs[:, 1]]


def csr(pairs, T):
    o = np.argsort(pairs[:, 0], kind="stable")
    p = pairs[o]
    ptr = np.zeros(T + 1, np.int64)
    np.add.at(ptr, p[:, 0] + 1, 1)
    return np.cumsum(ptr), p[:, 1]


down_p, down_c = csr(down_pairs, T)                 # tile -> tiles holding its flow destinations
up_p, up_c = csr(down_pairs[:, ::-1], T)            # tile -> tiles holding its upslope contributors
nbr_p, nbr_c = csr(nbr_pairs, T)


def deps(p, t):
    out = [(p, int(u)) for u in up_c[up_p[t]:up_p[t + 1]]] if kind[p] == 0 else []
    if p > 0:
        out.append((p - 1, t))
        wide = kind[p] == 1 or kind[p - 1] == 1
        lst = nbr_c[nbr_p[t]:nbr_p[t + 1]] if wide else down_c[down_p[t]:down_p[t + 1]]
        out += [(p - 1, int(u)) for u in lst]
    return out


pq, tq = np.unravel_index(np.argmax(en), en.shape)
path = []
while True:
    ds = deps(pq, tq)
    if not ds:
        path.append((pq, tq, st[pq, tq], 0.0, -1))
        break
    ends = [en[a, b] for a, b in ds]
    k = int(np.argmax(ends))
    path.append((pq, tq, st[pq, tq] - ends[k], dur[pq, tq], k))
    pq, tq = ds[k]
path = path[::-1]
print(f"critical path: {len(path)} tasks; sum of task time {sum(x[3] for x in path)/1e3:.1f} ms, sum of hand-off gaps {sum(x[2] for x in path[1:])/1e3:.1f} ms")
pd_ = np.array([(a, tdepth[b], g, d) for a, b, g, d, _ in path])
for lo_, hi_ in ((0, 0), (1, 3), (4, 20), (21, 100), (101, 1000)):
    m = (pd_[:, 1] >= lo_) & (pd_[:, 1] <= hi_)
    print(f"  depth {lo_}-{hi_}: {m.sum()} tasks, task time {pd_[m, 3].sum()/1e3:.1f} ms, gaps {pd_[m, 2].sum()/1e3:.1f} ms, mean gap {pd_[m, 2].mean() if m.any() else 0:.1f} us")
# how the path moves: consecutive (phase, depth) along the last 60 tasks
print("tail of the path (phase, depth, gap us, dur us):", [(int(a), int(b), int(g), int(d)) for a, b, g, d in pd_[-40:]])
seg_len = np.diff(np.nonzero(np.diff(pd_[:, 0]) != 0)[0])
print("tasks between phase changes along the path: mean", seg_len.mean() if len(seg_len) else 0, "max", seg_len.max() if len(seg_len) else 0)
np.savez_compressed("gpurun_out/trace_full.npz", st=st.astype(np.float32), en=en.astype(np.float32), kind=kind, hi=hi, first=first,
                    tdepth=tdepth, down_p=down_p, down_c=down_c, up_p=up_p, up_c=up_c, nbr_p=nbr_p, nbr_c=nbr_c,
                    stream_rank=np.cumsum(bnd == 3)[first] - 1)
