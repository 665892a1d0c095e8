"""Developer aid: 24 single steps (tribs_gpu_step path) of the 1M-cell valley, for ncu captures of the single-step kernels
(PROBE_MET=1: with the weather station, so that kernel (1) runs as a launch of its own)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tribs_b200.model import Simulator  # noqa: E402
from tribs_b200.synthbasin import build_basin  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
basin = build_basin(side, side, seed=1, runtime_h=48.0)
if os.environ.get("PROBE_MET"):       # weather station + ET + Rutter interception: et_cells_kernel at step 16, under rain
    from tribs_b200.synthbasin import add_met_forcing
    add_met_forcing(basin, hours=72)
sim = Simulator(basin, rain_schedule=lambda t: 20.0 if t <= 6.0 else 0.0)
for _ in range(24):
    sim.one_step(fetch_qeff=False)
sim.dev.wait()
print("ok", sim.dev.launch_count())
