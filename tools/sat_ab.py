"""Developer aid: time of the single-step saturated-zone kernels (tribs_gpu_phase_ms) on the 1M-cell valley.
TRIBS_GPU_LIB selects the library variant (TRIBS_NVCC_EXTRA=-DTRIBS_SAT_MIN_BLOCKS=n builds one)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tribs_b200.model import Simulator  # noqa: E402
from tribs_b200.synthbasin import build_basin  # noqa: E402

sim = Simulator(build_basin(1000, 1000, seed=1, runtime_h=48.0), rain_schedule=lambda t: 20.0 if t <= 6.0 else 0.0)
for _ in range(16):
    sim.one_step(fetch_qeff=False)
sim.dev.wait()
sim.dev.set_timing(True)
sim.dev.phase_ms(reset=True)
for _ in range(64):
    sim.one_step(fetch_qeff=False)
sim.dev.wait()
ph = sim.dev.phase_ms(reset=True)
print(os.environ.get("TRIBS_GPU_LIB", "default"), {k: round(v / 8, 4) for k, v in ph.items() if k == "sat"}, "ms per groundwater step;", {k: round(v, 2) for k, v in ph.items()})
