import sys, time, os, math
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(),'tests'))
import numpy as np, torch
from tribs_b200.model import Simulator
from tribs_b200.synthbasin import build_basin
side=int(sys.argv[1]); B=int(sys.argv[2]); reps=int(sys.argv[3])
b=build_basin(side,side,runtime_h=240.0)
sim=Simulator(b, rain_schedule=lambda t: 20.0 if t<=6 else 0.0)
dev=sim.dev
for r in range(reps):
    dev.wait(); t0=time.perf_counter()
    sim.run_batch(B); dev.wait()
    dt=time.perf_counter()-t0
    print(f"side={side} B={B} rep={r} batch_ms={dt*1e3:.2f} ms/step={dt*1e3/B:.3f} Mcs/s={side*side*B/dt/1e6:.1f}", flush=True)
