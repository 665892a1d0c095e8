import sys, time, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(),'tests'))
import numpy as np
from tribs_b200.model import Simulator
from tribs_b200.synthbasin import build_basin
side=int(sys.argv[1]); B=int(sys.argv[2]); reps=int(sys.argv[3]); mode=sys.argv[4] if len(sys.argv)>4 else 'nodeps'
b=build_basin(side,side,runtime_h=240.0)
if mode=='nodeps':
    b['flow_dest'][:] = -1      # every cell independent: pure throughput of the column update
sim=Simulator(b, rain_schedule=lambda t: 20.0 if t<=6 else 0.0)
dev=sim.dev
dev.set_timing(True)
for r in range(reps):
    dev.wait(); t0=time.perf_counter()
    sim.run_batch(B); dev.wait()
    dt=time.perf_counter()-t0
    ph=dev.phase_ms(reset=True)
    print(f"{mode} side={side} B={B} rep={r} batch_ms={dt*1e3:.2f} ms/step={dt*1e3/B:.3f} Mcs/s={side*side*B/dt/1e6:.1f} phases={ {k: round(v,2) for k,v in ph.items()} }", flush=True)
