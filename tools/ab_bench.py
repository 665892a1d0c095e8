"""Quick A/B timing of compile-time variants of libtribs_gpu.so (no torch import, seconds per run).

  TRIBS_GPU_LIB=tribs_b200/libtribs_gpu_cm.so python tools/ab_bench.py --side 1000

Same basin, storm and batching as bench.py's resident leg (pipelined batches of 64 steps); the
time is the library's own CUDA-event timing of the batch kernel (tribs_gpu_phase_ms) next to the
wall clock around run + wait. Prints one JSON line. Not a bench value: bench.py is the contract."""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=64)
    ap.add_argument("--batch", type=int, default=64)
    a = ap.parse_args()
    from bench import storm_rain as bench_rain
    from tribs_b200 import gpu
    from tribs_b200.model import Simulator
    from tribs_b200.synthbasin import build_basin
    t0 = time.perf_counter()
    basin = build_basin(a.side, a.side, seed=1, runtime_h=240.0, stream_every=0)
    sim = Simulator(basin, rain_schedule=None)
    dev, t = sim.dev, sim.timer
    setup_s = time.perf_counter() - t0

    def host_steps(k):
        out = []
        for _ in range(k):
            sim.advance()
            now = t.getCurrentTime()
            out.append(dict(current_time=now, elapsed_steps=t.getElapsedSteps(now),
                            hourly_reset=int(not math.fmod(now - t.getTimeStep(), t.etistep)),
                            sat=bool(sim.gw_on and not math.fmod(now, t.getGWTimeStep())), rain=bench_rain(now)))
        return out

    def loop(k):
        done = 0
        while done < k:
            kk = min(a.batch, k - done)
            dev.run(host_steps(kk))
            done += kk

    loop(a.warmup)
    dev.wait()
    dev.set_timing(True)
    dev.phase_ms(reset=True)
    w0 = time.perf_counter()
    loop(a.steps)
    dev.wait()
    wall_ms = (time.perf_counter() - w0) * 1e3
    ph = dev.phase_ms(reset=True)
    n = int(basin["n_active"][0])
    print(json.dumps({"lib": os.path.basename(gpu.LIB_PATH), "cells": n, "steps": a.steps, "setup_s": round(setup_s, 2),
                      "wall_ms_per_step": wall_ms / a.steps, "phase_ms_per_step": {k: v / a.steps for k, v in ph.items()},
                      "cell_timesteps_per_s_wall": n * a.steps / (wall_ms * 1e-3)}))


if __name__ == "__main__":
    main()
