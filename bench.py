#!/usr/bin/env python
"""bench.py — Voronoi cell-timesteps/s of the per-timestep water balance on B200.

A "step" is one reference time step of the hot path over the whole synthetic basin:
UnSaturatedZone every step, ResetGW+SaturatedZone every GWSTEP/TIMESTEP-th step (8th at the
standard 3.75 min / 30 min), hillslope routing + basin accumulations and Reset every step.

  value  device-resident throughput (rain already on the device), CUDA-event timed on the
         library's compute stream, max over ranks.
  e2e    the same loop through the C ABI with HOST buffers: every step uploads a per-cell
         rain array from pinned host memory and reads back what the host model consumes each
         step (the stream cells' lateral inflow for the channel router + basin accumulators).
  --impl reference   the unmodified reference (oracle/_ref/tribs_ref_harness) on the host CPU.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "voronoi_cell_timesteps_per_sec"
UNIT = "cell-timesteps/s"
# compulsory bytes per cell and invocation of the dominant kernel (unsat sweep)
ET_BYTES_PER_CELL = 548.0      # 236 B read + 312 B written per cell per fused ET call
SNOW_BYTES_PER_CELL = 920.0    # snow pack per call
UNSAT_BYTES_PER_CELL = 520.0
SAT_BYTES_PER_CELL = 710.0
STEP_BYTES_PER_CELL = 677.0      # whole step, storm-only configuration


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 6 for k in range(4) if r[2 + k].lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def storm_rain(t, pmean=20.0, stdur=6.0, istdur=18.0):
    """STOCHASTICMODE 1 as the reference runs it: one storm, then dry (tSimul.cpp:386-406)."""
    return pmean if t <= stdur else 0.0


# -------------------------------------------------------------------------------- GPU arm
WORKLOAD = {
    2: "synthetic V-valley TIN, uniform storm 20 mm/h, TIMESTEP 3.75 min, GWSTEP 30 min, ET/snow off",
    3: "synthetic V-valley TIN, rain gauges + weather station, Penman-Monteith ET + Rutter interception every "
       "METSTEP 60 min, TIMESTEP 3.75 min, GWSTEP 30 min",
    5: "synthetic V-valley TIN, winter weather station + gauges, snow pack (OPTSNOW 1, no canopy) every METSTEP 60 min, "
       "perched zone over a shallow water table, TIMESTEP 3.75 min, GWSTEP 30 min",
}


def self_check(rank, world, steps=32, side=120):
    """N > 1, before anything is timed: the `world` ranks advance a basin of side x side*world cells together
    (peer-memory halos, pipelined batches) and rank 0 compares every owned cell value, the hydrograph arrays, the
    basin accumulators and the per-step lateral inflows with ONE GPU running the same partition layout.
    Bit equality or the run stops (the reference's serial == MPI contract, test_big_spring.py:5-22)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from tribs_b200.model import Simulator
    from tribs_b200.partition import cell_owner
    from tribs_b200.peers import PeerSimulator
    from tribs_b200.synthbasin import build_basin
    fields = ["NwtOld", "MuOld", "MiOld", "NtOld", "NfOld", "RuOld", "RiOld", "Qpout", "intstorm", "srf_hr", "cumsrf",
              "SoilMoisture", "RootMoisture", "AvSoilMoisture", "Recharge", "RechDisch", "esrf", "Transmissivity", "TTime",
              "UnSaturatedStorage", "SaturatedStorage", "gwchange", "satsrfOccur", "sbsrfOccur"]
    results = ["mhydro", "HsrfRout", "SbsrfRout", "PsrfRout", "SatsrfRout", "crr", "rmax", "rmin", "frac", "msm", "msmRt",
               "msmU", "mgw", "sat", "qunsat"]
    basin = build_basin(side, side * world, seed=2, runtime_h=48.0, gw_depth_mm=900.0)
    owner = cell_owner(basin, world)
    rain = lambda t: 25.0 if t <= 1.0 else 0.0      # storm, then the first dry steps (every column changes state)
    ps = PeerSimulator(basin, rank, world, owner=owner, max_batch_steps=16)
    q = []
    for _ in range(steps // 16):
        ps.run_batch(16, rain_of_time=rain)
        q.append(ps.dev.retrieve_qeff_steps(16))
    got = ps.gather(fields)
    lim = int(basin["res_limit"][0])
    res = {nm: ps.dev.results(nm, lim) for nm in results}
    glob = ps.dev.globals()
    bad = []
    if rank == 0:
        ref = Simulator(basin, part=dict(rank=0, nranks=1, cell_owner=owner))
        qr = []
        for _ in range(steps // 16):
            ref.run_batch(16, rain_of_time=rain, collect_qeff=True)
            qr.append(ref.batch_qeff)
        want = ref.dev.sync(fields)
        bad += [f for f in fields if not np.array_equal(want[f], got[f])]
        bad += [nm for nm in results if not np.array_equal(ref.dev.results(nm, lim), res[nm])]
        g1 = ref.dev.globals()
        bad += [k for k in ("Stok", "TotRain", "TotGWchange", "TotMoist", "psrf_last", "BasinRain", "BasinRunoff") if g1[k] != glob[k]]
        if not np.array_equal(np.concatenate(qr), np.concatenate(q)):
            bad.append("qeff_steps")
        if not (np.abs(want["cumsrf"]).max() > 0 and np.abs(res["mhydro"]).max() > 0):
            bad.append("no runoff produced: the check is vacuous")
        ref.dev.close()
    flag = torch.tensor([float(len(bad))], device="cuda")
    dist.broadcast(flag, src=0)
    ps.dev.close()
    info = {"cells": int(basin["n_active"][0]), "ranks": world, "steps": steps, "fields": len(fields), "result_arrays": len(results),
            "bitwise_equal_to_one_gpu": not bad, "different": bad}
    if rank == 0:
        print("bench.py self-check: " + json.dumps(info), file=sys.stderr, flush=True)
    if flag.item() != 0:
        raise SystemExit("bench.py: the partitioned run differs from the one-GPU run; refusing to time it")
    return info


def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from tribs_b200 import gpu
    from tribs_b200.model import Simulator
    from tribs_b200.synthbasin import build_basin

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; tribs_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    t_start = time.perf_counter()

    def say(what):
        if os.environ.get("TRIBS_BENCH_VERBOSE"):
            print(f"bench.py[rank {rank}] {time.perf_counter() - t_start:7.1f} s  {what}", file=sys.stderr, flush=True)

    cfg = args.config
    side = args.side if args.side > 0 else {2: 1000, 3: 3163, 5: 2500}[cfg]       # cfg 5: 6.25M cells per GPU (50M on 8)
    B = max(1, min(args.batch, args.steps))
    runtime_h = 24.0 + (args.warmup + 2 * args.steps + 64) * 0.0625
    check = None
    partitioned = world > 1 and args.mgpu == "partition"
    strong = cfg == 3       # configs[2]: ONE 10M-cell basin on 1/2/4/8 GPUs; configs[1] grows with the GPUs
    if partitioned and not args.no_self_check:
        check = self_check(rank, world)
    ny = side if (strong or not partitioned) else side * world
    # configs[1] on N GPUs: a valley N times as long (side x side*N cells, side^2 per GPU). Either way the basin is
    # cut into reach partitions along the river and boundary values travel through peer memory inside the batch kernel.
    basin = build_basin(side, ny, seed=1 if partitioned else 1 + rank, runtime_h=runtime_h, stream_every=args.stream_every,
                        **({"gw_depth_mm": 600.0} if cfg == 5 else {}))
    if cfg == 3:
        from tribs_b200.synthbasin import add_met_forcing
        add_met_forcing(basin, hours=int(runtime_h) + 24)     # weather station, land class, Penman-Monteith + Rutter
    elif cfg == 5:
        from tribs_b200.synthbasin import add_met_forcing
        add_met_forcing(basin, hours=int(runtime_h) + 24, i_option=0, winter=True, snow=True)
    if partitioned:
        from tribs_b200.partition import cell_owner
        from tribs_b200.peers import PeerSimulator
        owner = cell_owner(basin, world)
        sim = PeerSimulator(basin, rank, world, owner=owner, max_batch_steps=B)
        n = int((owner == rank).sum())
    else:
        n = int(basin["n_active"][0])
        sim = Simulator(basin, rain_schedule=None)
    dev = sim.dev
    say(f"basin of {int(basin['n_active'][0])} cells on the device")
    gw_every = int(round(dev.gwstep / dev.tstep))
    stream = torch.cuda.ExternalStream(dev.stream_handle())
    n_stream = dev.n_stream
    n_host = dev.n      # cells in the arrays the host exchanges with this rank
    dev.reserve_batch(B)      # the batch buffers exist before anything is timed

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # e2e forcing: the host hands over one value per rain gauge and step (tRainfall's gauge input, RAINSOURCE 3;
    # the Thiessen map cell -> gauge is static and lives on the device), from pinned host memory
    n_gauges = 16
    bx, by = np.asarray(basin["x"]), np.asarray(basin["y"])
    gx = np.minimum((4 * (bx - bx.min()) / max(np.ptp(bx), 1e-9)).astype(np.int32), 3)
    gy = np.minimum((4 * (by - by.min()) / max(np.ptp(by), 1e-9)).astype(np.int32), 3)
    dev.set_rain_gauges((gy * 4 + gx).astype(np.int32), n_gauges)
    pinned_rows = torch.empty((B, n_gauges), dtype=torch.float64).pin_memory()
    gauge_shape = 0.5 + np.arange(n_gauges) / n_gauges if cfg != 2 else np.ones(n_gauges)     # spatially varying (cfg 3, 5)
    row_of = {"k": 0}

    def rain_resident(now):
        r = storm_rain(now)
        return gpu.GaugeRain(r * gauge_shape) if cfg != 2 else r          # cfg 2: tRainfall::NewRain(double)

    def rain_pinned(now):
        row = pinned_rows[row_of["k"] % B]
        row_of["k"] += 1
        row.copy_(torch.from_numpy(storm_rain(now) * gauge_shape))
        return gpu.GaugeRain(row.numpy())

    def resident_batch(k):
        sim.run_batch(k, rain_of_time=rain_resident)      # with ET on, kernel (1) rides inside the batch (tribs_gpu_run_et)

    # A third leg repeats e2e with the channels routed on the device as well (tKinemat's kinematic-wave reaches, SURVEY
    # row f2: one launch per batch, the outlet discharge per step comes back instead of the lateral inflows): the
    # reference's SurfaceFlow, which its arm times, includes them.
    channels = not args.no_channels and args.stream_every == 0
    if channels:
        sim.flow.route_channels(basin)
        sim.flow.on_device = False        # `value` and `e2e` time the per-cell path, whose boundary is the lateral inflow

    def e2e_batch(k):
        row_of["k"] = 0
        with_channels = sim.flow.on_device
        sim.run_batch(k, rain_of_time=rain_pinned, collect_qeff=not with_channels)   # H2D: k x n_gauges rain values (+ station records)
        g = dev.globals()                                               # D2H: lateral inflows (or outlet discharge) per step, accumulators
        return (sim.flow.qout_series if with_channels else sim.batch_qeff), g

    def loop(fn_batch, k):
        done = 0
        while done < k:
            kk = min(B, k - done)
            fn_batch(kk)
            done += kk

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record(stream)
        fn(k)
        e1.record(stream)
        dev.wait()
        w1 = time.perf_counter()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, (w1 - w0) * 1e3

    # warm-up, then `value` and `e2e` over the SAME simulated steps (warmup+1 .. warmup+steps, what the reference
    # arm times too): the state after the warm-up is kept on the device and restored between the two legs
    loop(resident_batch, args.warmup)
    dev.wait()
    say("warm-up done")
    dev.snapshot()
    sim.host_snapshot()
    dev.set_timing(True)
    dev.phase_ms(reset=True)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = dev.launch_count()
    ms, wall_ms = timed(lambda k: loop(resident_batch, k), args.steps)
    launches = dev.launch_count() - l0
    clocks = sampler.stop()
    phases = dev.phase_ms(reset=True)
    dev.set_timing(False)
    # dominant kernel: the pipelined batch kernel, which runs the unsaturated sweep of every step and the
    # saturated zone of every gw_every-th step
    n_launch = math.ceil(args.steps / B)
    unsat_ms = phases["unsat"] / max(n_launch, 1)
    steps_per_launch = args.steps / n_launch
    bytes_per_cell_step = UNSAT_BYTES_PER_CELL + SAT_BYTES_PER_CELL / gw_every + ({3: ET_BYTES_PER_CELL, 5: SNOW_BYTES_PER_CELL}.get(cfg, 0.0) / 16.0)
    bytes_per_launch = n * steps_per_launch * bytes_per_cell_step

    say("resident leg timed")
    dev.rollback()
    sim.host_restore()
    ems, ewall = timed(lambda k: loop(e2e_batch, k), args.steps)
    chan_ms = None
    if channels:
        dev.rollback()
        sim.host_restore()
        sim.flow.on_device = True
        cms, cwall = timed(lambda k: loop(e2e_batch, k), args.steps)
        chan_ms = max(cms, cwall)
        chan_state = dev.channel_sync()
    v0 = args.warmup + 1
    windows = {"value": [v0, v0 + args.steps - 1], "e2e": [v0, v0 + args.steps - 1], "storm_ends_at_step": 96}
    # host work (numpy fill, ctypes) is part of the user-visible e2e time: use the wall clock
    e2e_ms = max(ems, ewall)
    if world > 1:
        tt = torch.tensor([e2e_ms, chan_ms or 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms, chan_ms = float(tt[0].item()), (float(tt[1].item()) if channels else None)

    if world > 1:   # cells advanced by all ranks together
        tt = torch.tensor([float(n)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt)
        total_cells = int(tt.item())
    else:
        total_cells = n
    value = total_cells * args.steps / (ms * 1e-3)
    e2e = total_cells * args.steps / (e2e_ms * 1e-3)
    peak, peak_kind = peaks()
    achieved = bytes_per_launch / (unsat_ms * 1e-3) / 1e9 if unsat_ms > 0 else 0.0
    traffic = None      # DRAM bytes of one launch: bytes per cell-timestep of the committed ncu capture x this launch's units
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["pipeline_kernel"]
        traffic = tr["dram_bytes_per_cell_step"] * n * steps_per_launch
    except (OSError, KeyError, ValueError):
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD[cfg], "baseline_config": f"configs[{cfg - 1}]",
                   "cells_per_gpu": n, "total_cells": total_cells, "stream_cells": n_stream, "sweep_levels": dev.levels(),
                   "l2": "state arrays (54 x 8 B x cells) exceed the 126 MB L2; no explicit flush",
                   "batch_steps": B, "timed_steps": windows,
                   "parallelism": "1 GPU" if world == 1 else (
                       f"one basin of {total_cells} cells, {world} reach partitions along the river (one per GPU, "
                       f"{n} cells on rank 0), pipelined batches of {B} steps; river inflow, water-table and downslope-state "
                       "halos stored into peer memory by the batch kernel, dependent work released with system-scope "
                       "atomics over NVLink (no host exchange, no NCCL on the data path)" if partitioned
                       else f"{world} independent basin replicas, no exchange")},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 8 * n_gauges * world,
                "d2h_bytes_per_step": (8 * (n_stream + 1) + 8 * len(gpu.GLOBALS)) * world},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "pipeline_kernel", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                     "peak_source": peak_kind, "bytes_per_cell_step": bytes_per_cell_step,
                     "avg_launch_ms": unsat_ms, "steps_per_launch": steps_per_launch,
                     "note": "not HBM-bound: a dependency-chained FP64 state machine; bound by instruction issue of the working "
                             f"warps and by the ridge-to-outlet chain ({dev.levels()} levels) at the start and end of a batch; "
                             "see DESIGN.md 4.2"},
        "phase_ms_per_step": {k: v / args.steps for k, v in phases.items()},
    }
    say("end-to-end leg timed")
    if channels:
        line["e2e_with_channels"] = {
            "value": total_cells * args.steps / (chan_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * n_gauges * world,
            "d2h_bytes_per_step": (8 + 8 * len(gpu.GLOBALS)) * world,
            "what": "e2e plus tribs_gpu_channel_route after every batch (kinematic-wave reaches on the device, every rank routes "
                    "the whole net); the outlet discharge per step replaces the lateral inflows in the D2H copy",
            "stream_nodes": n_stream + 1, "newton_iterations": chan_state["newton_iterations"],
            "outlet_discharge_m3s": float(sim.flow.Qout)}
    if check is not None:
        line["self_check"] = check
    if world == 1 and not args.no_et_kernel and cfg == 2:
        line["kernels"] = {"et_cells_kernel": time_et_kernel(dev, basin, n, stream, peak)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample(min(side, 1000), cfg)
    if rank == 0:
        print(json.dumps(line))
    dev.close()
    if world > 1:
        dist.destroy_process_group()


def time_et_kernel(dev, basin, n, stream, peak, calls=24):
    """Kernel (1) timed alone on the state the water-balance run left behind: one fused
    potential-evaporation + ET-partition + Rutter-interception sweep per call (what the reference
    does once per METSTEP), daytime and raining so that every branch works. Reported next to the
    headline because it is the path's streaming, dependency-free kernel."""
    import numpy as np
    import torch
    from tribs_b200 import met
    from tribs_b200.synthbasin import add_met_forcing
    add_met_forcing(basin)
    dev.et_init(basin)
    host = met.EvapoTransHost(basin)
    rain = np.full(dev.n, 6.0)
    dev.step(0, 0.0625, 1, 0, rain=rain)

    def call(hour):
        pot = host.potential_call(hour % 24, 0.0625 + hour)
        cmp_ = host.components_call(0.0625 + hour)
        dev.et_step(potential=pot, components=cmp_, cell_record=host.cell_record, elev=host.elev)

    for h in range(3):
        call(h)
    host.hourlyTimeStep = 10            # mid-morning records: sun up
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_calls(k):
        torch.cuda.synchronize()
        ms = 0.0
        for h in range(k):
            host.hourlyTimeStep = 10 + (h % 6)
            pot = host.potential_call(10 + (h % 6), 10.0625 + h)
            cmp_ = host.components_call(10.0625 + h)
            e0.record(stream)
            dev.et_step(potential=pot, components=cmp_, cell_record=None, elev=host.elev)
            e1.record(stream)
            dev.wait()
            ms += e0.elapsed_time(e1)
        return ms / k

    ms = timed_calls(calls)
    # dry spell: no rain, empty canopy -> the Rutter ODE is skipped, the energy balance remains
    dev.step(0, 0.0625, 1, 0, rain=np.zeros(dev.n))
    dev.et_push({"CanStorage": np.zeros(dev.n)})
    timed_calls(2)
    ms_dry = timed_calls(8)
    gbs = n * ET_BYTES_PER_CELL / (ms * 1e-3) / 1e9
    gbs_dry = n * ET_BYTES_PER_CELL / (ms_dry * 1e-3) / 1e9
    return {"avg_launch_ms": ms, "bytes_per_cell_call": ET_BYTES_PER_CELL, "achieved": gbs, "unit": "GB/s",
            "frac": gbs / peak if peak else None, "cells_per_s": n / (ms * 1e-3),
            "dry": {"avg_launch_ms": ms_dry, "achieved": gbs_dry, "frac": gbs_dry / peak if peak else None},
            "note": "raining + wet canopy (Rutter RK4, >= 200 exp per cell: FP64-bound) vs dry; includes the H2D copy "
                    "of the hour's station records (72 B per station) issued inside the call"}


# -------------------------------------------------------------------------------- CPU arms
def _harness():
    from refrun import HARNESS, have_harness
    return HARNESS if have_harness() else None


def _run_reference_sample(n_side: int, hours: float, procs: int = 1):
    """Runs `procs` concurrent copies of the unmodified serial reference on an n_side^2-point
    V-valley deck for `hours`; returns aggregate cell-timesteps/s of their simulation loops."""
    import tempfile
    from refrun import HARNESS
    from tribs_b200.synth import BasinSpec, write_basin
    ps, dirs = [], []
    for k in range(procs):
        d = tempfile.mkdtemp(prefix="tribsbench_")
        write_basin(BasinSpec(n=n_side, runtime_h=hours, seed=1 + k), d)
        dirs.append(d)
    t0 = time.perf_counter()
    for d in dirs:
        ps.append(subprocess.Popen([HARNESS, "basin.in", "-K", "--out", d, "--no-basin"], cwd=d,
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True))
    rate, cells, steps = 0.0, 0, 0
    for p in ps:
        out = p.communicate()[0]
        for ln in out.splitlines():
            if ln.startswith("ref_timing"):
                j = json.loads(ln[len("ref_timing "):])
                rate += j["cell_steps_per_s"]
                cells, steps = j["cells"], j["steps"]
    wall = time.perf_counter() - t0
    import shutil
    for d in dirs:
        shutil.rmtree(d, ignore_errors=True)
    return rate, cells, steps, wall


def _reference_on_bench_basin(side: int, steps: int, cfg: int = 2):
    """The unmodified serial reference on the GPU arm's own basin: build_basin(side, side, seed=1) is
    written as OPTMESHINPUT 9 files (tribs_b200/meshb.py), which the reference reads without its
    super-linear mesh / flow-network set-up, and runs the first `steps` steps of the storm."""
    import shutil
    import tempfile
    from refrun import HARNESS, write_meshb_deck
    d = tempfile.mkdtemp(prefix="tribsbench_")
    t0 = time.perf_counter()
    try:
        write_meshb_deck(side, d, seed=1, runtime_h=24.0, scenario={3: "met", 5: "snow"}.get(cfg, "deep"),
                         **({"extra": {"OPTINTERCEPT": 0}} if cfg == 5 else {}))
        r = subprocess.run([HARNESS, "basin.in", "-K", "--out", d, "--no-basin", "--max-steps", str(steps)], cwd=d,
                           stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=900)
        for ln in r.stdout.splitlines():
            if ln.startswith("ref_timing"):
                j = json.loads(ln[len("ref_timing "):])
                return j["cell_steps_per_s"], j["cells"], j["steps"], j.get("setup_s", 0.0), time.perf_counter() - t0
    except (OSError, subprocess.SubprocessError):
        pass
    finally:
        shutil.rmtree(d, ignore_errors=True)
    return None


def cpu_baseline_sample(side: int = 1000, cfg: int = 2):
    if _harness():
        got = _reference_on_bench_basin(side, 20 if cfg != 2 else 8, cfg)
        if got:
            rate, cells, steps, setup_s, wall = got
            return {"value": rate, "unit": UNIT, "cores": 1, "kind": "reference",
                    "sample": f"unmodified serial reference (g++ -O2) on a basin of the GPU arm's generator ({cells} cells, read as "
                              f"OPTMESHINPUT 9 files" + ({3: ", rain gauges + weather station + ET + Rutter", 5: ", winter station + snow pack"}.get(cfg, "")) +
                              f"), the first {steps} steps, simulation_loop only; "
                              f"{wall:.0f} s wall incl. writing the deck and {setup_s:.0f} s of reference set-up"}
        rate, cells, steps, wall = _run_reference_sample(120, 12.0, 1)
        return {"value": rate, "unit": UNIT, "cores": 1, "kind": "reference",
                "sample": f"unmodified serial reference (g++ -O2), {cells}-cell V-valley of the same generator, "
                          f"{steps} steps (12 h, storm), simulation_loop only; {wall:.1f} s wall incl. setup"}
    import numpy as np
    from oracle.pyoracle import OracleModel
    from tribs_b200.synthbasin import build_basin
    b = build_basin(160, 160)
    m = OracleModel(b, lambda t: storm_rain(t))
    t0 = time.perf_counter()
    for _ in range(96):
        m.step()
    dt = time.perf_counter() - t0
    return {"value": int(b["n_active"][0]) * 96 / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "C oracle (bit-exact restatement), 25.6k-cell synthetic V-valley, 96 steps"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    total = args.steps + args.warmup
    if _harness():
        # One unmodified serial reference process per host core (no MPI toolchain on the box: the optimistic
        # all-core bound), each on a basin of the GPU arm's generator read as OPTMESHINPUT 9 files. A "step" is one
        # simulated TIMESTEP of all processes together; the first `warmup` steps are left out of the timing
        # (--time-from). The basin side is bounded so that the run ends within a few minutes.
        import shutil
        import tempfile
        from refrun import HARNESS, write_meshb_deck
        procs = max(1, min(cores, 32))
        try:
            avail_gb = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") / 2 ** 30
        except (ValueError, OSError):
            avail_gb = 64.0
        want = args.side if args.side > 0 else {2: 1000, 3: 3163, 5: 2500}[args.config]
        side = int(min(want, math.sqrt(150.0 * 0.3e6 / total), math.sqrt(0.5 * avail_gb / procs / 6e-6)))
        side = max(side, 64)
        base = tempfile.mkdtemp(prefix="tribsbench_ref_")
        src = os.path.join(base, "deck")
        os.makedirs(src)
        write_meshb_deck(side, src, seed=1, runtime_h=24.0, scenario={3: "met", 5: "snow"}.get(args.config, "deep"),
                         **({"extra": {"OPTINTERCEPT": 0}} if args.config == 5 else {}))
        dirs = []
        for k in range(procs):
            d = os.path.join(base, f"p{k}")
            shutil.copytree(src, d)
            dirs.append(d)
        t0 = time.perf_counter()
        ps = [subprocess.Popen([HARNESS, "basin.in", "-K", "--out", d, "--no-basin", "--max-steps", str(total),
                                "--time-from", str(args.warmup)], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
              for d in dirs]
        rate, loop_s, cells, steps = 0.0, 0.0, 0, 0
        for p in ps:
            for ln in p.communicate()[0].splitlines():
                if ln.startswith("ref_timing"):
                    j = json.loads(ln[len("ref_timing "):])
                    rate += j["cell_steps_per_s"]
                    loop_s = max(loop_s, j["loop_s"])
                    cells, steps = j["cells"], j["steps"]
        wall = time.perf_counter() - t0
        shutil.rmtree(base, ignore_errors=True)
        if steps != args.steps:
            raise SystemExit("bench.py --impl reference: a reference process did not finish")
        value, ms, kind = rate, loop_s / steps * 1e3, "reference"
        sample = (f"{procs} concurrent processes of the unmodified serial reference (no MPI toolchain on the box, so "
                  f"this is the optimistic all-core bound), each on a {cells}-cell basin of the GPU arm's generator read "
                  f"as OPTMESHINPUT 9 files; {args.warmup} warm-up + {steps} timed simulation steps of the storm, "
                  f"simulation_loop only; {wall:.0f} s wall incl. set-up")
    else:
        from oracle.pyoracle import OracleModel
        from tribs_b200.synthbasin import build_basin
        b = build_basin(96, 96)
        m = OracleModel(b, lambda t: storm_rain(t))
        vals = []
        for k in range(total):
            t0 = time.perf_counter()
            for _ in range(32):
                m.step()
            dt = time.perf_counter() - t0
            if k >= args.warmup:
                vals.append(int(b["n_active"][0]) * 32 / dt)
        value, ms, kind, procs = statistics.mean(vals), 0.0, "port", 1
        sample = "C oracle port, 9.2k-cell synthetic V-valley, 32 steps per sample"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if args.config == 3 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD[args.config], "baseline_config": f"configs[{args.config - 1}]",
                       "sample": "bounded CPU sample of the GPU arm's workload, see cpu_baseline"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=32)
    ap.add_argument("--impl", default="tribs_b200", choices=["tribs_b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[2, 3, 5],
                    help="BASELINE.json configs[1] (1M-cell valley per GPU, uniform storm, ET off; grows with the GPUs) or "
                         "configs[2] (one 10M-cell valley on all GPUs, gauge rain + weather station + ET + Rutter) or "
                         "configs[4] (6.25M cells per GPU, snow pack + shallow water table, weak scaling)")
    ap.add_argument("--side", type=int, default=0, help="cells per side of the synthetic basin (default: 1000 / 3163)")
    ap.add_argument("--batch", type=int, default=256, help="steps per pipelined launch (tribs_gpu_run), at most --steps")
    ap.add_argument("--mgpu", default="partition", choices=["partition", "replicas"],
                    help="N>1: one reach-partitioned basin with halo exchange (default) or independent replicas")
    ap.add_argument("--no-channels", action="store_true", help="e2e leg: leave the channel routing to the host (copy the lateral inflows out)")
    ap.add_argument("--no-self-check", action="store_true", help="N>1: skip the bit-equality check against one GPU")
    ap.add_argument("--stream-every", type=int, default=0, help="tributary stream row every K rows (0: one river, "
                    "hillslopes as long as half the basin width)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-et-kernel", action="store_true", help="skip the separate timing of kernel (1)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
