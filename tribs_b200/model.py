"""Host-side mirror of the reference call surface for the water-balance path.

`tHydroModel`, `tKinemat` and `Simulator` below keep the reference's public method names,
argument meaning and calling order (reference: src/tHydro/tHydroModel.h:45-106,
src/tFlowNet/tKinemat.h:127-131, src/tSimulator/tSimul.cpp:220-295, 506-520) but their bodies
only forward to the C ABI (include/tribs_gpu.h). The decisions the reference takes on the host
with its own timer arithmetic (`fmod` on accumulated hours, tRunTimer.cpp:436-452) are taken
here, on the host, in the same way and passed down as flags.
"""
from __future__ import annotations

import math

import numpy as np

from . import gpu, met


class tRunTimer:
    """The subset of the reference timer the path depends on (src/tSimulator/tRunTimer.cpp)."""

    def __init__(self, tstep_h, gwstep_h, etistep_h, end_h):
        self.tstep, self.gwstep, self.etistep, self.endTime = tstep_h, gwstep_h, etistep_h, end_h
        self.currentTime = 0.0

        self.calendar = None        # met.Calendar once a start date is known

    def Advance(self, dt):
        self.currentTime += dt
        if self.calendar is not None:
            self.calendar.advance(self.currentTime, dt)
        return self.currentTime < self.endTime

    @property
    def hour(self):
        return self.calendar.hour if self.calendar is not None else 0

    def getElapsedMETSteps(self, t):
        return int(t / self.metstep)

    def getElapsedETISteps(self, t):
        return int(t / self.etistep)

    def IsFinished(self):
        return self.currentTime >= self.endTime

    def getCurrentTime(self):
        return self.currentTime

    def getTimeStep(self):
        return self.tstep

    def getGWTimeStep(self):
        return self.gwstep

    def getElapsedSteps(self, t):
        return int(t / self.tstep)


class tHydroModel:
    """UnSaturatedZone / ResetGW / SaturatedZone / Reset forwarded to the GPU."""

    def __init__(self, dev: gpu.GpuBasin, timer: tRunTimer):
        self.dev, self.timer = dev, timer
        self._gw_reset_pending = False

    def _flags(self):
        t = self.timer
        now = t.getCurrentTime()
        hourly = int(not math.fmod(now - t.getTimeStep(), t.etistep))   # tHydroModel.cpp:700-702
        return now, t.getElapsedSteps(now), hourly

    def UnSaturatedZone(self, dt):
        now, el, hr = self._flags()
        self.dev.step(gpu.PHASE_UNSAT, now, el, hr)

    def ResetGW(self):
        # folded into the SAT phase (a pointer swap on the device); kept for call-surface parity
        self._gw_reset_pending = True

    def SaturatedZone(self, dtGW):
        if not self._gw_reset_pending:
            raise RuntimeError("SaturatedZone() must follow ResetGW(), as in Simulator::SubSurfaceHydroProcesses")
        self._gw_reset_pending = False
        now, el, hr = self._flags()
        self.dev.step(gpu.PHASE_SAT, now, el, hr)

    def Reset(self):
        now, el, hr = self._flags()
        self.dev.step(gpu.PHASE_RESET, now, el, hr)


class tIntercept:
    """Only the option query the simulator makes (tIntercept.cpp:555-563); the interception
    models themselves run inside tEvapoTrans.callEvapoTrans on the device."""

    def __init__(self, basin: dict):
        self.interceptOption = int(basin["et_Ioption"][0]) if "et_Ioption" in basin else 0

    def getIoption(self):
        return self.interceptOption


class tEvapoTrans:
    """callEvapoPotential / callEvapoTrans forwarded to kernel (1). The per-call members of the
    reference class (met clock, sun geometry, record-0 decisions) live in met.EvapoTransHost."""

    def __init__(self, dev: gpu.GpuBasin, timer: tRunTimer, basin: dict):
        self.dev, self.timer = dev, timer
        self.evapotransOption = int(basin["et_option"][0]) if "et_option" in basin else 0
        self.host = None
        self._pending = None
        if self.evapotransOption:
            if int(basin["rain_optStorm"][0]):
                raise NotImplementedError("stochastic weather generator (tHydroMetStoch) is not part of the path")
            if int(basin["et_metdataOption"][0]) != 1:
                raise NotImplementedError("only METDATAOPTION 1 (stations) is wired on the host side")
            self.host = met.EvapoTransHost(basin)
            timer.metstep = self.host.metstep
            timer.calendar = met.Calendar(basin["timer_start"])
            dev.et_init(basin)
            self.snowOption = int(basin["et_snowOption"][0]) if "et_snowOption" in basin else 0
            if self.snowOption:
                dev.snow_init(basin)

    snowOption = 0

    def getEToption(self):
        return self.evapotransOption

    def getSnowOpt(self):
        return self.snowOption

    def callSnowPack(self, intercept, flag: int):
        """tSnowPack::callSnowPack(tIntercept*, int): one sweep that replaces both ET sweeps."""
        now = self.timer.getCurrentTime()
        hourly = self.host.hourlyTimeStep
        step, rec = self.host.potential_call(self.timer.hour, now)
        step["te_eti"] = float(self.timer.getElapsedETISteps(now))
        self.dev.snow_step(step, rec, hourly, cell_record=self.host.cell_record, elev=self.host.elev)

    def callEvapoPotential(self):
        """Queues the potential-evaporation sweep; it is launched together with callEvapoTrans when
        both are due in the same step (the usual case), otherwise by flush()."""
        self._pending = self.host.potential_call(self.timer.hour, self.timer.getCurrentTime())

    def callEvapoTrans(self, intercept: tIntercept, flag: int):
        comp = self.host.components_call(self.timer.getCurrentTime())
        self.dev.et_step(potential=self._pending, components=comp,
                         cell_record=self.host.cell_record, elev=self.host.elev)
        self._pending = None

    def flush(self):
        if self._pending is not None:
            self.dev.et_step(potential=self._pending, components=None,
                             cell_record=self.host.cell_record, elev=self.host.elev)
            self._pending = None


class tKinemat:
    """SurfaceFlow(): hillslope routing + basin accumulations on the GPU. The kinematic-wave channel solver either
    stays with the host model and consumes `lateral_inflow`, or - after route_channels() - runs on the device too
    (tribs_gpu_channel_route) and leaves the outlet discharge of every step in `Qout` / `qout_series`."""

    def __init__(self, dev: gpu.GpuBasin, timer: tRunTimer):
        self.dev, self.timer = dev, timer
        self.lateral_inflow = None
        self.on_device = False
        self.defer = False           # True: the caller routes the channels itself once every in-process rank has launched
        self.pending_steps: list[int] = []
        self.Qout = 0.0
        self.qout_series: list[float] = []

    def route_channels(self, basin: dict, net: dict | None = None):
        """Hands tKinemat's reach lists, channel geometry and initial levels to the device (tKinemat.cpp:826-830,
        1136-1252). From here on SurfaceFlow() / Simulator.run_batch() route the channels as well. `net`: reach lists other
        than the basin's own (channel.channel_network's layout)."""
        from .channel import channel_network
        n = int(basin["n_active"][0])

        def init(key):
            if key not in basin:
                return None
            a = np.zeros(n + 1)
            a[:n] = np.asarray(basin[key], float)[:n]
            return a
        self.dev.channel_init(channel_network(basin) if net is None else net, self.timer.getTimeStep() * 3600.0, init("init_Hlevel"), init("init_Qstrm"),
                              percolation_option=int(basin["chan_percolation"][0]) if "chan_percolation" in basin else 0,
                              reservoir_option=int(basin["chan_optres"][0]) if "chan_optres" in basin else 0)
        self.on_device = True

    def channels_after_batch(self, n_steps: int, deferred: bool = False):
        """tribs_gpu_channel_route waits for the batch; ranks that live in one process (peers.LocalPeers) must all have
        launched theirs before any of them waits, so there the call is deferred to the end of LocalPeers.run_batch."""
        if self.on_device and n_steps:
            if self.defer and not deferred:
                self.pending_steps.append(n_steps)
                return
            q = self.dev.channel_route(n_steps)
            self.qout_series.extend(q.tolist())
            self.Qout = float(q[-1])

    def SurfaceFlow(self, fetch_qeff: bool = True):
        now = self.timer.getCurrentTime()
        self.dev.step(gpu.PHASE_ROUTE, now, self.timer.getElapsedSteps(now), 0)
        if fetch_qeff:
            self.lateral_inflow = self.dev.retrieve_qeff(now)
        if self.on_device:
            self.Qout = float(self.dev.channel_route(0)[0])
            self.qout_series.append(self.Qout)


class tCOutput:
    """tCOutput::WriteDynamicVars (tOutput.cpp:1238-1400) from the device state: the fields the file
    needs are synced (only when an output is due), the reference's text format is written by
    output.write_dynamic_vars. Qstrm / Hlev belong to the host's channel router and are passed in."""

    NEEDS = ["NwtNew", "MuNew", "MiNew", "NfNew", "NtNew", "Qpout", "Qpin", "srf_hr", "Rain", "SnTempC", "IceWE", "LiqWE",
             "SnSub", "SnEvap", "LiqRouted", "Unode", "SnLHF", "SnSHF", "SnGHF", "SnPHF", "SnRLout", "SnRLin", "SnRSin",
             "Uerror", "IntSWE", "IntSub", "IntSnUnload", "SoilMoistureSC", "RootMoistureSC", "CanStorage", "ActEvap",
             "EvapSoil", "EvapoTrans", "GFlux", "HFlux", "LFlux", "FlowVelocity"]

    def __init__(self, dev: gpu.GpuBasin, basin: dict, prefix: str, et_on: bool, snow_on: bool):
        self.dev, self.basin, self.prefix, self.et_on, self.snow_on = dev, basin, prefix, et_on, snow_on

    def WriteDynamicVars(self, time_h: float, channel: dict | None = None) -> str:
        from . import output
        main = [nm for nm in self.NEEDS if nm in gpu.F]
        state = self.dev.sync(main)
        if self.et_on:
            state.update(self.dev.et_sync([nm for nm in self.NEEDS if nm in gpu.EF and nm not in gpu.F]))
        if self.snow_on:
            state.update(self.dev.snow_sync([nm for nm in self.NEEDS if nm in gpu.SNF and nm not in gpu.F and nm not in gpu.EF]))
        state.update(channel or {})
        path = output.dynamic_vars_name(self.prefix, time_h)
        output.write_dynamic_vars(path, self.basin, state)
        return path


class Simulator:
    """simulation_loop() for the path: Advance, rain, UnSat, [ResetGW, Sat], SurfaceFlow, Reset."""

    def __init__(self, basin: dict, rain_schedule=None, gw_on: bool | None = None, part: dict | None = None):
        self.dev = gpu.GpuBasin(basin, part=part)
        b = basin
        self.timer = tRunTimer(float(b["tstep"][0]), float(b["gwstep"][0]), float(b["etistep"][0]),
                               float(b["endtime"][0]))
        self.hydro = tHydroModel(self.dev, self.timer)
        self.flow = tKinemat(self.dev, self.timer)
        self.intercept = tIntercept(b)
        self.evapotrans = tEvapoTrans(self.dev, self.timer, b)
        self.met_clock = met.MetClock(self.timer.tstep, float(b["metstep"][0]) if "metstep" in b else self.timer.etistep, self.timer.etistep)
        self.rain_schedule = rain_schedule
        self.gw_on = bool(int(b["GW_model_label"][0])) if gw_on is None else gw_on
        self.step_no = 0
        self._record_sent = [False, False]
        self.et_in_batch = True      # False: cut the batches at the ET sweeps and run kernel (1) between them

    def host_snapshot(self):
        """Host-side counterpart of GpuBasin.snapshot(): the timer, the met clock and tEvapoTrans' per-call members."""
        import copy
        self._host_snap = copy.deepcopy((self.timer.currentTime, self.timer.calendar, self.met_clock,
                                         getattr(self.evapotrans.host, "__dict__", None), self.step_no, list(self._record_sent)))

    def host_restore(self):
        import copy
        now, cal, clock, host, step_no, sent = copy.deepcopy(self._host_snap)
        self.timer.currentTime, self.timer.calendar, self.met_clock, self.step_no, self._record_sent = now, cal, clock, step_no, sent
        if host is not None:
            self.evapotrans.host.__dict__.update(host)

    def UpdatePrecipitationInput(self):
        if self.rain_schedule is None:
            return
        r = self.rain_schedule(self.timer.getCurrentTime())
        now = self.timer.getCurrentTime()
        self.dev.step(0, now, self.timer.getElapsedSteps(now), 0, rain=r)

    def SurfaceHydroProcesses(self):
        """tSimul.cpp:419-498 without the snow branch: which of the two ET sweeps are due."""
        met_due, eti_due = self.met_clock.due(self.timer.getCurrentTime())
        et, icp = self.evapotrans, self.intercept
        if et.getEToption() != 0 and et.getSnowOpt() != 0:       # snow active: tSimul.cpp:464-492
            if met_due:
                et.callSnowPack(icp, 1 if icp.getIoption() != 0 else 0)
        elif et.getEToption() != 0:
            if met_due:
                et.callEvapoPotential()
            if eti_due:
                et.callEvapoTrans(icp, 1 if icp.getIoption() != 0 else 0)
            et.flush()
        elif icp.getIoption() != 0:
            raise NotImplementedError("interception without the evaporation scheme needs tEvapoTrans' land data wiring")

    def SubSurfaceHydroProcesses(self):
        t = self.timer
        self.hydro.UnSaturatedZone(t.getTimeStep())
        gw_label = math.fmod(t.getCurrentTime(), t.getGWTimeStep())      # tSimul.cpp:511
        if self.gw_on and not gw_label:
            self.hydro.ResetGW()
            self.hydro.SaturatedZone(t.getGWTimeStep())
            return True
        return False

    def advance(self):
        self.timer.Advance(self.timer.getTimeStep())
        self.step_no += 1

    def one_step(self, fetch_qeff=True):
        self.advance()
        self.UpdatePrecipitationInput()
        self.SurfaceHydroProcesses()
        did_gw = self.SubSurfaceHydroProcesses()
        self.flow.SurfaceFlow(fetch_qeff)
        self.hydro.Reset()
        return did_gw

    def run_batch(self, k: int, rain_of_time=None, per_cell_rain=None, collect_qeff: bool = False):
        """k whole steps in pipelined launches (tribs_gpu_run). The host-side decisions of every
        step (GW due, hourly reset, elapsed steps, rain) are taken here exactly as one_step()
        would take them, then handed down together. With the evaporation scheme on, a batch ends
        where SurfaceHydroProcesses changes the inputs of the sweep (every METSTEP / ET-I step):
        kernel (1) runs there and the next launch starts with that step."""
        t = self.timer
        steps, out = [], []
        et_on = self.evapotrans.getEToption() != 0

        qeff = []

        def flush():
            if steps:
                self.dev.run(steps)
                if collect_qeff:      # per-step lateral inflows for a channel router on the host
                    qeff.append(self.dev.retrieve_qeff_steps(len(steps)))
                self.flow.channels_after_batch(len(steps))
                out.extend(steps)
                steps.clear()

        for _ in range(k):
            self.advance()
            now = t.getCurrentTime()
            sched = rain_of_time if rain_of_time is not None else self.rain_schedule
            rain = sched(now) if sched is not None else None
            if per_cell_rain is not None:
                per_cell_rain.fill(0.0 if rain is None else rain)
                rain = per_cell_rain.copy()
            # per-cell arrays are uploaded every step: an identity cache cannot see in-place refills, scalar
            # steps in between or uploads made by other entry points (a static gauge map + K gauge values per
            # step is the cheap form of spatially varying rain, see tribs_gpu_set_rain_gauges)
            step_rain = rain
            et_arg = snow_arg = None
            if et_on and self.evapotrans.getSnowOpt() != 0 and self.intercept.getIoption() == 0 and self.et_in_batch:
                # OPTSNOW 1 without a canopy: callSnowPack needs nothing but the cell itself and rides inside the batch
                met_due, _ = self.met_clock.due(now)
                if met_due:
                    host = self.evapotrans.host
                    hourly = host.hourlyTimeStep
                    step, rec = host.potential_call(t.hour, now)
                    step["te_eti"] = float(t.getElapsedETISteps(now))
                    snow_arg = (step, rec, hourly, None if self._record_sent[0] else host.cell_record, host.elev)
                    self._record_sent[0] = True
            elif et_on and self.evapotrans.getSnowOpt() == 0 and self.et_in_batch:
                # kernel (1) rides inside the batch as one more phase of every tile (tribs_gpu_run_et)
                met_due, eti_due = self.met_clock.due(now)
                if met_due or eti_due:
                    host = self.evapotrans.host
                    pot = host.potential_call(t.hour, now) if met_due else None
                    comp = host.components_call(now) if eti_due else None
                    # the cell -> station map is static: it travels once per sweep kind (the library keeps it)
                    need = (pot is not None and not self._record_sent[0]) or (comp is not None and not self._record_sent[1])
                    et_arg = (pot, comp, host.cell_record if need else None, host.elev)
                    if need:
                        self._record_sent = [self._record_sent[0] or pot is not None, self._record_sent[1] or comp is not None]
            elif et_on:
                met_due, eti_due = self.met_clock.due(now)
                if met_due or eti_due:
                    flush()
                    if step_rain is not None:      # the ET sweep reads the Rain array of this step
                        self.dev.step(0, now, t.getElapsedSteps(now), 0, rain=step_rain)
                        step_rain = None if isinstance(rain, np.ndarray) else step_rain
                    if self.evapotrans.getSnowOpt() != 0:
                        if met_due:
                            self.evapotrans.callSnowPack(self.intercept, 0)
                    else:
                        if met_due:
                            self.evapotrans.callEvapoPotential()
                        if eti_due:
                            self.evapotrans.callEvapoTrans(self.intercept, 1 if self.intercept.getIoption() != 0 else 0)
                        self.evapotrans.flush()
            steps.append(dict(current_time=now, elapsed_steps=t.getElapsedSteps(now),
                              hourly_reset=int(not math.fmod(now - t.getTimeStep(), t.etistep)),
                              sat=bool(self.gw_on and not math.fmod(now, t.getGWTimeStep())),
                              rain=step_rain, et=et_arg, snow=snow_arg))
        flush()
        self.batch_qeff = np.concatenate(qeff) if qeff else None
        return out

    def simulation_loop(self, max_steps=None):
        k = 0
        while not self.timer.IsFinished() and (max_steps is None or k < max_steps):
            self.one_step()
            k += 1
        return k
