"""Shared comparison logic: step a model phase by phase and compare with reference snapshots."""
from __future__ import annotations

import math

import numpy as np

from refrun import evap_of, has_et, rel_err

STATE_FIELDS = ["NwtNew", "MuNew", "MiNew", "NtNew", "NfNew", "RuNew", "RiNew", "Qpin", "Qpout", "intstorm",
                "srf", "hsrf", "sbsrf", "psrf", "esrf", "srf_hr", "cumsrf", "SoilMoisture", "SoilMoistureSC",
                "SoilMoistureUNSC", "RootMoisture", "RootMoistureSC", "AvSoilMoisture", "Recharge", "UnSatFlowIn",
                "UnSatFlowOut", "hsrfOccur", "sbsrfOccur", "psrfOccur", "satOccur", "RechDisch", "Rain"]
GW_FIELDS = ["gwchange", "satsrf", "satsrfOccur", "Transmissivity"]
RESULT_ARRAYS = ["mhydro", "HsrfRout", "SbsrfRout", "PsrfRout", "SatsrfRout", "crr", "rmax", "rmin", "frac",
                 "msm", "msmRt", "msmU", "mgw", "sat", "qunsat"]


class GpuStepper:
    """Drives tribs_b200 through the C ABI one phase at a time."""

    def __init__(self, basin, rain_schedule):
        from tribs_b200 import gpu, model
        self.gpu = gpu
        self.sim = model.Simulator(basin, rain_schedule)
        self.n = self.sim.dev.n
        self.boundary = np.asarray(basin["boundary"])
        self.stream_cells = self.sim.dev.stream_cells()
        self.limit = int(basin["res_limit"][0])
        self.evap, self._evap_now = evap_of(basin), None
        self.has_et = self.sim.evapotrans.getEToption() != 0

    @property
    def time(self):
        return self.sim.timer.getCurrentTime()

    def advance(self):
        self.sim.advance()
        self.sim.UpdatePrecipitationInput()
        self.sim.SurfaceHydroProcesses()
        if self.evap > 0.0:   # the harness's injected soil evaporation: EvapSoil = rate on dry steps
            r = self.sim.rain_schedule(self.time)
            want = 0.0 if r > 0.0 else self.evap
            if want != self._evap_now:
                self.sim.dev.push({"EvapSoil": np.full(self.n, want)})
                self._evap_now = want

    def unsat(self):
        self.sim.hydro.UnSaturatedZone(self.sim.timer.getTimeStep())

    def gw_due(self):
        t = self.sim.timer
        return self.sim.gw_on and math.fmod(t.getCurrentTime(), t.getGWTimeStep()) == 0.0

    def sat(self):
        self.sim.hydro.ResetGW()
        self.sim.hydro.SaturatedZone(self.sim.timer.getGWTimeStep())

    def route(self):
        self.sim.flow.SurfaceFlow(True)
        return self.sim.flow.lateral_inflow

    def pending(self, n_bins):
        """[n_stream, n_bins] dtReff bins from the current one on: the device's (TimeInd, Qeff) stacks."""
        return self.sim.dev.pending_qeff(self.time, n_bins)[:-1]

    def balance(self):
        pass   # tWaterBalance rides on the ROUTE phase of the device path

    def reset(self):
        self.sim.hydro.Reset()

    def get(self, names):
        from tribs_b200.gpu import EF, F, SNF
        main = [nm for nm in names if nm in F]
        out = self.sim.dev.sync(main) if main else {}
        et = [nm for nm in names if nm in EF and nm not in F]
        if et:
            out.update(self.sim.dev.et_sync(et))
        sn = [nm for nm in names if nm in SNF and nm not in F and nm not in EF]
        if sn:
            out.update(self.sim.dev.snow_sync(sn))
        return out

    def result(self, name):
        return self.sim.dev.results(name, self.limit)

    def globals(self):
        return self.sim.dev.globals()


class OracleStepper:
    def __init__(self, basin, rain_schedule):
        from oracle.pyoracle import OracleModel
        self.m = OracleModel(basin, rain_schedule)
        self.n = self.m.n
        self.boundary = np.asarray(basin["boundary"])
        self.stream_cells = np.nonzero(self.boundary == 3)[0].astype(np.int32)
        self.evap = evap_of(basin)
        self.et = None
        if has_et(basin):
            from oracle.pyoracle import OracleET
            from tribs_b200 import met
            self.host = met.EvapoTransHost(basin)
            self.et = OracleET(basin, self.m, self.host)
            self.cal = met.Calendar(basin["timer_start"])
            self.clock = met.MetClock(self.m.B.tstep, self.host.metstep, self.host.etistep)
        self.has_et = self.et is not None

    @property
    def time(self):
        return self.m.time

    def advance(self):
        self.m.advance()
        if self.et is not None:       # Simulator::SurfaceHydroProcesses, tSimul.cpp:419-461
            self.cal.advance(self.m.time, self.m.B.tstep)
            met_due, eti_due = self.clock.due(self.m.time)
            self._met_due = met_due
            if self.et.snow:     # tSimul.cpp:464-492: callSnowPack at met_hour replaces both sweeps
                if met_due:
                    hourly = self.host.hourlyTimeStep
                    step, rec = self.host.potential_call(self.cal.hour, self.m.time)
                    step["te_eti"] = float(int(self.m.time / self.host.etistep))
                    self.et.snow_pack(step, rec, hourly)
            else:
                if met_due:
                    self.et.potential(*self.host.potential_call(self.cal.hour, self.m.time))
                if eti_due:
                    self.et.components(*self.host.components_call(self.m.time))
        if self.evap > 0.0:
            self.m.state["EvapSoil"][:] = np.where(self.m.state["Rain"] > 0.0, 0.0, self.evap)

    def unsat(self):
        self.m.unsat()

    def gw_due(self):
        return self.m.gw_due()

    def sat(self):
        self.m.sat()

    def route(self):
        self.m.route()
        q = self.m.retrieve_qeff()
        return np.concatenate([q[self.stream_cells], q[-1:]])

    def pending(self, n_bins):
        cur = int(math.ceil(self.time / self.m.B.dtReff))
        out = np.zeros((self.stream_cells.size, n_bins))
        for k, i in enumerate(self.stream_cells):
            for tind, q in zip(*self.m.stack_of(int(i))):
                if not 0 <= tind - cur < n_bins:
                    raise AssertionError(f"oracle stack of cell {i} holds bin {tind}, current bin {cur}")
                out[k, tind - cur] = q
        return out

    def balance(self):
        self.m.balance(self.et, getattr(self, "_met_due", False), self.host.metstep if self.et is not None else 1.0)

    def reset(self):
        self.m.reset()

    def get(self, names):
        return {k: (self.m.state[k] if k in self.m.state else self.et.f[k]) for k in names}

    def result(self, name):
        return self.m.res[name]

    def globals(self):
        from oracle.pyoracle import GLOBALS
        g = dict(zip(GLOBALS, self.m.glob))
        b6 = getattr(self.m, "basin6", np.zeros(6))
        g.update(BasinRain=b6[0], BasinEvap=b6[1], BasinRunoff=b6[5])
        return g


# absolute floors of the relative comparison for ET fields that pass through zero (degC, W/m2)
ET_FLOOR = {"SurfTemp": 1.0, "SoilTemp": 1.0, "AirTemp": 1.0, "DewTemp": 1.0, "NetRad": 1.0, "GFlux": 1.0,
            "HFlux": 1.0, "LFlux": 1.0}


class Worst:
    def __init__(self):
        self.w = {}

    def add(self, key, ref, got, info, floor):
        ref, got = np.ravel(ref), np.ravel(got)
        e = rel_err(ref, got, floor)
        k = int(np.argmax(e)) if e.size else 0
        v = float(e[k]) if e.size else 0.0
        if key not in self.w or v > self.w[key][0]:
            self.w[key] = (v, info, k, float(np.ravel(ref)[k]) if e.size else 0.0, float(np.ravel(got)[k]) if e.size else 0.0)

    def max(self):
        return max((v[0] for v in self.w.values()), default=0.0)

    def report(self, top=12):
        rows = sorted(self.w.items(), key=lambda kv: -kv[1][0])[:top]
        return "\n".join(f"  {k}: rel={v[0]:.3e} at {v[1]} cell {v[2]} ref={v[3]!r} got={v[4]!r}" for k, v in rows)


QEFF_BINS = 24      # dtReff bins compared from the current one on (hillslope travel times stay far below 12 h here)


def reference_stacks(snaps, pre, stream_cells, cur, n_bins=QEFF_BINS):
    """The reference's (TimeInd, Qeff) stacks of a snapshot (tKinemat.cpp:1023-1067, dumped after
    tKinemat::SurfaceFlow, i.e. after RetrieveQeff consumed what it consumes) as a dense
    [n_stream, n_bins] array of dtReff bins cur, cur+1, ..."""
    slot = {int(c): k for k, c in enumerate(stream_cells)}
    out = np.zeros((len(stream_cells), n_bins))
    node, off = snaps[pre + "qeff_node"], snaps[pre + "qeff_off"]
    tind, val = snaps[pre + "qeff_tind"], snaps[pre + "qeff_val"]
    for k, c in enumerate(node):
        for e in range(off[k], off[k + 1]):
            b = int(tind[e]) - cur
            if not 0 <= b < n_bins:
                raise AssertionError(f"reference stack of cell {c} holds bin {tind[e]}, current bin {cur}")
            out[slot[int(c)], b] = val[e]
    return out


def run_and_compare(stepper, snaps, n_steps, floor=1e-9, stack_check=True, dt_reff=0.5):
    """Steps `stepper` n_steps times; wherever the reference has a snapshot, compares.
    Returns a Worst() table keyed by (phase, field). With stack_check, kernel (4)'s own output is
    compared too: the pending dtReff bins of every stream cell against the reference's stacks and,
    on the steps where RetrieveQeff (tKinemat.cpp:1504-1551) leaves the current bin in place, the
    lateral inflow it returned."""
    w = Worst()
    hill = stepper.boundary == 0
    for st in range(1, n_steps + 1):
        stepper.advance()
        pre = f"s{st}_m_"
        if pre + "PotEvap" in snaps:
            names = [k[len(pre):] for k in snaps if k.startswith(pre) and k != pre + "sun"]
            got = stepper.get(names)
            for f in names:
                w.add(("m", f), snaps[pre + f], got[f], st, ET_FLOOR.get(f, floor))
        stepper.unsat()
        pre = f"s{st}_u_"
        if pre + "NwtNew" in snaps:
            got = stepper.get(STATE_FIELDS)
            for f in STATE_FIELDS:
                if pre + f in snaps:
                    w.add(("u", f), snaps[pre + f], got[f], st, floor)
        gw = stepper.gw_due()
        if gw:
            stepper.sat()
        qeff = stepper.route()
        pre = f"s{st}_r_"
        if pre + "NwtNew" in snaps:
            names = STATE_FIELDS + ["satsrf", "FlowVelocity", "TTime"] + (GW_FIELDS if gw else [])
            names = list(dict.fromkeys(names))
            got = stepper.get(names)
            for f in names:
                if pre + f not in snaps:
                    continue
                ref, g = snaps[pre + f], got[f]
                if f == "FlowVelocity":   # stream cells are rewritten by the channel router afterwards
                    ref, g = ref[hill], g[hill]
                w.add(("r", f), ref, g, st, floor)
            if stack_check and pre + "qeff_node" in snaps:
                now = stepper.time
                cur = int(math.ceil(now / dt_reff))
                ref = reference_stacks(snaps, pre, stepper.stream_cells, cur)
                w.add(("r", "QeffBins"), ref, stepper.pending(QEFF_BINS), st, 1e-12)
                if math.ceil(now / dt_reff) != math.floor(now / dt_reff):   # bin not consumed: still at the stack's front
                    w.add(("r", "Qeff"), ref[:, 0], np.asarray(qeff)[:-1], st, 1e-12)
            for nm in RESULT_ARRAYS:
                if pre + nm in snaps:
                    w.add(("res", nm), snaps[pre + nm], stepper.result(nm), st, floor)
            if pre + "globals" in snaps:
                g = stepper.globals()
                ref = snaps[pre + "globals"]
                w.add(("glob", "Stok"), ref[0:1], np.array([g["Stok"]]), st, floor)
                w.add(("glob", "TotRain"), ref[1:2], np.array([g["TotRain"]]), st, floor)
                w.add(("glob", "TotGWchange"), ref[2:3], np.array([g["TotGWchange"]]), st, 1e-6)
                w.add(("glob", "TotMoist"), ref[3:4], np.array([g["TotMoist"]]), st, 1e-6)
        stepper.balance()
        pre = f"s{st}_b_"
        if pre + "UnSaturatedStorage" in snaps:
            names = ["UnSaturatedStorage", "SaturatedStorage"] + (["CanopyStorVol"] if stepper.has_et else [])
            got = stepper.get(names)
            for f in names:
                w.add(("b", f), snaps[pre + f], got[f], st, floor)
            g, ref = stepper.globals(), snaps[pre + "BasinStorages"]
            for k, nm in ((0, "BasinRain"), (1, "BasinEvap"), (5, "BasinRunoff")):
                w.add(("b", nm), ref[k:k + 1], np.array([g[nm]]), st, 1e-6)
        stepper.reset()
    return w

