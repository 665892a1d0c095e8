"""Host side of kernel (1): what tEvapoTrans keeps per *call* rather than per cell.

The per-cell loops of ``tEvapoTrans::callEvapoPotential`` / ``callEvapoTrans`` and
``tIntercept::callInterception`` run on the device (``tribs_gpu_et_step``). What stays on the
host is what the reference computes once per call and what it reads from files:

* the run calendar (``tRunTimer::correctCalendarTime``, src/tSimulator/tRunTimer.cpp:316-372),
* the met clock ``hourlyTimeStep`` and ``setTime`` (src/tHydro/tEvapoTrans.cpp:472-503),
* sun geometry ``SetSunVariables`` / ``ComputeHourAngle`` / ``julianDay`` (:1523-1700),
* the weather-station records of the hour (``newHydroMetData``, :3962-4010) including the
  one-off decisions taken at record 0 (site latitude/longitude/GMT, which humidity variable is
  observed, :3989-4003) and the sticky no-data sky-cover switch (:724-727).

Everything here is O(stations) per call; nothing touches per-cell data.
"""
from __future__ import annotations

import math

import numpy as np

NODATA = 9999.99


class Calendar:
    """year/month/day/hour/minute of tRunTimer, advanced like ``tRunTimer::Advance``."""

    DAYS = [0, 31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]

    def __init__(self, start):
        self.year, self.month, self.day, self.hour, self.minute = (int(v) for v in start)

    def advance(self, current_time: float, dt: float):
        minutet = current_time - math.floor(current_time)
        if minutet == 0 and dt < 1:
            self.minute, cnt = 0, 1
        elif minutet == 0 and dt >= 1:
            self.minute, cnt = 0, int(math.ceil(dt))
        elif minutet != 0 and dt < 1:
            self.minute, cnt = int(minutet * 60), 0
        else:
            self.minute, cnt = int((dt - math.floor(dt)) * 60), int(math.floor(dt))
        for _ in range(cnt):
            self.hour += 1
            if self.hour == 24:
                self.day += 1
                self.hour = 0
            if self.month == 2 and self.day > self.DAYS[2]:
                if self.hour == 0 and ((self.year % 4) != 0 or self.day > 29):
                    self.month += 1
                    self.day = 1
            elif self.day > self.DAYS[self.month]:
                self.month += 1
                self.day = 1
                if self.month > 12:
                    self.year += 1
                    self.month = 1


def julian_day(year, month, day):
    dim = [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]
    if (year % 4 == 0 and year % 100 != 0) or (year % 400 == 0):
        dim[1] = 29
    return sum(dim[:month - 1]) + day


def hour_angle(tt, delta):
    if tt < 12.0 + delta:
        dtau = (tt + 12.0 - delta) * 15.0
    elif abs(tt - 12.0 - delta) < 1.0e-3:
        dtau = (tt + 12.0 - delta) * 15.0 if delta > 0.0 else (tt - 12.0 - delta) * 15.0
    else:
        dtau = (tt - 12.0 - delta) * 15.0
    return dtau * (4.0 * math.atan(1.0)) / 180.0


def sun_variables(latitude, longitude, gmt, node_hour, cal_ymd, etistep_h):
    """tEvapoTrans::SetSunVariables; returns alphaD, sinAlpha, sunaz, Io."""
    pi = 4.0 * math.atan(1.0)
    dgmt = float(gmt)
    long_sm = 15.0 * abs(int(gmt))
    long_m = abs(longitude)
    ts = float(node_hour)
    jday = julian_day(*cal_ymd)
    r = 1.0 + 0.017 * math.cos((2.0 * pi / 365.0) * (186.0 - jday))
    dele = (23.45 * pi / 180.0) * math.cos((2.0 * pi / 365.0) * (172.0 - jday))
    phi = latitude * pi / 180.0
    if gmt:
        delta_t = (1.0 / 15.0) * (long_sm - long_m) * (dgmt / abs(int(gmt)))
    else:
        delta_t = -long_m * 15.0
    tau1 = hour_angle(ts + 1.0, delta_t)
    tsp1 = ts + etistep_h if ts < 23.0 else 0.0
    tau2 = hour_angle(tsp1 + 1.0, delta_t)
    if abs(tau1 - tau2) >= pi:
        if tau2 < tau1:
            tau2 += 2.0 * pi
        else:
            tau1 += 2.0 * pi
    tau = (tau1 + tau2) / 2.0
    if tau > 2.0 * pi:
        tau -= 2.0 * pi
    sin_alpha = math.sin(dele) * math.sin(phi) + math.cos(dele) * math.cos(phi) * math.cos(tau)
    alpha_d = math.asin(sin_alpha) * 180.0 / pi
    den = math.tan(dele) * math.cos(phi) - math.sin(phi) * math.cos(tau)
    num = -math.sin(tau)
    # C: atan(x/0.0) = atan(+-inf); Python raises on the division
    sunaz = math.atan(num / den) if den != 0.0 else math.atan(math.copysign(math.inf, num) if num != 0.0 else math.nan)
    if 0.0 < tau <= pi:
        sunaz += pi if sunaz > 0.0 else 2.0 * pi
    elif pi <= tau <= 2.0 * pi:
        if sunaz < 0.0:
            sunaz += pi
    io = 1367.0 / (r * r)      # the compiler turns the reference's pow(r, 2.0) into r*r
    return alpha_d, sin_alpha, sunaz, io


class EvapoTransHost:
    """Per-call state of tEvapoTrans with station data (METDATAOPTION 1)."""

    SERIES = ("airTemp", "dewTemp", "rHumidity", "vPress", "atmPress", "windSpeed", "skyCover", "radGlobal")

    def __init__(self, basin):
        g = lambda k: np.asarray(basin[k])
        self.n = int(g("n_active")[0])
        self.ns = int(g("met_n_stations")[0])
        self.nt = int(g("met_n_times")[0])
        self.series = {k: g("met_" + k).reshape(self.ns, self.nt) for k in self.SERIES}
        self.station_id = g("met_station_id")
        self.elev = g("met_other").astype(np.float64)
        self.lat, self.long, self.gmt_st = g("met_lat"), g("met_long"), g("met_gmt")
        self.calendar = g("met_calendar").reshape(self.nt, 4)
        self.ts_option = int(g("met_tsOption")[0])
        assigned = g("met_assigned_station")
        index_of = {int(s): k for k, s in enumerate(self.station_id)}
        self.cell_record = np.array([index_of[int(a)] for a in assigned], dtype=np.int32)
        self.etistep = float(g("etistep")[0])
        self.metstep = float(g("metstep")[0])
        # members of tEvapoTrans::initializeVariables (tEvapoTrans.cpp:381-431)
        self.hourlyTimeStep = 0
        self.latitude = self.longitude = 0.0
        self.gmt = 0
        self.dewHumFlag = 0
        self.skycover_flag = 0
        self.sun = (0.0, 0.0, 0.0, 0.0)
        self.hour = 0

    def record(self, t):
        return {k: np.ascontiguousarray(v[:, t]) for k, v in self.series.items()}

    def SetEnvironment(self, timer_hour):
        """setTime + SetSunVariables for the call (tEvapoTrans.cpp:517-524)."""
        t = self.hourlyTimeStep
        y, mo, d, h = (int(v) for v in self.calendar[t])
        self.hour = h
        self.sun = sun_variables(self.latitude, self.longitude, self.gmt, timer_hour, (y, mo, d), self.etistep)

    def _first_record_decisions(self, rec):
        """What newHydroMetData does at record 0 for every cell; the last cell's station wins
        for the site members (tEvapoTrans.cpp:3989-4003)."""
        last = int(self.cell_record[-1])
        self.latitude, self.longitude, self.gmt = float(self.lat[last]), float(self.long[last]), int(self.gmt_st[last])
        nd = lambda v: abs(v - NODATA) < 1.0e-3
        dew, vp, rh = rec["dewTemp"][last], rec["vPress"][last], rec["rHumidity"][last]
        if nd(dew) and nd(vp):
            self.dewHumFlag = 0
        elif nd(rh) and nd(vp):
            self.dewHumFlag = 1
        elif nd(rh) and nd(dew):
            self.dewHumFlag = 2

    def potential_call(self, timer_hour, current_time):
        """Arguments of the device sweep that replaces callEvapoPotential's cell loop."""
        self.SetEnvironment(timer_hour)
        t = self.hourlyTimeStep
        rec = self.record(t)
        first = (t == 0)
        if first:
            self._first_record_decisions(rec)
        # sticky switch: the first cell (in sweep order) whose station reports no sky cover turns
        # the computed sky cover on for itself and everything after it, for the rest of the run
        if self.skycover_flag:
            sky_from = 0
        else:
            # per station, not per cell: the first cell (sweep order) of every station is known once and for all
            if not hasattr(self, "_first_cell_of_record"):
                first_cell = np.full(len(self.elev), self.n, np.int64)
                np.minimum.at(first_cell, self.cell_record, np.arange(self.n))
                self._first_cell_of_record = first_cell
            bad = np.abs(rec["skyCover"] - NODATA) < 1.0e-3
            sky_from = int(self._first_cell_of_record[bad].min()) if bad.any() else self.n
            if sky_from < self.n:
                self.skycover_flag = 1
        step = dict(alphaD=self.sun[0], sinAlpha=self.sun[1], sunaz=self.sun[2], Io=self.sun[3], hour=self.hour,
                    first_call=int(first), dew_hum_flag=self.dewHumFlag, ts_option=self.ts_option,
                    sky_flag_from=sky_from, te_met=float(int(current_time / self.metstep)), te_eti=0.0)
        self.hourlyTimeStep += 1
        return step, rec

    def components_call(self, current_time):
        """Arguments for callEvapoTrans: the record at the (already advanced) met clock."""
        rec = self.record(min(self.hourlyTimeStep, self.nt - 1))
        step = dict(alphaD=self.sun[0], sinAlpha=self.sun[1], sunaz=self.sun[2], Io=self.sun[3], hour=self.hour,
                    first_call=0, dew_hum_flag=self.dewHumFlag, ts_option=self.ts_option,
                    sky_flag_from=0 if self.skycover_flag else self.n,
                    te_met=0.0, te_eti=float(int(current_time / self.etistep)))
        return step, rec


class MetClock:
    """Simulator::get_next_met (tSimul.cpp:666-678): when the MET and ET/I sweeps are due."""

    def __init__(self, tstep, metstep, etistep):
        self.met_hour = self.eti_hour = tstep
        self.metstep, self.etistep = metstep, etistep

    def due(self, current_time):
        if self.met_hour < current_time:
            self.met_hour += self.metstep
        if self.eti_hour < current_time:
            self.eti_hour += self.etistep
        return current_time == self.met_hour, current_time == self.eti_hour


class GaugeRain:
    """tRainfall with RAINSOURCE 3 (src/tRasTin/tRainfall.cpp:434-503): at every RAININTRVL the
    hour's gauge values are spread over the cells by the Thiessen assignment."""

    def __init__(self, basin):
        g = lambda k: np.asarray(basin[k])
        self.n = int(g("n_active")[0])
        ng, self.nt = len(g("rain_gauge_id")), int(g("rain_n_times")[0])
        self.series = g("rain_series").reshape(ng, self.nt)
        index_of = {int(s): k for k, s in enumerate(g("rain_gauge_id"))}
        self.cell_gauge = np.array([index_of[int(a)] for a in g("rain_assigned_gauge")], dtype=np.int64)
        self.elev = g("rain_gauge_elev").astype(np.float64)
        self.lapse = float(g("rain_precLapseRate")[0])
        self.dt = float(g("rain_dt")[0])
        self.tstep = float(g("tstep")[0])
        self.z = g("z").astype(np.float64)
        self.hourlyTimeStep = 0
        self.current = np.zeros(self.n)
        self.current_gauges = np.zeros(ng)

    def is_gauge_time(self, current_time):
        """tRunTimer::isGaugeTime (tRunTimer.cpp:270-276)."""
        q = (current_time - self.tstep) / self.dt
        return q == math.floor(q)

    def gauge_values(self, current_time):
        """The same forcing per GAUGE (for tribs_gpu_set_rain_gauges(self.cell_gauge) + rain_mode GAUGES): what the
        cells of each gauge's Thiessen polygon receive. Only without a precipitation lapse rate, which makes the
        value depend on the cell's elevation."""
        if self.lapse != 0.0:
            raise ValueError("PRECLAPSE != 0: rain differs from cell to cell inside a Thiessen polygon; use the per-cell form")
        if self.is_gauge_time(current_time):
            t = min(self.hourlyTimeStep, self.nt - 1)
            r = self.series[:, t]
            self.current_gauges = np.where(r >= 1e-5, r, 0.0)
            self.hourlyTimeStep += 1
        return self.current_gauges

    def __call__(self, current_time):
        """Per-cell rain (mm/h) valid at *current_time*."""
        if self.is_gauge_time(current_time):
            t = min(self.hourlyTimeStep, self.nt - 1)
            r = self.series[:, t][self.cell_gauge]
            adj = r + self.lapse * (self.z - self.elev[self.cell_gauge])
            self.current = np.where((r >= 1e-5) & (adj >= 1e-5), adj, 0.0)
            self.hourlyTimeStep += 1
        return self.current
