"""TEST INFRASTRUCTURE — ctypes binding of oracle/libtribs_oracle.so (the plain-C CPU
restatement) plus a tiny driver that steps it exactly like the reference's time loop
(reference: src/tSimulator/tSimul.cpp:248-293, 386-406, 506-520).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this module.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libtribs_oracle.so")
HDR = os.path.join(HERE, "tribs_oracle.h")


def _xlist(macro: str) -> list[str]:
    src = open(HDR).read()
    m = re.search(r"#define\s+" + macro + r"\(X\)\s*\\?\n?((?:.*\\\n)*.*)\n", src)
    return re.findall(r"X\((\w+)\)", m.group(1))


STATIC_F64 = _xlist("ORC_STATIC_F64")
STATIC_I32 = _xlist("ORC_STATIC_I32")
EDGE_F64 = _xlist("ORC_EDGE_F64")
EDGE_I32 = _xlist("ORC_EDGE_I32")
STATE_F64 = _xlist("ORC_STATE_F64")
GLOBALS = _xlist("ORC_GLOBALS")
RESULTS = _xlist("ORC_RESULTS")

PD = C.POINTER(C.c_double)
PI = C.POINTER(C.c_int32)


class Basin(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("n_edges", C.c_int32),
        ("sf", PD * len(STATIC_F64)), ("si", PI * len(STATIC_I32)),
        ("nbr_rowptr", PI),
        ("ef", PD * len(EDGE_F64)), ("ei", PI * len(EDGE_I32)),
        ("BasArea", C.c_double), ("BasArea_flow", C.c_double), ("IntStormMAX", C.c_double),
        ("surfaceSoilDepth", C.c_double), ("rootZoneDepth", C.c_double),
        ("EToption", C.c_int32), ("Ioption", C.c_int32), ("SnOpt", C.c_int32), ("RdstrOption", C.c_int32),
        ("velcoef", C.c_double), ("velratio", C.c_double), ("flowexp", C.c_double),
        ("baseflow", C.c_double), ("kincoef", C.c_double), ("dtReff", C.c_double), ("hillvel0", C.c_double),
        ("outlet_node", C.c_int32), ("outlet_contr_area", C.c_double),
        ("tstep", C.c_double), ("gwstep", C.c_double), ("etistep", C.c_double), ("output_interval", C.c_double),
    ]


class State(C.Structure):
    _fields_ = [("f", PD * len(STATE_F64)), ("glob", C.c_double * len(GLOBALS))]


class Stack(C.Structure):
    _fields_ = [("n", C.c_int32), ("cap", C.c_int32), ("tind", PI), ("q", PD)]


class Results(C.Structure):
    _fields_ = [("limit", C.c_int32), ("a", PD * len(RESULTS)), ("stacks", C.POINTER(Stack))]


def build(force: bool = False) -> str:
    newest = max(os.path.getmtime(os.path.join(HERE, f)) for f in
                 ("tribs_oracle.c", "tribs_oracle.h", "tribs_oracle_et.c", "tribs_oracle_et.h"))
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        subprocess.check_call(["make", "-C", HERE, "port"], stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.orc_unsat.argtypes = [C.POINTER(Basin), C.POINTER(State), C.c_double, C.c_double]
        L.orc_reset_gw.argtypes = [C.POINTER(Basin), C.POINTER(State)]
        L.orc_sat.argtypes = [C.POINTER(Basin), C.POINTER(State), C.c_double, C.c_double]
        L.orc_route.argtypes = [C.POINTER(Basin), C.POINTER(State), C.POINTER(Results), C.c_double]
        L.orc_reset.argtypes = [C.POINTER(Basin), C.POINTER(State)]
        L.orc_balance.argtypes = [C.POINTER(Basin), C.POINTER(State), C.c_int, C.c_int, C.c_double, PD, PD, PD, PD, PD, PD]
        L.orc_retrieve_qeff.argtypes = [C.POINTER(Basin), C.POINTER(Results), C.c_double, PD]
        L.orc_state_counts.argtypes = [C.POINTER(C.c_longlong), C.c_int]
        for nm, nargs in (("lambertw", 1),):
            getattr(L, "orc_kat_" + nm).restype = C.c_double
            getattr(L, "orc_kat_" + nm).argtypes = [C.c_double] * nargs
        for nm, nargs in (("total_moist", 1), ("upper_moist", 2), ("lower_moist", 2), ("newton1", 2),
                          ("recharge", 2), ("unsat_lat", 3), ("sat_lat", 4), ("z1z2", 3)):
            fn = getattr(L, "orc_kat_" + nm)
            fn.restype = C.c_double
            fn.argtypes = [PD] + [C.c_double] * nargs
        L.orc_kat_newton2.restype = C.c_double
        L.orc_kat_newton2.argtypes = [PD, C.c_double, C.c_double, C.c_double, C.c_int]
        L.orc_sizes.restype = C.c_int
        L.orc_sizes.argtypes = [C.c_int]
        assert [L.orc_sizes(k) for k in range(7)] == [len(x) for x in (
            STATIC_F64, STATIC_I32, EDGE_F64, EDGE_I32, STATE_F64, GLOBALS, RESULTS)]
        assert L.orc_sizes(99) == C.sizeof(Basin)
        _lib = L
    return _lib


# basin.trb name -> oracle name where they differ
_BASIN_ALIAS = {
    "nbr_len": "spoke_len", "nbr_vedglen": "spoke_vedglen", "nbr_len_in": "spoke_compl_len",
    "nbr_vedglen_in": "spoke_compl_vedglen", "nbr_col": "spoke_nbr", "nbr_listpos": "spoke_edge_listpos",
}


def _pd(a):
    return a.ctypes.data_as(PD)


def _pi(a):
    return a.ctypes.data_as(PI)


class OracleModel:
    """Steps the C oracle over a flattened basin (dict as read from basin.trb)."""

    def __init__(self, basin: dict, rain_schedule=None):
        self.L = lib()
        b = basin
        self.n = n = int(b["n_active"][0])
        self.keep = {}
        B = Basin()
        B.n = n
        rp = np.ascontiguousarray(b["spoke_rowptr"], np.int32)
        B.n_edges = int(rp[-1])
        self.keep["rowptr"] = rp
        B.nbr_rowptr = _pi(rp)
        for k, nm in enumerate(STATIC_F64):
            a = np.ascontiguousarray(b[nm], np.float64); self.keep["sf" + nm] = a; B.sf[k] = _pd(a)
        for k, nm in enumerate(STATIC_I32):
            a = np.ascontiguousarray(b[nm], np.int32); self.keep["si" + nm] = a; B.si[k] = _pi(a)
        for k, nm in enumerate(EDGE_F64):
            a = np.ascontiguousarray(b[_BASIN_ALIAS[nm]], np.float64); self.keep["ef" + nm] = a; B.ef[k] = _pd(a)
        for k, nm in enumerate(EDGE_I32):
            a = np.ascontiguousarray(b[_BASIN_ALIAS[nm]], np.int32); self.keep["ei" + nm] = a; B.ei[k] = _pi(a)
        sc = lambda k: float(b[k][0])
        B.BasArea = sc("BasArea"); B.BasArea_flow = sc("BasArea_flow"); B.IntStormMAX = sc("IntStormMAX")
        B.surfaceSoilDepth = sc("surfaceSoilDepth"); B.rootZoneDepth = sc("rootZoneDepth")
        B.EToption = int(b["EToption"][0]); B.Ioption = int(b["Ioption"][0])
        B.SnOpt = int(b["SnOpt"][0]); B.RdstrOption = int(b["RdstrOption"][0])
        B.velcoef = sc("velcoef"); B.velratio = sc("velratio"); B.flowexp = sc("flowexp")
        B.baseflow = sc("baseflow"); B.kincoef = sc("kincoef"); B.dtReff = sc("dtReff"); B.hillvel0 = sc("hillvel")
        B.outlet_node = int(b["outlet_stream_node"][0]); B.outlet_contr_area = sc("outlet_contr_area")
        B.tstep = sc("tstep"); B.gwstep = sc("gwstep"); B.etistep = sc("etistep")
        B.output_interval = sc("output_interval")
        self.B = B
        self.gw_on = int(b["GW_model_label"][0]) != 0
        self.endtime = sc("endtime")

        S = State()
        self.state = {}
        for k, nm in enumerate(STATE_F64):
            src = "init_" + nm
            if nm == "TTime":
                a = np.array(b["ttime"], np.float64)
            elif src in b:
                a = np.array(b[src], np.float64)
            else:
                a = np.zeros(n)
            self.state[nm] = a
            S.f[k] = _pd(a)
        self.S = S
        self.glob = np.ctypeslib.as_array(S.glob)

        R = Results()
        R.limit = int(b["res_limit"][0])
        self.res = {}
        for k, nm in enumerate(RESULTS):
            a = np.zeros(R.limit)
            if nm == "rmin":
                a[:] = 9999.99
            self.res[nm] = a
            R.a[k] = _pd(a)
        self.stacks = (Stack * (n + 1))()
        R.stacks = C.cast(self.stacks, C.POINTER(Stack))
        self.R = R
        self.time = 0.0
        self.step_no = 0
        self.rain_schedule = rain_schedule

    # -- phases -------------------------------------------------------------------
    def advance(self):
        self.time += self.B.tstep
        self.step_no += 1
        if self.rain_schedule is not None:
            self.state["Rain"][:] = self.rain_schedule(self.time)

    def unsat(self):
        self.L.orc_unsat(C.byref(self.B), C.byref(self.S), self.B.tstep, self.time)

    def gw_due(self) -> bool:
        return self.gw_on and math.fmod(self.time, self.B.gwstep) == 0.0

    def sat(self):
        self.L.orc_reset_gw(C.byref(self.B), C.byref(self.S))
        self.L.orc_sat(C.byref(self.B), C.byref(self.S), self.B.gwstep, self.time)

    def route(self):
        self.L.orc_route(C.byref(self.B), C.byref(self.S), C.byref(self.R), self.time)

    def retrieve_qeff(self):
        out = np.zeros(self.n + 1)
        self.L.orc_retrieve_qeff(C.byref(self.B), C.byref(self.R), self.time, _pd(out))
        return out

    def state_counts(self, reset=False):
        out = (C.c_longlong * 16)()
        self.L.orc_state_counts(out, int(reset))
        return list(out)

    def reset(self):
        self.L.orc_reset(C.byref(self.B), C.byref(self.S))

    def balance(self, et=None, met_time=False, metstep_h=1.0):
        """Simulator::UpdateWaterBalance (tSimul.cpp:576-585); *et* is an OracleET or None."""
        if not hasattr(self, "basin6"):
            self.basin6 = np.zeros(6)
        gw_time = math.fmod(self.time, self.B.gwstep) == 0.0
        f = (lambda nm: _pd(et.f[nm])) if et is not None else (lambda nm: None)
        self.L.orc_balance(C.byref(self.B), C.byref(self.S), int(gw_time), int(met_time), metstep_h,
                           f("InterceptLoss"), f("EvapWetCanopy"), f("EvapoTrans"), f("CanStorage"), f("CanopyStorVol"),
                           _pd(self.basin6))

    def step(self):
        self.advance()
        self.unsat()
        if self.gw_due():
            self.sat()
        self.route()
        self.retrieve_qeff()
        self.reset()

    def stack_of(self, i):
        st = self.stacks[i]
        return ([st.tind[k] for k in range(st.n)], [st.q[k] for k in range(st.n)])


# ------------------------------------------------------------------ ET / interception sweeps
def _xlist_et() -> list[str]:
    src = open(os.path.join(HERE, "tribs_oracle_et.h")).read()
    m = re.search(r"#define\s+ORC_ET_FIELDS\(X\)\s*\\?\n?((?:.*\\\n)*.*)\n", src)
    return re.findall(r"X\((\w+)\)", m.group(1))


ET_FIELDS = _xlist_et()


def _xlist_snow() -> list[str]:
    src = open(os.path.join(HERE, "tribs_oracle_et.h")).read()
    m = re.search(r"#define\s+ORC_SNOW_FIELDS\(X\)\s*\\?\n?((?:.*\\\n)*.*)\n", src)
    return re.findall(r"X\((\w+)\)", m.group(1))


SNOW_FIELDS = _xlist_snow()


class SnowArgs(C.Structure):
    _fields_ = [("min_sn_temp", C.c_double), ("snliqfrac", C.c_double), ("hill_albedo_option", C.c_int32),
                ("hourly_time_step", C.c_int32), ("LiqWE", PD), ("IceWE", PD), ("LiqRouted", PD)]


class EtBasin(C.Structure):
    _fields_ = [("n", C.c_int32)] + [(k, PD) for k in (
        "z", "flow_slope", "aspect", "ThetaS", "ThetaR", "VolHeatCond", "SoilHeatCap", "soil_cutoff",
        "root_cutoff", "bedrock")] + [
        ("land_class", PI), ("land_table", PD),
        ("et_option", C.c_int32), ("i_option", C.c_int32), ("gflux_option", C.c_int32), ("shelter_option", C.c_int32),
        ("metstep_min", C.c_double), ("etistep_h", C.c_double), ("tlinke", C.c_double), ("temp_lapse", C.c_double),
        ("max_interstorm", C.c_double), ("hor_angle_0000", PD)]


class Met(C.Structure):
    _fields_ = [("n_records", C.c_int32), ("cell_record", PI)] + [(k, PD) for k in (
        "air_temp", "dew_temp", "rel_humid", "vap_press", "atm_press", "wind_speed", "sky_cover", "rad_global", "elev")]


class EtStep(C.Structure):
    _fields_ = [("alphaD", C.c_double), ("sinAlpha", C.c_double), ("sunaz", C.c_double), ("Io", C.c_double),
                ("hour", C.c_int32), ("first_call", C.c_int32), ("dew_hum_flag", C.c_int32), ("ts_option", C.c_int32),
                ("sky_flag_from", C.c_int32), ("te_met", C.c_double), ("te_eti", C.c_double)]


class EtShared(C.Structure):
    _fields_ = [(k, PD) for k in ("Rain", "NetPrecipitation", "EvapSoil", "EvapDryCanopy", "SoilMoisture", "RootMoisture")]


_MET_KEYS = (("air_temp", "airTemp"), ("dew_temp", "dewTemp"), ("rel_humid", "rHumidity"), ("vap_press", "vPress"),
             ("atm_press", "atmPress"), ("wind_speed", "windSpeed"), ("sky_cover", "skyCover"), ("rad_global", "radGlobal"))


class OracleET:
    """tEvapoTrans / tIntercept cell loops of the C oracle, sharing the arrays of an OracleModel."""

    def __init__(self, basin: dict, model: OracleModel, host):
        self.L = lib()
        self.L.orc_et_potential.argtypes = [C.POINTER(EtBasin), C.POINTER(EtStep), C.POINTER(Met),
                                            C.POINTER(PD), C.POINTER(EtShared)]
        self.L.orc_et_components.argtypes = [C.POINTER(EtBasin), C.POINTER(EtStep), C.POINTER(Met),
                                             C.POINTER(PD), C.POINTER(EtShared), C.c_int]
        assert self.L.orc_et_n_fields() == len(ET_FIELDS)
        b, n = basin, model.n
        self.host, self.model, self.keep = host, model, {}
        B = EtBasin()
        B.n = n
        for dst, src in (("z", "z"), ("flow_slope", "flow_slope"), ("aspect", "aspect"), ("ThetaS", "ThetaS"),
                         ("ThetaR", "ThetaR"), ("VolHeatCond", "VolHeatCond"), ("SoilHeatCap", "SoilHeatCap"),
                         ("soil_cutoff", "soil_cutoff"), ("root_cutoff", "root_cutoff"), ("bedrock", "bedrock")):
            a = np.ascontiguousarray(b[src], np.float64); self.keep[dst] = a; setattr(B, dst, _pd(a))
        lc = np.ascontiguousarray(b["land_class"], np.int32); self.keep["lc"] = lc; B.land_class = _pi(lc)
        lt = np.ascontiguousarray(b["land_table"], np.float64); self.keep["lt"] = lt; B.land_table = _pd(lt)
        B.et_option = int(b["et_option"][0]); B.i_option = int(b["et_Ioption"][0])
        B.gflux_option = int(b["et_gFluxOption"][0]); B.shelter_option = int(b["et_shelterOption"][0])
        B.metstep_min = float(b["et_timeStep"][0]); B.etistep_h = float(b["etistep"][0])
        B.tlinke = float(b["et_Tlinke"][0]); B.temp_lapse = float(b["et_tempLapseRate"][0])
        B.max_interstorm = float(b["icp_maxInterStormPeriod"][0])
        if "hor_angle_0000" in b:         # tShelter's horizon angle at azimuth 0 (OPTRADSHELT 1-3)
            ha = np.ascontiguousarray(b["hor_angle_0000"], np.float64); self.keep["ha"] = ha; B.hor_angle_0000 = _pd(ha)
        self.B = B
        self.f = {nm: np.array(b["init_" + nm], np.float64) if "init_" + nm in b else np.zeros(n) for nm in ET_FIELDS}
        self.fp = (PD * len(ET_FIELDS))(*[_pd(self.f[nm]) for nm in ET_FIELDS])
        sh = EtShared()
        for k, _ in EtShared._fields_:
            setattr(sh, k, _pd(model.state[k]))
        self.sh = sh
        self.flag = int(B.i_option != 0)
        self.snow = int(b["et_snowOption"][0]) != 0 if "et_snowOption" in b else False
        if self.snow:
            self.L.orc_snow_pack.argtypes = [C.POINTER(EtBasin), C.POINTER(SnowArgs), C.POINTER(EtStep), C.POINTER(Met),
                                             C.POINTER(PD), C.POINTER(PD), C.POINTER(EtShared), C.c_int, PD]
            assert self.L.orc_snow_n_fields() == len(SNOW_FIELDS)
            for nm in SNOW_FIELDS:
                self.f[nm] = np.array(b["init_" + nm], np.float64) if "init_" + nm in b else np.zeros(n)
            self.sfp = (PD * len(SNOW_FIELDS))(*[_pd(self.f[nm]) for nm in SNOW_FIELDS])
            self.members = np.array([0.88, 0.0, 0.0])     # albedo, Uerr, snCanWE (tSnowPack.cpp:103-150)
            a = SnowArgs()
            a.min_sn_temp = float(b["sn_minSnTemp"][0]); a.snliqfrac = float(b["sn_snliqfrac"][0])
            a.hill_albedo_option = int(b["sn_hillAlbedoOption"][0])
            a.LiqWE, a.IceWE, a.LiqRouted = _pd(model.state["LiqWE"]), _pd(model.state["IceWE"]), _pd(model.state["LiqRouted"])
            self.sn_args = a

    def _args(self, step: dict, rec: dict):
        s = EtStep(**step)
        m = Met()
        m.n_records = self.host.ns
        cr = self.host.cell_record; m.cell_record = _pi(cr)
        keep = [cr]
        for dst, src in _MET_KEYS:
            a = np.ascontiguousarray(rec[src], np.float64); keep.append(a); setattr(m, dst, _pd(a))
        el = np.ascontiguousarray(self.host.elev, np.float64); keep.append(el); m.elev = _pd(el)
        return s, m, keep

    def potential(self, step, rec):
        s, m, keep = self._args(step, rec)
        rc = self.L.orc_et_potential(C.byref(self.B), C.byref(s), C.byref(m), self.fp, C.byref(self.sh))
        assert rc == 0, "option set not covered by the oracle"

    def snow_pack(self, step, rec, hourly):
        """tSnowPack::callSnowPack (replaces both ET sweeps under OPTSNOW); hourly = the met clock of the call."""
        s, m, keep = self._args(step, rec)
        self.sn_args.hourly_time_step = int(hourly)
        rc = self.L.orc_snow_pack(C.byref(self.B), C.byref(self.sn_args), C.byref(s), C.byref(m), self.fp, self.sfp,
                                  C.byref(self.sh), self.flag, _pd(self.members))
        assert rc == 0

    def components(self, step, rec):
        s, m, keep = self._args(step, rec)
        rc = self.L.orc_et_components(C.byref(self.B), C.byref(s), C.byref(m), self.fp, C.byref(self.sh), self.flag)
        assert rc == 0


def stochastic_storm_schedule(pmean: float, stdur: float, istdur: float):
    """Rain rate seen by the cells under STOCHASTICMODE 1 (mean values, no sampling).
    One storm of PMEAN for STDUR hours; the rate is then set to 0 during the interstorm
    (tSimul.cpp:398-399) and, because tStorm::GenerateStorm leaves `p` untouched in mode 1
    (tStorm.cpp:228-242), every later "storm" rains 0 — the reference's actual behaviour.
    (reference: src/tSimulator/tSimul.cpp:386-406, tRunTimer::UpdateStorm tRunTimer.cpp:559)"""
    # stateless (several in-process ranks evaluate the same schedule, each over its own batch of times): the rate is
    # PMEAN up to the end of the first storm and 0 from the first step inside the interstorm on
    def rain(t):
        return pmean if (t <= stdur or istdur <= 0.0) else 0.0

    return rain


# ------------------------------------------------------------------ channel router (tribs_oracle_channel.c)
class _Channel(C.Structure):
    _fields_ = [("n_cells", C.c_int), ("n_reach", C.c_int), ("off", PI), ("node", PI), ("width", PD), ("length", PD),
                ("slope", PD), ("roughness", C.c_double), ("uniform_width", C.c_double)]


class _ChannelState(C.Structure):
    _fields_ = [("hlevel", PD), ("qstrm", PD), ("velocity", PD), ("outlet_hlev", PD)]


from tribs_b200.channel import channel_network  # noqa: E402  (host-side flattening of tKinemat's reach lists)


class OracleChannel:
    """One tKinemat::RunRoutingModel call per step (without its hillslope half, which OracleModel.route is)."""

    def __init__(self, basin: dict, net: dict | None = None):
        self.L = lib()
        self.L.orc_channel_step.argtypes = [C.POINTER(_Channel), C.POINTER(_ChannelState), PD, C.c_double, PD,
                                            C.POINTER(C.c_int)]
        self.L.orc_channel_step.restype = C.c_int
        assert int(basin["chan_percolation"][0]) == 0 and int(basin["chan_optres"][0]) == 0
        self.net = net = channel_network(basin) if net is None else net
        n = net["n_cells"]
        self.n_reach = len(net["off"]) - 1
        self.hlevel, self.qstrm, self.velocity = np.zeros(n + 1), np.zeros(n + 1), np.zeros(n + 1)
        self.outlet_hlev = np.zeros(self.n_reach)
        for nm, key in (("hlevel", "init_Hlevel"), ("qstrm", "init_Qstrm"), ("velocity", "init_FlowVelocity")):
            if key in basin:
                getattr(self, nm)[:n] = np.asarray(basin[key], float)[:n]
        self.C = _Channel(n, self.n_reach, _pi(net["off"]), _pi(net["node"]), _pd(net["width"]), _pd(net["length"]),
                          _pd(net["slope"]), net["roughness"], net["uniform_width"])
        self.S = _ChannelState(_pd(self.hlevel), _pd(self.qstrm), _pd(self.velocity), _pd(self.outlet_hlev))
        self.dt = float(basin["tstep"][0]) * 3600.0
        self.iters = 0

    def step(self, qeff) -> float:
        """qeff: what RetrieveQeff returned for this step, per cell index (+ the basin outlet slot)."""
        qeff = np.ascontiguousarray(qeff, float)
        assert qeff.size == self.net["n_cells"] + 1
        qout, it = C.c_double(0.0), C.c_int(0)
        rc = self.L.orc_channel_step(C.byref(self.C), C.byref(self.S), _pd(qeff), self.dt, C.byref(qout), C.byref(it))
        if rc:
            raise RuntimeError(f"orc_channel_step: reach solver failed ({rc})")
        self.iters += it.value
        return qout.value
