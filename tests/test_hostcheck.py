"""CPU test: the device physics headers (tribs_b200/csrc/physics.cuh, cell_update.cuh) compiled
with g++ must agree bit for bit with the oracle when swept sequentially. This is what lets the
CUDA kernels be debugged in a container without a GPU; nothing here is shipped."""
import ctypes as C
import math
import os

import numpy as np
import pytest

from refrun import ROOT, SCENARIOS, load_golden, storm_of

PD = C.POINTER(C.c_double)
PI = C.POINTER(C.c_int32)


def _ptab(arrs, T):
    return (T * len(arrs))(*[a.ctypes.data_as(T) for a in arrs])


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_device_headers_match_oracle_on_host(name, built):
    from oracle.pyoracle import OracleModel, stochastic_storm_schedule
    from tribs_b200.gpu import FIELDS
    L = C.CDLL(os.path.join(ROOT, "tests", "hostcheck", "libhostcheck.so"))
    assert L.hc_n_fields() == len(FIELDS)
    b, snaps, _ = load_golden(name)
    n = int(b["n_active"][0])
    m = OracleModel(b, stochastic_storm_schedule(*storm_of(b)))
    rain = stochastic_storm_schedule(*storm_of(b))
    alpha = [math.atan(x) for x in b["flow_slope"]]
    cosv = np.array([abs(math.cos(a)) if a > 0 else 1.0 for a in alpha])
    sinv = np.array([abs(math.sin(a)) if a > 0 else 0.0 for a in alpha])
    cos_gw = np.array([math.cos(a) for a in alpha])
    sf = [b["varea"], b["z"], b["bedrock"], cosv, sinv, cos_gw, b["flow_vedglen"] * 1000., b["hillpath"],
          b["streampath"], b["contr_area"], b["Ks"], b["ThetaS"], b["ThetaR"], b["PoreSize"], b["AirEBubPres"],
          b["DecayF"], b["SatAnRatio"], b["UnsatAnRatio"]]
    sf = [np.ascontiguousarray(a, np.float64) for a in sf]
    si = [np.ascontiguousarray(b[k], np.int32) for k in
          ["boundary", "flow_dest", "stream_node", "spoke_rowptr", "spoke_nbr", "spoke_edge_listpos"]]
    ef = [np.ascontiguousarray(b[k], np.float64) for k in
          ["spoke_len", "spoke_vedglen", "spoke_compl_len", "spoke_compl_vedglen"]]
    fields = {}
    for f in FIELDS:
        if f == "TTime":
            fields[f] = np.array(b["ttime"])
        elif "init_" + f in b:
            fields[f] = np.array(b["init_" + f])
        else:
            fields[f] = np.zeros(n)
    mc = np.array([b["IntStormMAX"][0], b["surfaceSoilDepth"][0], b["rootZoneDepth"][0], b["BasArea"][0],
                   b["BasArea_flow"][0]])
    mi = np.array([b["EToption"][0], b["Ioption"][0], b["SnOpt"][0], b["RdstrOption"][0]], np.int32)
    glob = np.zeros(8)
    sc = (C.c_longlong * 16)()
    L.hc_unsat.argtypes = [C.c_int, C.POINTER(PD), C.POINTER(PI), C.POINTER(PD), C.POINTER(PD), PD, PI, C.c_double,
                           C.c_double, C.c_double, C.c_int, PD, C.POINTER(C.c_longlong)]
    L.hc_sat.argtypes = [C.c_int, C.POINTER(PD), C.POINTER(PI), C.POINTER(PD), C.POINTER(PD), PD, PI, C.c_double,
                         C.c_double, C.c_double, C.c_double, PD, C.POINTER(C.c_longlong), C.POINTER(C.c_int)]
    pairs = [(o, o.replace("Old", "New")) for o in ["NwtOld", "MuOld", "MiOld", "NtOld", "NfOld", "RuOld", "RiOld"]]

    def swap():
        for o, nw in pairs:
            fields[o], fields[nw] = fields[nw], fields[o]

    skip = ("FlowVelocity", "TTime", "Qstrm", "LiqWE", "IceWE", "LiqRouted")
    tstep, gwstep, etistep = b["tstep"][0], b["gwstep"][0], b["etistep"][0]
    t = 0.0
    for st in range(1, int(snaps["n_steps"][0]) + 1):
        t += tstep
        fields["Rain"][:] = rain(t)
        m.advance()
        m.unsat()
        ft = _ptab([fields[f] for f in FIELDS], PD)
        L.hc_unsat(n, _ptab(sf, PD), _ptab(si, PI), _ptab(ef, PD), ft, mc.ctypes.data_as(PD), mi.ctypes.data_as(PI),
                   tstep, t, float(int(t / tstep)), int(not math.fmod(t - tstep, etistep)), glob.ctypes.data_as(PD), sc)
        if m.gw_due():
            m.sat()
            swap()
            ov = C.c_int(0)
            ft = _ptab([fields[f] for f in FIELDS], PD)
            L.hc_sat(n, _ptab(sf, PD), _ptab(si, PI), _ptab(ef, PD), ft, mc.ctypes.data_as(PD),
                     mi.ctypes.data_as(PI), gwstep, t, float(int(t / tstep)), glob[4], glob.ctypes.data_as(PD), sc,
                     C.byref(ov))
            assert ov.value == 0
        for f in FIELDS:
            if f not in skip:
                assert np.array_equal(fields[f], m.state[f]), (st, f)
        m.route(); m.retrieve_qeff(); m.reset()
        swap()
        for f in ("Qpin", "srf", "sbsrf", "hsrf", "psrf", "satsrf", "gwchange"):
            fields[f][:] = 0.0
    assert np.array_equal(glob[:4], m.glob[:4])


# ------------------------------------------------------------------ kernel (1) / snow device headers
def _hc_lib():
    L = C.CDLL(os.path.join(ROOT, "tests", "hostcheck", "libhostcheck.so"))
    L.hc_et_step.argtypes = [C.c_int, PI, PD, C.POINTER(PD), PI, PD, C.c_int, PI, PD, PI, C.POINTER(PD), C.POINTER(PD),
                             C.POINTER(PD), C.POINTER(PD), PD, C.POINTER(PD)]
    return L


class _EtHostCheck:
    """Runs et_update.cuh / snow_update.cuh (compiled for the host) on a copy of the state the
    oracle is about to update and compares what both leave behind."""

    MAIN = ("Rain", "NetPrecipitation", "EvapSoil", "EvapDryCanopy", "SoilMoisture", "RootMoisture", "LiqWE", "IceWE",
            "LiqRouted")

    def __init__(self, basin, stepper, tol):
        from oracle.pyoracle import ET_FIELDS, SNOW_FIELDS
        from tribs_b200.gpu import ET_FIELDS as GET, FIELDS, SNOW_FIELDS as GSN
        assert ET_FIELDS == GET and SNOW_FIELDS == GSN
        self.L, self.b, self.o, self.tol = _hc_lib(), basin, stepper, tol
        self.et_names, self.sn_names, self.fields = ET_FIELDS, SNOW_FIELDS, FIELDS
        assert self.L.hc_n_et_fields() == len(ET_FIELDS) and self.L.hc_n_snow_fields() == len(SNOW_FIELDS)
        self.worst, self.calls, self.rounds = 0.0, 0, []
        b = basin
        self.n = int(b["n_active"][0])
        self.ci = np.array([b["et_option"][0], b["et_Ioption"][0], b["et_gFluxOption"][0], b["et_shelterOption"][0]], np.int32)
        self.cd = np.array([b["et_timeStep"][0], b["etistep"][0], b["et_Tlinke"][0], b["et_tempLapseRate"][0],
                            b["icp_maxInterStormPeriod"][0]], np.float64)
        self.est = [np.ascontiguousarray(b[k], np.float64) for k in
                    ("z", "flow_slope", "aspect", "VolHeatCond", "SoilHeatCap", "soil_cutoff", "root_cutoff", "ThetaR",
                     "bedrock", "varea")]
        self.est.append(np.ascontiguousarray(b["hor_angle_0000"], np.float64) if "hor_angle_0000" in b else np.zeros(self.n))
        self.lc = np.ascontiguousarray(b["land_class"], np.int32)
        self.lt = np.ascontiguousarray(b["land_table"], np.float64)

    def snapshot(self):
        et = self.o.et
        self.snap_et = {k: et.f[k].copy() for k in self.et_names}
        self.snap_sn = {k: et.f[k].copy() for k in self.sn_names} if et.snow else {}
        self.snap_main = {k: self.o.m.state[k].copy() for k in self.MAIN}
        self.snap_members = et.members.copy() if et.snow else None

    def run_and_compare(self, pot, cmp_, snow_hourly=None):
        host = self.o.host
        step = (pot or cmp_)[0]
        si = np.array([pot is not None, cmp_ is not None, int(self.ci[1] != 0), step["hour"],
                       pot[0]["first_call"] if pot else 0, step["dew_hum_flag"], step["ts_option"],
                       (pot or cmp_)[0]["sky_flag_from"], pot is not None], np.int32)
        sd = np.array([step["alphaD"], step["sinAlpha"], step["sunaz"], step["Io"], pot[0]["te_met"] if pot else 0.0,
                       (cmp_ or pot)[0]["te_eti"]], np.float64)
        keys = ("airTemp", "dewTemp", "rHumidity", "vPress", "atmPress", "windSpeed", "skyCover", "radGlobal")

        def met(rec):
            arrs = [np.ascontiguousarray(rec[k], np.float64) for k in keys] + [np.ascontiguousarray(host.elev, np.float64)]
            return arrs, _ptab(arrs, PD)

        keep_p, mp = met((pot or cmp_)[1])
        keep_c, mc = met((cmp_ or pot)[1])
        etf = [self.snap_et[k] for k in self.et_names]
        fields = [self.snap_main[k] if k in self.snap_main else np.zeros(self.n) for k in self.fields]
        snow, snf_tab, snf = None, None, []
        if snow_hourly is not None:
            snow = np.array([self.b["sn_minSnTemp"][0], self.b["sn_snliqfrac"][0], self.b["sn_hillAlbedoOption"][0],
                             snow_hourly, self.snap_members[0], self.snap_members[2], 0.0], np.float64)
            snf = [self.snap_sn[k] for k in self.sn_names]
            snf_tab = _ptab(snf, PD)
        self.L.hc_et_step(self.n, self.ci.ctypes.data_as(PI), self.cd.ctypes.data_as(PD), _ptab(self.est, PD),
                          self.lc.ctypes.data_as(PI), self.lt.ctypes.data_as(PD), self.lt.size // 15,
                          si.ctypes.data_as(PI), sd.ctypes.data_as(PD), host.cell_record.ctypes.data_as(PI), mp, mc,
                          _ptab(etf, PD), _ptab(fields, PD), snow.ctypes.data_as(PD) if snow is not None else None, snf_tab)
        et = self.o.et
        floor = {"SurfTemp": 1.0, "SoilTemp": 1.0, "AirTemp": 1.0, "DewTemp": 1.0, "NetRad": 1.0, "GFlux": 1.0, "HFlux": 1.0,
                 "LFlux": 1.0, "SnTempC": 1.0, "Uerror": 1e9, "CumUerror": 1e9, "DU": 1e-3, "Unode": 1e-3}
        for names, got_of in ((self.et_names, lambda k: self.snap_et[k]), (self.sn_names if snow is not None else [], lambda k: self.snap_sn[k]),
                              (self.MAIN, lambda k: fields[self.fields.index(k)])):
            for k in names:
                if k == "CanopyStorVol":   # the device applies CanopyBalance inside this sweep, the oracle at the end of the step
                    continue
                ref = et.f[k] if k in et.f else self.o.m.state[k]
                err = np.abs(ref - got_of(k)) / np.maximum(np.abs(ref), floor.get(k, 1e-9))
                assert err.max() <= self.tol, (k, self.calls, float(err.max()), ref[np.argmax(err)], got_of(k)[np.argmax(err)])
                self.worst = max(self.worst, float(err.max()))
        if snow is not None and self.ci[1] != 0:     # canopy: the members the call leaves behind, and the probe rounds
            assert snow[4] == et.members[0] or abs(snow[4] - et.members[0]) <= self.tol, (snow[4], et.members[0])
            assert abs(snow[5] - et.members[2]) <= self.tol * max(1.0, abs(et.members[2])), (snow[5], et.members[2])
            self.rounds.append(int(snow[6]))
        self.calls += 1


def _drive_with_host_check(basin, n_steps, tol):
    from common import OracleStepper
    from refrun import schedule_for
    o = OracleStepper(basin, schedule_for(basin))
    hc = _EtHostCheck(basin, o, tol)
    for _ in range(n_steps):
        o.m.advance()
        o.cal.advance(o.m.time, o.m.B.tstep)
        met_due, eti_due = o.clock.due(o.m.time)
        o._met_due = met_due
        if o.et.snow:
            if met_due:
                hourly = o.host.hourlyTimeStep
                step, rec = o.host.potential_call(o.cal.hour, o.m.time)
                step["te_eti"] = float(int(o.m.time / o.host.etistep))
                hc.snapshot()
                o.et.snow_pack(step, rec, hourly)
                hc.run_and_compare((step, rec), (step, rec), snow_hourly=hourly)
        elif met_due or eti_due:
            pot = o.host.potential_call(o.cal.hour, o.m.time) if met_due else None
            cmp_ = o.host.components_call(o.m.time) if eti_due else None
            hc.snapshot()
            if pot:
                o.et.potential(*pot)
            if cmp_:
                o.et.components(*cmp_)
            hc.run_and_compare(pot, cmp_)
        o.unsat()
        if o.gw_due():
            o.sat()
        o.route(); o.balance(); o.reset()
    return hc


def test_et_device_header_matches_oracle_on_host(built):
    """et_update.cuh (the source of et_cells_kernel) compiled with g++: every call of the fused
    potential-ET + ET-partition + Rutter sweep against the bit-exact oracle on the same inputs."""
    basin, snaps, _ = load_golden("met")
    hc = _drive_with_host_check(basin, int(snaps["n_steps"][0]), tol=1e-9)
    assert hc.calls >= 40 and hc.worst < 1e-9


@pytest.mark.parametrize("variant", ["deardorff", "priestley-taylor", "gray", "gradient-gflux", "sky-cover-nodata"])
def test_et_device_header_option_variants_on_host(variant, built):
    from refrun import have_harness, run_reference
    from test_oracle import et_variant_spec
    if not have_harness():
        pytest.skip("oracle/_ref/tribs_ref_harness not built")
    basin, _, _, _ = run_reference(et_variant_spec(variant, 10, 30.0, 4))
    hc = _drive_with_host_check(basin, 30 * 16, tol=1e-9)
    assert hc.calls >= 25


@pytest.mark.parametrize("extra", [{"OPTINTERCEPT": 0}, {"OPTINTERCEPT": 0, "HILLALBOPT": 2}, {"OPTINTERCEPT": 0, "HILLALBOPT": 1}])
def test_snow_device_header_matches_oracle_on_host(extra, built):
    from refrun import have_harness, run_reference, spec_for
    if not have_harness():
        pytest.skip("oracle/_ref/tribs_ref_harness not built")
    basin, _, _, _ = run_reference(spec_for("snow", 10, runtime_h=96.0, seed=9, pmean=8.0, extra=extra))
    hc = _drive_with_host_check(basin, 96 * 16, tol=1e-9)
    assert hc.calls >= 90
    assert np.abs(hc.o.et.f["CumMelt"]).max() > 0


@pytest.mark.parametrize("extra", [{}, {"HILLALBOPT": 2}, {"OPTINTERCEPT": 1}])
def test_snow_with_canopy_device_header_matches_oracle_on_host(extra, built):
    """OPTSNOW 1 with a canopy: the reference carries `albedo` / `snCanWE` from cell to cell in sweep
    order. The host check runs the same probe / last-writer scan / commit protocol as
    snow_step_with_canopy (tribs_gpu.cu) on the device's cell update and must land on what the
    sequential oracle (bit-exact against the reference) leaves behind."""
    from refrun import have_harness, run_reference, spec_for
    if not have_harness():
        pytest.skip("oracle/_ref/tribs_ref_harness not built")
    basin, _, _, _ = run_reference(spec_for("snow", 10, runtime_h=96.0, seed=9, pmean=8.0, extra=extra))
    assert int(basin["et_Ioption"][0]) != 0
    hc = _drive_with_host_check(basin, 96 * 16, tol=1e-9)
    assert hc.calls >= 90
    assert np.abs(hc.o.et.f["CumIntSub"]).max() > 0 and np.abs(hc.o.et.f["CumMelt"]).max() > 0
    assert max(hc.rounds) >= 2 and max(hc.rounds) <= 4, hc.rounds


@pytest.mark.parametrize("name,option", [("met", 1), ("met", 2), ("met", 3), ("snow", 1), ("snow", 2), ("snow", 3)])
def test_radiation_sheltering_device_header_matches_oracle_on_host(name, option, built):
    """OPTRADSHELT 1-3 in et_update.cuh / snow_update.cuh (tShelter's sky-view factor from the node, the
    horizon angle at azimuth 0, land-view reflection) against the oracle, which is bit-exact against the
    reference for these options (tests/test_oracle.py)."""
    from refrun import have_harness, run_reference, spec_for
    if not have_harness():
        pytest.skip("oracle/_ref/tribs_ref_harness not built")
    basin, _, _, _ = run_reference(spec_for(name, 10, runtime_h=40.0, seed=5, dem_ridge=300.0, extra={"OPTRADSHELT": option}))
    assert int(basin["et_shelterOption"][0]) == option and "hor_angle_0000" in basin
    hc = _drive_with_host_check(basin, 40 * 16, tol=1e-9)
    assert hc.calls >= 35
