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
