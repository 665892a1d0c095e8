"""Experiment (not a test): how far do last-ulp differences in exp / log / pow move the results?

The device libm differs from glibc in the last ulp, and -DTRIBS_CONST_MATH replaces exp / log / pow
by const_math.cuh (<= 1 ulp from glibc, pow as exp(y ln x)). This script sweeps the *host* build of
the device headers with those routines (tests/hostcheck built with -DTRIBS_CONST_MATH
-DTRIBS_CONST_MATH_HOST) next to the bit-exact oracle on the bench basin and prints the largest
relative differences per field - an emulation, on a box without a GPU, of what the parity tests see.

  g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared -DTRIBS_CONST_MATH -DTRIBS_CONST_MATH_HOST \
      -o /tmp/libhostcheck_cm.so tests/hostcheck/hostcheck.cpp
  python tests/exp_perturb.py /tmp/libhostcheck_cm.so 300 128"""
import ctypes as C
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import storm_rain  # noqa: E402
from oracle.pyoracle import OracleModel  # noqa: E402
from tribs_b200.gpu import FIELDS  # noqa: E402
from tribs_b200.synthbasin import build_basin  # noqa: E402

PD, PI = C.POINTER(C.c_double), C.POINTER(C.c_int32)


def _ptab(arrs, T):
    return (T * len(arrs))(*[a.ctypes.data_as(T) for a in arrs])


def main(lib, side, n_steps):
    L = C.CDLL(lib)
    b = build_basin(side, side, seed=1, runtime_h=240.0)
    n = int(b["n_active"][0])
    m = OracleModel(b, storm_rain)
    alpha = np.arctan(b["flow_slope"])
    cosv = np.where(alpha > 0, np.abs(np.cos(alpha)), 1.0)
    sinv = np.where(alpha > 0, np.abs(np.sin(alpha)), 0.0)
    sf = [b["varea"], b["z"], b["bedrock"], cosv, sinv, np.cos(alpha), b["flow_vedglen"] * 1000., b["hillpath"],
          b["streampath"], b["contr_area"], b["Ks"], b["ThetaS"], b["ThetaR"], b["PoreSize"], b["AirEBubPres"],
          b["DecayF"], b["SatAnRatio"], b["UnsatAnRatio"]]
    sf = [np.ascontiguousarray(a, np.float64) for a in sf]
    si = [np.ascontiguousarray(b[k], np.int32) for k in
          ["boundary", "flow_dest", "stream_node", "spoke_rowptr", "spoke_nbr", "spoke_edge_listpos"]]
    ef = [np.ascontiguousarray(b[k], np.float64) for k in ["spoke_len", "spoke_vedglen", "spoke_compl_len", "spoke_compl_vedglen"]]
    fields = {f: (np.array(b["ttime"]) if f == "TTime" else np.array(b["init_" + f]) if "init_" + f in b else np.zeros(n))
              for f in FIELDS}
    mc = np.array([b["IntStormMAX"][0], b["surfaceSoilDepth"][0], b["rootZoneDepth"][0], b["BasArea"][0], b["BasArea_flow"][0]])
    mi = np.array([b["EToption"][0], b["Ioption"][0], b["SnOpt"][0], b["RdstrOption"][0]], np.int32)
    glob, sc = np.zeros(8), (C.c_longlong * 16)()
    L.hc_unsat.argtypes = [C.c_int, C.POINTER(PD), C.POINTER(PI), C.POINTER(PD), C.POINTER(PD), PD, PI, C.c_double,
                           C.c_double, C.c_double, C.c_int, PD, C.POINTER(C.c_longlong)]
    L.hc_sat.argtypes = [C.c_int, C.POINTER(PD), C.POINTER(PI), C.POINTER(PD), C.POINTER(PD), PD, PI, C.c_double,
                         C.c_double, C.c_double, C.c_double, PD, C.POINTER(C.c_longlong), C.POINTER(C.c_int)]
    pairs = [(o, o.replace("Old", "New")) for o in ["NwtOld", "MuOld", "MiOld", "NtOld", "NfOld", "RuOld", "RiOld"]]

    def swap():
        for o, nw in pairs:
            fields[o], fields[nw] = fields[nw], fields[o]

    watch = ["NwtNew", "MuNew", "NfNew", "NtNew", "RuNew", "Qpout", "srf", "cumsrf", "Recharge"]
    worst = {f: (0.0, 0, 0) for f in watch}
    over = {f: 0 for f in watch}
    tstep, gwstep, etistep = b["tstep"][0], b["gwstep"][0], b["etistep"][0]
    t = 0.0
    for st in range(1, n_steps + 1):
        t += tstep
        fields["Rain"][:] = storm_rain(t)
        m.advance(); m.unsat()
        L.hc_unsat(n, _ptab(sf, PD), _ptab(si, PI), _ptab(ef, PD), _ptab([fields[f] for f in FIELDS], PD), mc.ctypes.data_as(PD),
                   mi.ctypes.data_as(PI), tstep, t, float(int(t / tstep)), int(not math.fmod(t - tstep, etistep)), glob.ctypes.data_as(PD), sc)
        if m.gw_due():
            m.sat(); swap()
            ov = C.c_int(0)
            L.hc_sat(n, _ptab(sf, PD), _ptab(si, PI), _ptab(ef, PD), _ptab([fields[f] for f in FIELDS], PD), mc.ctypes.data_as(PD),
                     mi.ctypes.data_as(PI), gwstep, t, float(int(t / tstep)), glob[4], glob.ctypes.data_as(PD), sc, C.byref(ov))
        for f in watch:
            ref, got = m.state[f], fields[f]
            e = np.abs(ref - got) / np.maximum(np.abs(ref), 1e-6)
            k = int(np.argmax(e))
            over[f] += int((e > 1e-6).sum())
            if e[k] > worst[f][0]:
                worst[f] = (float(e[k]), st, k)
        m.route(); m.retrieve_qeff(); m.reset(); swap()
        for f in ("Qpin", "srf", "sbsrf", "hsrf", "psrf", "satsrf", "gwchange"):
            fields[f][:] = 0.0
    print(f"cells {n} steps {n_steps}")
    for f in watch:
        print(f"  {f:10s} worst rel {worst[f][0]:.3e} (step {worst[f][1]}, cell {worst[f][2]}); cell-steps above 1e-6: {over[f]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))
