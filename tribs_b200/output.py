"""Writers of the reference's output files from structure-of-arrays state (host side).

write_dynamic_vars(): tCOutput::WriteDynamicVars (src/tInOut/tOutput.cpp:1238-1400), the spatial
`<OUTFILENAME>.HHHH_MMd` file: one comma-separated row of 52 values per active cell, in sweep
order, written between SurfaceFlow and Reset of an output step (Simulator::OutputSimulatedVars,
tSimul.cpp:266-272). The reference prints doubles through an ofstream with setprecision(p) and no
float-field flag, i.e. printf("%.{p}g"); the column without a setprecision of its own (Qpin)
inherits the precision of the column before it. The state comes from tribs_gpu_sync /
tribs_gpu_et_sync / tribs_gpu_snow_sync (names = the C ABI's field names); Qstrm and Hlev belong to
the host's channel router and are passed through."""
from __future__ import annotations

import math

import numpy as np

# (header, precision, source); the sources are resolved in _columns()
_LAND = ["CanStorParam", "IntercepCoeff", "ThroughFall", "CanFieldCap", "DrainCoeff", "DrainExpPar", "LandUseAlb",
         "VegHeight", "OptTransmCoeff", "StomRes", "VegFraction", "LeafAI"]
HEADER = ["ID", "Nwt", "Mu", "Mi", "Nf", "Nt", "Qpout", "Qpin", "Srf", "Rain", "ST", "IWE", "LWE", "SnSub", "SnEvap", "SnMelt",
          "Upack", "sLHF", "sSHF", "sGHF", "sPHF", "sRLo", "sRLi", "sRSi", "Uerr", "IntSWE", "IntSub", "IntUnl", "SoilMoist",
          "RootMoist", "CanStorage", "ActEvp", "EvpSoil", "ET", "GFlux", "HFlux", "LFlux", "Qstrm", "Hlev", "FlwVlc"] + _LAND


def _columns(basin: dict, state: dict):
    n = int(basin["n_active"][0])
    zero = np.zeros(n)

    def f(name):
        return np.asarray(state[name], float) if name in state else zero

    slope = np.arctan(np.asarray(basin["flow_slope"], float))
    cos_slope = np.maximum(np.cos(slope), 1e-9)
    varea = np.asarray(basin["varea"], float)
    cols = [(5, f("NwtNew") / cos_slope), (5, f("MuNew") / cos_slope), (5, f("MiNew") / cos_slope),
            (5, f("NfNew") / cos_slope), (5, f("NtNew") / cos_slope), (5, f("Qpout") * 1.0e-6 / varea),
            (5, f("Qpin") * 1.0e-6 / varea), (4, f("srf_hr")), (3, f("Rain")), (3, f("SnTempC")), (5, f("IceWE")),
            (5, f("LiqWE")), (7, f("SnSub")), (7, f("SnEvap")), (7, f("LiqRouted")), (5, f("Unode")), (5, f("SnLHF")),
            (5, f("SnSHF")), (5, f("SnGHF")), (5, f("SnPHF")), (5, f("SnRLout")), (5, f("SnRLin")), (5, f("SnRSin")),
            (5, f("Uerror")), (5, f("IntSWE")), (5, f("IntSub")), (5, f("IntSnUnload")), (3, f("SoilMoistureSC")),
            (3, f("RootMoistureSC")), (3, f("CanStorage")), (3, f("ActEvap")), (5, f("EvapSoil")), (5, f("EvapoTrans")),
            (3, f("GFlux")), (3, f("HFlux")), (3, f("LFlux")), (3, f("Qstrm")), (3, f("Hlevel")), (3, f("FlowVelocity"))]
    for name in _LAND:                       # per-cell land parameters: given, or zero before the first ET call
        cols.append((5, np.asarray(basin[name], float) if name in basin else f(name)))
    return cols


def dynamic_vars_name(prefix: str, time_h: float) -> str:
    hour = int(math.floor(time_h))
    minute = int(math.floor((time_h - hour) * 60))
    return f"{prefix}.{hour:04d}_{minute:02d}d"


def write_dynamic_vars(path: str, basin: dict, state: dict, header: bool = True) -> None:
    """One `.HHHH_MMd` file (not the time-0 variant, which carries two extra columns)."""
    ids = np.asarray(basin["id"])
    cols = _columns(basin, state)
    fmts = ["%d"] + [f"%.{p}g" for p, _ in cols]
    with open(path, "w") as out:
        if header:
            out.write(",".join(HEADER) + "\n")
        rows = zip(ids.tolist(), *[c.tolist() for _, c in cols])
        line = ",".join(fmts) + "\n"
        out.writelines(line % r for r in rows)
