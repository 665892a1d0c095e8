"""Writers of the reference's output files from structure-of-arrays state (host side).

write_dynamic_vars(): tCOutput::WriteDynamicVars (src/tInOut/tOutput.cpp:1238-1400), the spatial
`<OUTFILENAME>.HHHH_MMd` file: one comma-separated row of 52 values per active cell, in sweep
order, written between SurfaceFlow and Reset of an output step (Simulator::OutputSimulatedVars,
tSimul.cpp:266-272). The reference prints doubles through an ofstream with setprecision(p) and no
float-field flag, i.e. printf("%.{p}g"); the column without a setprecision of its own (Qpin)
inherits the precision of the column before it. The state comes from tribs_gpu_sync /
tribs_gpu_et_sync / tribs_gpu_snow_sync (names = the C ABI's field names); Qstrm and Hlev belong to
the host's channel router and are passed through.

PixelWriter: tCOutput::WritePixelInfo (tOutput.cpp:1080-1230), the `<OUTFILENAME><id>.pixel` time
series of one cell: one row of 82 values per output interval, fixed notation; the stream keeps its
precision from row to row, so the third column has 6 decimals in the first row and 7 afterwards."""
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


PIXEL_HEADER = ("NodeID Time_hr Nwt_mm Nf_mm Nt_mm Mu_mm Mi_mm QpOut_mm_h QpIn_mm_h Trnsm_m2_h GWflx_m3_h Srf_mm Rain_mm_h "
                "SoilMoist_[] RootMoist_[] AirT_oC DewT_oC SurfT_oC SoilT_oC Press_Pa RelHum_[] SkyCov_[] Wind_m_s NetRad_W_m2 "
                "ShrtRadIn_W_m2 ShortRadInSlope_W_m2 ShrtRadIn_dir_W_m2 ShrtRadIn_dif_W_m2 ShortAbsbVeg_W_m2 ShortAbsbSoi_W_m2 "
                "LngRadIn_W_m2 LngRadOut_W_m2A PotEvp_mm_h ActEvp_mm_h EvpTtrs_mm_h EvpWetCan_mm_h EvpDryCan_mm_h EvpSoil_mm_h "
                "Gflux_W_m2 HFlux_W_m2 Lflux_W_m2 NetPrecip_mm_hr LiqWE_cm IceWE_cm SnWE_cm SnSub_cm SnEvap_cm U_kJ_m2 RouteWE_cm "
                "SnTemp_C SurfAge_h DU_kJ_m2_etistep snLHF_kJ_m2_etistep snSHF_kJ_m2_etistep snGHF_kJ_m2_etistep "
                "snPHF_kJ_m2_etistep snRLout_kJ_m2_etistep snRLin_kJ_m2_etistep snRSin_kJ_m2_etistep Uerror_kJ_m2_etistep "
                "IntSWEq_cm IntSub_cm IntSnUnload_cm CanStorage_mm CumIntercept_mm Interception_mm Recharge_mm/hr RunOn_mm "
                "Srf_Hour_mm Qstrm_m3_s Hlevel_m CanStorParam_mm IntercepCoeff_[] ThroughFall_[] CanFieldCap_mm DrainCoeff_mm_hr "
                "DrainExpPar_1_mm LandUseAlb_[] VegHeight_m OptTransmCoeff_[] StomRes_s_m VegFraction[] LeafAI_[] ")
# plain columns 16..68 (after RootMoist, before Qstrm), all written with the stream's precision and no width
_PIXEL_PLAIN = ["AirTemp", "DewTemp", "SurfTemp", "SoilTemp", "AirPressure", "RelHumid", "SkyCover", "WindSpeed", "NetRad",
                "ShortRadIn", "ShortRadSlope", "ShortRadIn_dir", "ShortRadIn_dif", "ShortAbsbVeg", "ShortAbsbSoi", "LongRadIn",
                "LongRadOut", "PotEvap", "ActEvap", "EvapoTrans", "EvapWetCanopy", "EvapDryCanopy", "EvapSoil", "GFlux", "HFlux",
                "LFlux", "NetPrecipitation", "LiqWE", "IceWE", "_SnWE", "SnSub", "SnEvap", "Unode", "LiqRouted", "SnTempC",
                "CrustAge", "DU", "SnLHF", "SnSHF", "SnGHF", "SnPHF", "SnRLout", "SnRLin", "SnRSin", "Uerror", "IntSWE", "IntSub",
                "IntSnUnload", "CanStorage", "CumIntercept", "InterceptLoss", "Recharge", "RunOn", "srf_hr"]


class PixelWriter:
    """The `.pixel` file of one cell (index `cell` in sweep order; its node id goes into the rows)."""

    def __init__(self, path: str, basin: dict, cell: int, header: bool = True):
        self.f = open(path, "w")
        self.cell, self.first = int(cell), True
        self.node_id = int(np.asarray(basin["id"])[cell])
        slope = math.atan(float(np.asarray(basin["flow_slope"])[cell]))
        self.cos_slope = max(math.cos(slope), 1e-9)
        self.varea = float(np.asarray(basin["varea"])[cell])
        self.stream = int(np.asarray(basin["boundary"])[cell]) == 3
        self.land = [float(np.asarray(basin[k])[cell]) if k in basin else None for k in _LAND]
        if header:
            self.f.write(PIXEL_HEADER + "\n")

    def write_row(self, time_h: float, state: dict) -> None:
        c = self.cell

        def v(name):
            if name == "_SnWE":
                return v("LiqWE") + v("IceWE")
            return float(np.asarray(state[name])[c]) if name in state else 0.0

        hour = int(math.floor(time_h))
        ext = "%04d.%02d" % (hour, int((time_h - hour) * 100))
        cs = self.cos_slope
        p3 = 6 if self.first else 7                       # precision carried over from the previous row
        out = ["%8d" % self.node_id, "%13s" % ext, " ", "%10.*f" % (p3, v("NwtNew") / cs), " ",
               "%6.7f" % (v("NfNew") / cs), " ", "%6.7f" % (v("NtNew") / cs), " ", "%7.7f" % (v("MuNew") / cs), " ",
               "%7.7f" % (v("MiNew") / cs), "   ", "%10.7f" % (v("Qpout") * 1.0e-6 / self.varea), "  ",
               "%10.7f" % (v("Qpin") * 1.0e-6 / self.varea), "   ", "%10.7f" % (v("Transmissivity") * 1.0e-6), "    ",
               "%10.7f" % (v("gwchange") * 1.0e-9), "  ", "%8.7f" % v("srf"), "  ", "%10.7f" % v("Rain"), "  ",
               "%10.7f" % v("SoilMoistureSC"), "  ", "%10.7f" % v("RootMoistureSC"), " "]
        for name in _PIXEL_PLAIN:
            out += ["%.7f" % v(name), " "]
        if self.stream:
            out += ["%10.7f" % v("Qstrm"), " ", "%6.7f" % v("Hlevel"), " "]
        else:
            out.append("0.0 0.0 ")
        land = [x if x is not None else v(k) for x, k in zip(self.land, _LAND)]
        out.append("%10.7f" % land[0])
        for x in land[1:]:
            out += [" ", "%.7f" % x]
        self.f.write("".join(out) + "\n")
        self.f.flush()
        self.first = False

    def close(self):
        self.f.close()
