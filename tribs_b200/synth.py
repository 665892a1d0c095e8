"""Synthetic basin generator for the tRIBS hot-path work.

Writes a complete, self-contained reference-format input deck (``.in`` control
file, points file, soil/land tables, ASCII grids) for a V-shaped valley TIN, so
that the *unmodified* reference serial model and the B200 path
are driven from identical files.

Formats follow the reference readers:
  * points file ``NP`` + ``x y z b`` rows, boundary codes 0 interior / 1 closed /
    2 outlet / 3 stream            (reference: src/tMesh/tMesh.cpp:560-640)
  * ``.in`` = ``KEYWORD:`` line followed by a value line; lookup is a prefix match
    so CHANNELWIDTHCOEFF must precede CHANNELWIDTH
                                   (reference: src/tInOut/tInputFile.cpp:76-108)
  * soil table ``nclass 12`` / land table ``nclass 15``
                                   (reference: src/tRasTin/tInvariant.cpp:35-110, 544-590)
  * ArcInfo ASCII grids            (reference: src/tRasTin/tResample.cpp:1066-1129)
"""
from __future__ import annotations

import math
import os
import random
from dataclasses import dataclass, field


@dataclass
class BasinSpec:
    n: int = 24                 # points per side (n*n points incl. boundary ring)
    spacing: float = 30.0       # m
    jitter: float = 3.0         # +- m on interior points
    seed: int = 1
    runtime_h: float = 12.0
    timestep_min: float = 3.75
    gwstep_min: float = 30.0
    metstep_min: float = 60.0
    gw_depth_mm: float = 3000.0   # uniform initial water-table depth
    gw_plane: float = 0.0         # extra depth per metre of |x-xc| (mm/m): tilted table
    bedrock_m: float = 15.0
    pmean: float = 20.0           # mm/h  storm intensity (STOCHASTICMODE 1)
    stdur: float = 6.0            # h
    istdur: float = 18.0          # h
    intstormmax: float = 10000.0
    mesh_option: int = 8
    slope_y: float = 0.02
    slope_x: float = 0.05
    tributary_every: int = 0      # >0: add a stream row every k rows (dendritic net)
    soil: tuple = (10.0, 0.45, 0.05, 0.4, -200.0, 0.001, 100.0, 100.0, 0.45, 0.7, 1400000.0)
    land: tuple = (0.5, 0.2, 0.5, 1.0, 0.1, 3.7, 0.2, 1.0, 0.8, 100.0, 0.5, 2.0, 0.3, 0.25)
    opt_spatial: int = 0
    opt_interhydro: int = 0
    forcing: str = "storm"        # "storm": STOCHASTICMODE 1 uniform storm; "met": rain gauges +
                                  # one weather station (Penman-Monteith ET, Rutter interception)
    n_gauges: int = 2
    winter: bool = False          # met deck: sub-zero air temperatures and OPTSNOW 1
    start_hour: int = 0           # met deck: hour of day the run (STARTDATE) and the forcing files begin at
    rain_hour: int = 3            # met deck: hour of day the daily wet spell begins
    sky_nodata: bool = False      # met deck: write 9999.99 as sky cover (the model then derives it from RH / rain)
    dem_ridge: float = 0.0        # OPTRADSHELT 1-3: height of a wall on the east margin of the DEM (m), to cast shade
    inject_evap: float = 0.0      # test forcing, not a deck keyword: soil evaporation on dry steps (mm/h)
    extra: dict = field(default_factory=dict)   # keyword overrides


def make_points(spec: BasinSpec):
    """Return list of (x, y, z, b) in file order."""
    rng = random.Random(spec.seed)
    n, h = spec.n, spec.spacing
    ic = n // 2
    xc = ic * h
    pts = []
    for j in range(n):
        for i in range(n):
            x, y = i * h, j * h
            edge = (i == 0 or j == 0 or i == n - 1 or j == n - 1)
            if edge:
                b = 2 if (j == 0 and i == ic) else 1
            else:
                b = 3 if i == ic else 0
                if spec.tributary_every and b == 0 and j % spec.tributary_every == 0 \
                        and 2 <= i <= n - 3:
                    b = 3
                if b == 0:
                    x += rng.uniform(-spec.jitter, spec.jitter)
                    y += rng.uniform(-spec.jitter, spec.jitter)
            z = 100.0 + spec.slope_y * y + spec.slope_x * abs(x - xc)
            if spec.tributary_every and not edge:
                # shallow side valleys draining along tributary rows
                jj = j % spec.tributary_every
                d = min(jj, spec.tributary_every - jj) * h
                z += 0.03 * d
            pts.append((x, y, z, b))
    return pts


def _asc(path, xll, yll, size, ncell, value):
    with open(path, "w") as f:
        f.write(f"ncols {ncell}\nnrows {ncell}\nxllcorner {xll}\nyllcorner {yll}\n"
                f"cellsize {size}\nNODATA_value -9999\n")
        row = " ".join([str(value)] * ncell)
        for _ in range(ncell):
            f.write(row + "\n")


def met_series(spec: BasinSpec):
    """Hourly forcing of the "met" decks: rain per gauge (mm/h) and one weather station's
    PA RH XC US TA IS columns. Deterministic, smooth diurnal cycles with two wet spells."""
    nh = int(math.ceil(spec.runtime_h)) + 2
    rain = [[0.0] * nh for _ in range(spec.n_gauges)]
    for g in range(spec.n_gauges):
        for h in range(nh):
            ph = (h + spec.start_hour) % 24
            if spec.rain_hour <= ph < spec.rain_hour + int(spec.stdur):
                rain[g][h] = round(spec.pmean * (0.6 + 0.4 * g)
                                   * (0.5 + 0.5 * math.sin(math.pi * (ph - spec.rain_hour + 0.5) / spec.stdur)), 3)
    met = []
    for h in range(nh):
        ph = (h + spec.start_hour) % 24
        day = math.sin(math.pi * (ph - 6) / 12.0) if 6 <= ph <= 18 else 0.0
        wet = any(rain[g][h] > 0 for g in range(spec.n_gauges))
        ta = round((-4.0 if spec.winter else 16.0) + (7.0 if spec.winter else 9.0) * math.sin(2 * math.pi * (ph - 9) / 24.0)
                   + (4.0 * math.sin(2 * math.pi * h / 90.0) if spec.winter else 0.0), 2)
        rh = round(min(98.0, 55.0 - 25.0 * math.sin(2 * math.pi * (ph - 9) / 24.0) + (25.0 if wet else 0.0)), 2)
        xc = 9.0 if wet else round(2.0 + 2.0 * math.cos(2 * math.pi * h / 37.0), 1)
        us = round(2.5 + 1.5 * math.sin(2 * math.pi * h / 17.0), 2)
        isw = round(850.0 * day * (0.25 if wet else 1.0), 1)
        met.append((980.0, rh, xc, us, ta, isw))
    return rain, met


def write_met_forcing(spec: BasinSpec, outdir: str) -> dict:
    """Rain-gauge (.sdf/.mdf, tRainfall.cpp:551-707) and weather-station files
    (tEvapoTrans.cpp:3122-3330, columns Y M D H PA RH XC US TA IS NR); returns the keywords."""
    rain, met = met_series(spec)
    nh = len(met)
    ext = (spec.n - 1) * spec.spacing

    def stamp(h):
        h += spec.start_hour
        return f"2000 6 {1 + h // 24} {h % 24}"

    with open(os.path.join(outdir, "gauges.sdf"), "w") as f:
        f.write(f"{spec.n_gauges} 7\n")
        for g in range(spec.n_gauges):
            gx = ext * (g + 0.5) / spec.n_gauges
            f.write(f"{g + 1} {outdir}/gauge{g + 1}.mdf {ext / 2:.3f} {gx:.3f} {nh} 5 110.0\n")
            with open(os.path.join(outdir, f"gauge{g + 1}.mdf"), "w") as d:
                d.write("Y M D H R\n")
                for h in range(nh):
                    d.write(f"{stamp(h)} {rain[g][h]}\n")
    with open(os.path.join(outdir, "met.sdf"), "w") as f:
        f.write("1 10\n")
        f.write(f"1 {outdir}/met1.mdf 34.0 {ext / 2:.3f} -107.0 {ext / 2:.3f} -7 {nh} 11 105.0\n")
    with open(os.path.join(outdir, "met1.mdf"), "w") as d:
        d.write("Y M D H PA RH XC US TA IS NR\n")
        for h in range(nh):
            pa, rh, xc, us, ta, isw = met[h]
            d.write(f"{stamp(h)} {pa} {rh} {9999.99 if spec.sky_nodata else xc} {us} {ta} {isw} 9999.99\n")
    kw = {"STOCHASTICMODE": 0, "RAINSOURCE": 3, "GAUGESTATIONS": f"{outdir}/gauges.sdf",
          "OPTEVAPOTRANS": 1, "OPTINTERCEPT": 2, "METDATAOPTION": 1,
          "HYDROMETSTATIONS": f"{outdir}/met.sdf"}
    if spec.winter:
        kw["OPTSNOW"] = 1
    return kw


def write_basin(spec: BasinSpec, outdir: str) -> str:
    """Write the input deck under *outdir*; returns the path of the ``.in`` file."""
    os.makedirs(outdir, exist_ok=True)
    os.makedirs(os.path.join(outdir, "out"), exist_ok=True)
    pts = make_points(spec)
    with open(os.path.join(outdir, "basin.points"), "w") as f:
        f.write(f"{len(pts)}\n")
        for x, y, z, b in pts:
            f.write(f"{x:.6f} {y:.6f} {z:.6f} {b}\n")
    with open(os.path.join(outdir, "soil.sdt"), "w") as f:
        f.write("1 12\n1 " + " ".join(repr(v) for v in spec.soil) + "\n")
    with open(os.path.join(outdir, "land.ldt"), "w") as f:
        f.write("1 15\n1 " + " ".join(repr(v) for v in spec.land) + "\n")
    ext = (spec.n - 1) * spec.spacing
    size = (ext + 400.0) / 4.0
    for name, val in (("soil.asc", 1), ("land.asc", 1), ("gw.asc", spec.gw_depth_mm)):
        _asc(os.path.join(outdir, name), -200.0, -200.0, size, 4, val)
    with open(os.path.join(outdir, "nodes.pix"), "w") as f:
        f.write("0\n")
    kv = [
        ("STARTDATE", f"06/01/2000/{spec.start_hour:02d}/00"), ("RUNTIME", spec.runtime_h),
        ("TIMESTEP", spec.timestep_min), ("GWSTEP", spec.gwstep_min),
        ("METSTEP", spec.metstep_min), ("RAININTRVL", 1), ("OPINTRVL", 1),
        ("SPOPINTRVL", 100000), ("INTSTORMMAX", spec.intstormmax), ("RAINSEARCH", 24),
        ("BASEFLOW", 0.01), ("VELOCITYCOEF", 0.5), ("KINEMVELCOEF", 0.1),
        ("VELOCITYRATIO", 5), ("FLOWEXP", 0.0), ("CHANNELROUGHNESS", 0.15),
        ("CHANNELWIDTHCOEFF", 1.0), ("CHANNELWIDTH", 10), ("CHANNELWIDTHEXPNT", 0.3),
        ("CHANNELWIDTHFILE", "none"), ("WIDTHINTERPOLATION", 0), ("OPTPERCOLATION", 0),
        ("TLINKE", 2.5), ("DEPTHTOBEDROCK", spec.bedrock_m),
        ("OPTMESHINPUT", spec.mesh_option), ("RAINSOURCE", 3), ("OPTEVAPOTRANS", 0),
        ("OPTSNOW", 0), ("HILLALBOPT", 0), ("OPTRADSHELT", 0), ("OPTINTERCEPT", 0),
        ("OPTLANDUSE", 0), ("OPTLUINTERP", 0), ("OPTSOILTYPE", 0), ("GFLUXOPTION", 2),
        ("METDATAOPTION", 0), ("CONVERTDATA", 0), ("OPTBEDROCK", 0), ("OPTGWFILE", 0),
        ("OPTRUNON", 0), ("OPTRESERVOIR", 0), ("OPTGROUNDWATER", 1),
        ("OPTINTERHYDRO", spec.opt_interhydro), ("OPTHEADER", 1),
        ("OPTSPATIAL", spec.opt_spatial), ("MINSNTEMP", -20), ("SNLIQFRAC", 0.06),
        ("TEMPLAPSE", 0), ("PRECLAPSE", 0), ("FORECASTMODE", 0), ("STOCHASTICMODE", 1),
        ("PMEAN", spec.pmean), ("STDUR", spec.stdur), ("ISTDUR", spec.istdur),
        ("SEED", 11), ("RESTARTMODE", 0), ("PARALLELMODE", 0), ("OPTVIZ", 0),
        ("POINTFILENAME", "basin.points"), ("SOILTABLENAME", "soil.sdt"),
        ("SOILMAPNAME", "soil.asc"), ("LANDTABLENAME", "land.ldt"),
        ("LANDMAPNAME", "land.asc"), ("GWATERFILE", "gw.asc"),
        ("OUTFILENAME", "out/sp"), ("OUTHYDROFILENAME", "out/hy"),
        ("OUTHYDROEXTENSION", "mrf"), ("NODEOUTPUTLIST", "nodes.pix"),
        ("HYDRONODELIST", "none"), ("OUTLETNODELIST", "none"),
    ]
    over = dict(spec.extra)
    if int(over.get("OPTRADSHELT", 0)) in (1, 2, 3):      # tShelter derives horizon angles from a DEM grid (DEMFILE)
        h, n = spec.spacing, spec.n
        xc = (n // 2) * h
        with open(os.path.join(outdir, "dem.asc"), "w") as f:
            m = n + 8                                      # 4 cells of margin around the points
            f.write(f"ncols {m}\nnrows {m}\nxllcorner {-4.5 * h}\nyllcorner {-4.5 * h}\ncellsize {h}\nNODATA_value -9999\n")
            for j in range(m - 1, -1, -1):                 # rows from north to south
                y = (j - 4) * h
                f.write(" ".join(repr(100.0 + spec.slope_y * y + spec.slope_x * abs((i - 4) * h - xc)
                                      + (spec.dem_ridge if i >= m - 3 else 0.0)) for i in range(m)) + "\n")
        over.setdefault("DEMFILE", "dem.asc")
    if spec.forcing == "met":
        for k, v in write_met_forcing(spec, outdir).items():
            over.setdefault(k, v)
    path = os.path.join(outdir, "basin.in")
    with open(path, "w") as f:
        for k, v in kv:
            v = over.pop(k, v)
            f.write(f"{k}:\n{v}\n\n")
        for k, v in over.items():
            f.write(f"{k}:\n{v}\n\n")
    return path


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("outdir")
    ap.add_argument("--n", type=int, default=24)
    ap.add_argument("--runtime", type=float, default=12.0)
    ap.add_argument("--trib", type=int, default=0)
    a = ap.parse_args()
    print(write_basin(BasinSpec(n=a.n, runtime_h=a.runtime, tributary_every=a.trib), a.outdir))
