"""Test helpers: run the reference harness (oracle/_ref/tribs_ref_harness — the unmodified
reference compiled in place) on a synthetic deck and load its raw FP64 dumps.
The binary is prebuilt by __graft_entry__.build(); it needs no access to /root/reference at
run time, so it also runs on the GPU box."""
from __future__ import annotations

import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tribs_b200.synth import BasinSpec, write_basin  # noqa: E402
from tribs_b200.trb import read_trb  # noqa: E402

HARNESS = os.path.join(ROOT, "oracle", "_ref", "tribs_ref_harness")

# scenarios shared by the golden fixtures and the live parity tests
SCENARIOS = {
    # storm + interstorm on a deep water table: wetting-front states, lateral unsaturated flow
    "deep": dict(runtime_h=48.0),
    # shallow water table, heavy rain: water table at the surface, exfiltration, GW runoff
    "shallow": dict(runtime_h=48.0, gw_depth_mm=400.0, pmean=30.0),
    # short storm, short INTSTORMMAX: wedge destruction / redistribution, Storm_Evol
    "redistribute": dict(runtime_h=60.0, intstormmax=8.0, stdur=3.0, istdur=12.0, gw_depth_mm=1500.0),
    # net evaporation on dry steps (injected into tCNode::EvapSoil by the harness, EToption on):
    # water table dropping from the surface, interstorm drying below the initial profile
    # rain gauges + one weather station: Penman-Monteith energy balance, Rutter interception
    # (kernel 1) feeding the unsaturated zone through EvapSoil / EvapDryCanopy / NetPrecipitation
    "met": dict(runtime_h=48.0, forcing="met", gw_depth_mm=1500.0, pmean=12.0, stdur=5.0),
    # the same with sub-zero temperatures and OPTSNOW 1: snow pack energy balance, canopy snow
    # interception, melt routed to the unsaturated zone (oracle only so far: no device kernel yet)
    "snow": dict(runtime_h=72.0, forcing="met", winter=True, gw_depth_mm=1500.0, pmean=6.0, stdur=5.0),
    "evap": dict(runtime_h=72.0, gw_depth_mm=350.0, pmean=25.0, stdur=4.0, istdur=20.0, inject_evap=0.6),
}


# decks without a small golden fixture of their own (tests/golden/make_fullsize_met.py): "met" beginning at 09:00 with the
# wet spell at noon - daytime ET on dry ground, then evaporation from a wet canopy
EXTRA_SCENARIOS = {
    "metday": dict(runtime_h=48.0, forcing="met", gw_depth_mm=1500.0, pmean=12.0, stdur=5.0, start_hour=9, rain_hour=12),
    # BASELINE configs[4] in small: winter station, OPTSNOW 1 without a canopy, shallow water table; begins at 06:00, snow
    # falls from 08:00 on while the air warms through 0 degC towards noon
    "snowplain": dict(runtime_h=72.0, forcing="met", winter=True, gw_depth_mm=600.0, pmean=6.0, stdur=5.0, start_hour=6,
                      rain_hour=8, extra={"OPTINTERCEPT": 0}),
}


def have_harness() -> bool:
    return os.path.exists(HARNESS) and os.access(HARNESS, os.X_OK)


def spec_for(name: str, n: int, **over) -> BasinSpec:
    kw = dict(SCENARIOS[name] if name in SCENARIOS else EXTRA_SCENARIOS[name])
    kw.update(over)
    return BasinSpec(n=n, **kw)


def run_reference(spec: BasinSpec, workdir: str | None = None, snap_from=1, snap_to=0, snap_every=1,
                  phases="ugr", kat=False, timeout=3600):
    """Returns (basin dict, snapshots dict, timing dict)."""
    own = workdir is None
    d = tempfile.mkdtemp(prefix="tribsref_") if own else workdir
    write_basin(spec, d)
    cmd = [HARNESS, "basin.in", "-K", "--out", d, "--snap-from", str(snap_from), "--snap-to", str(snap_to),
           "--snap-every", str(snap_every), "--phases", phases]
    if kat:
        cmd.append("--kat")
    if spec.inject_evap > 0.0:
        cmd += ["--inject-evap", repr(spec.inject_evap)]
    r = subprocess.run(cmd, cwd=d, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError("reference harness failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    import json
    m = re.search(r"ref_timing (\{.*\})", r.stdout)
    timing = json.loads(m.group(1)) if m else {}
    basin = read_trb(os.path.join(d, "basin.trb"))
    snaps = read_trb(os.path.join(d, "steps.trb")) if snap_to >= snap_from else {}
    katd = read_trb(os.path.join(d, "kat.trb")) if kat else {}
    basin["_spec_pmean"] = np.array([spec.pmean])
    basin["_spec_stdur"] = np.array([spec.stdur])
    basin["_spec_istdur"] = np.array([spec.istdur])
    basin["_spec_evap"] = np.array([spec.inject_evap])
    if own:
        import shutil
        shutil.rmtree(d, ignore_errors=True)
    return basin, snaps, timing, katd


def write_meshb_deck(side: int, workdir: str, seed: int = 1, runtime_h: float = 24.0, scenario: str = "deep", **over):
    """A deck the reference can run at any size: synthbasin.build_basin(side, side) written as
    OPTMESHINPUT 9 files (no mesh construction, no flow-network set-up in the reference), the initial
    water table per Voronoi cell (OPTGWFILE 1, tHydroModel.cpp:249-269; the coarse gw.asc grid would
    make tResample intersect it with every cell polygon). `scenario` / `over` as for spec_for (forcing
    files, options). Returns build_basin's dict."""
    from tribs_b200.synthbasin import build_basin
    spec = spec_for(scenario, side + 2, runtime_h=runtime_h, seed=seed, **over)
    write_basin(spec, workdir)                                                          # .in, soil / land grids, forcing
    basin = build_basin(side, side, seed=seed, runtime_h=runtime_h, mesh_dir=workdir, gw_depth_mm=spec.gw_depth_mm)
    inp = os.path.join(workdir, "basin.in")
    txt = re.sub(r"(OPTMESHINPUT:\s*\n)\s*\d+", r"\g<1>9", open(inp).read())
    txt = re.sub(r"(OPTGWFILE:\s*\n)\s*\d+", r"\g<1>1", txt)
    txt = re.sub(r"(GWATERFILE:\s*\n)\s*\S+", r"\g<1>gw.vor", txt)
    open(inp, "w").write(txt)
    with open(os.path.join(workdir, "gw.vor"), "w") as f:
        f.write("".join(f"{i} {v!r}\n" for i, v in enumerate(np.asarray(basin["init_NwtOld"]).tolist())))
    return basin


def run_reference_meshb(side: int, scenario: str, runtime_h: float, seed: int = 1, snap_from=1, snap_to=0, snap_every=1,
                        phases="ugr", timeout=3600, **over):
    """run_reference on a synthetic basin the reference READS (OPTMESHINPUT 9) instead of building it:
    set-up stays linear, so decks of 1e4-1e6 cells are affordable. Returns (basin, snapshots, timing)
    with the basin as flattened by the reference itself."""
    import json
    d = tempfile.mkdtemp(prefix="tribsref9_")
    write_meshb_deck(side, d, seed=seed, runtime_h=runtime_h, scenario=scenario, **over)
    cmd = [HARNESS, "basin.in", "-K", "--out", d, "--snap-from", str(snap_from), "--snap-to", str(snap_to),
           "--snap-every", str(snap_every), "--phases", phases]
    r = subprocess.run(cmd, cwd=d, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError("reference harness failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    m = re.search(r"ref_timing (\{.*\})", r.stdout)
    timing = json.loads(m.group(1)) if m else {}
    basin = read_trb(os.path.join(d, "basin.trb"))
    snaps = read_trb(os.path.join(d, "steps.trb")) if snap_to >= snap_from else {}
    spec = spec_for(scenario, side + 2, runtime_h=runtime_h, seed=seed, **over)
    basin["_spec_pmean"], basin["_spec_stdur"] = np.array([spec.pmean]), np.array([spec.stdur])
    basin["_spec_istdur"], basin["_spec_evap"] = np.array([spec.istdur]), np.array([spec.inject_evap])
    import shutil
    shutil.rmtree(d, ignore_errors=True)
    return basin, snaps, timing


def storm_of(basin):
    return float(basin["_spec_pmean"][0]), float(basin["_spec_stdur"][0]), float(basin["_spec_istdur"][0])


def schedule_for(basin):
    """The rain forcing of a reference run as a callable of model time (scalar or per-cell mm/h)."""
    if "rain_optStorm" in basin and not int(basin["rain_optStorm"][0]) and int(basin["rain_type"][0]) == 3:
        from tribs_b200.met import GaugeRain
        return GaugeRain(basin)
    from oracle.pyoracle import stochastic_storm_schedule
    return stochastic_storm_schedule(*storm_of(basin))


# scenarios the device path covers (tSnowPack has an oracle but no kernel yet)
GPU_SCENARIOS = [k for k in SCENARIOS if k != "snow"]


def has_et(basin) -> bool:
    return "et_option" in basin and int(basin["et_option"][0]) != 0


def evap_of(basin) -> float:
    return float(basin["_spec_evap"][0]) if "_spec_evap" in basin else 0.0


def load_golden(name: str):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    basin = {k[2:]: z[k] for k in z.files if k.startswith("b/")}
    snaps = {k[2:]: z[k] for k in z.files if k.startswith("s/")}
    kat = {k[2:]: z[k] for k in z.files if k.startswith("k/")}
    return basin, snaps, kat


def rel_err(ref, got, floor=1e-12):
    ref = np.asarray(ref, float)
    got = np.asarray(got, float)
    den = np.maximum(np.abs(ref), floor)
    e = np.abs(ref - got) / den
    e[(ref == got)] = 0.0
    return e
