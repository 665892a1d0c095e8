"""Full-size parity for BASELINE.json configs[2]: the 10M-cell basin under rain gauges + a weather station with
Penman-Monteith ET and Rutter interception, the configuration bench.py times by default.

tests/golden/fullsize_met_10m.npz holds what the UNMODIFIED REFERENCE leaves after 40 and 96 steps of a run that begins at
09:00 (`snap_steps`: dry ground under the late-morning sun, then the gauges' storm from step 48 on: wet-canopy evaporation,
wetting fronts; six ET sweeps, twelve groundwater steps; tests/golden/make_fullsize_met.py, scenario "metday"; the C oracle
reproduced the reference's full arrays - state and the outputs of kernel (1) - bit for bit before the fixture was written). The device runs the same 96 steps the way bench.py does: pipelined batches with
kernel (1) as a phase of the batch, the rain as per-gauge values + the static Thiessen map on the device.

Bar (north_star): per-cell Nwt, Nf, Mu ... and the hydrograph within 1e-6 relative."""
import os

import numpy as np
import pytest

from refrun import ROOT

pytestmark = pytest.mark.gpu

REL_TOL = 1e-6
FIXTURE = os.environ.get("TRIBS_FULLSIZE_FIXTURE") or os.path.join(ROOT, "tests", "golden", "fullsize_met_10m.npz")
SNOW_FIXTURE = os.environ.get("TRIBS_FULLSIZE_SNOW_FIXTURE") or os.path.join(ROOT, "tests", "golden", "fullsize_snow_6m.npz")


@pytest.mark.skipif(not os.path.exists(FIXTURE), reason="tests/golden/fullsize_met_10m.npz not generated")
def test_gpu_config3_full_size_matches_reference_fixture(built):
    check_against_fixture(FIXTURE, "config 3")


@pytest.mark.skipif(not os.path.exists(SNOW_FIXTURE), reason="tests/golden/fullsize_snow_6m.npz not generated")
def test_gpu_config5_full_size_matches_reference_fixture(built):
    """BASELINE.json configs[4] at the size one GPU carries in bench.py --config 5 (6.25M cells): winter station, OPTSNOW 1
    without a canopy (the snow pack as a phase of the batch), shallow water table; the run begins at 06:00, snow falls from
    08:00 on, the pack turns isothermal towards noon and melt reaches the soil in the afternoon (steps 96 and 160 of the
    reference; scenario "snowplain" of tests/golden/make_fullsize_met.py)."""
    z = check_against_fixture(SNOW_FIXTURE, "config 5")
    last = int(z["snap_steps"][-1])
    assert z[f"s{last}/CumMelt"].max() > 0 and z[f"s{last}/IceWE"].max() > 0 and z[f"s{last}/LiqRouted"].max() > 0


def check_against_fixture(path, label):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_fullsize_met import BLOCK, RESULTS, RUNTIME_H, STATE, apply_patches, block_sums
    from tribs_b200 import gpu, met
    from tribs_b200.model import Simulator
    from tribs_b200.synthbasin import build_basin
    z = np.load(path)
    side, seed = int(z["side"][0]), int(z["seed"][0])
    basin = apply_patches(build_basin(side, side, seed=seed, runtime_h=RUNTIME_H, gw_depth_mm=float(z["gw_depth_mm"][0])), z)
    for k in z.files:
        if k.startswith("extra/"):
            basin[k[6:]] = np.array(z[k])
    assert int(basin["n_active"][0]) == int(z["n_active"][0]) and int(z["oracle_bitexact"][0]) == 1
    idx = z["cells"]
    rain = met.GaugeRain(basin)
    sim = Simulator(basin)
    sim.dev.set_rain_gauges(rain.cell_gauge.astype(np.int32), len(rain.elev))
    done, worst = 0, {}
    for st in [int(s) for s in z["snap_steps"]]:
        sim.run_batch(st - done, rain_of_time=lambda t: gpu.GaugeRain(rain.gauge_values(t).copy()))
        done = st
        got = sim.dev.sync(list(STATE))
        for f in STATE:
            ref = z[f"s{st}/{f}"]
            err = np.abs(ref - got[f][idx]) / np.maximum(np.abs(ref), 1e-6)
            worst[(st, f)] = float(err.max())
            ref_b, got_b = z[f"s{st}/blk/{f}"], block_sums(got[f])
            scale = np.maximum(np.abs(ref_b), 1e-6 * BLOCK)
            worst[(st, "blk", f)] = float((np.abs(ref_b - got_b) / scale).max())
        et_fields = [str(f) for f in z["et_fields"]] if "et_fields" in z.files else []
        if et_fields:                                   # outputs of kernel (1): temperatures and fluxes pass through zero
            from common import ET_FLOOR
            floors = dict(ET_FLOOR)          # snow pack: degC, W/m2 and kJ/m2 quantities pass through zero
            floors.update({"SnTempC": 1.0, "SnLHF": 1.0, "SnSHF": 1.0, "SnGHF": 1.0, "SnPHF": 1.0, "SnRLin": 1.0, "SnRLout": 1.0,
                           "SnRSin": 1.0, "DU": 1e-3, "Unode": 1e-3})
            got_et = {}
            for getter, table in ((sim.dev.sync, gpu.F), (sim.dev.et_sync, gpu.EF), (sim.dev.snow_sync, gpu.SNF)):
                names = [f for f in et_fields if f in table and f not in got_et]
                if names:
                    got_et.update(getter(names))
            for f in et_fields:
                ref = z[f"s{st}/{f}"]
                err = np.abs(ref - got_et[f][idx]) / np.maximum(np.abs(ref), floors.get(f, 1e-6))
                worst[(st, "et", f)] = float(err.max())
        for r in RESULTS:
            ref = z[f"s{st}/res/{r}"]
            dev = np.asarray(sim.dev.results(r, ref.size))[: ref.size]
            worst[(st, "res", r)] = float(np.abs(ref - dev).max() / max(np.abs(ref).max(), 1e-6))
    rows = sorted(worst.items(), key=lambda kv: -kv[1])
    print(f"\n[{label}, full size] worst relative differences:\n" + "\n".join(f"  {k}: {v:.3e}" for k, v in rows[:10]))
    assert rows[0][1] <= REL_TOL, rows[:10]
    # the sample must have seen the physics: rain through the canopy, wetting fronts, lateral outflow
    last = int(z["snap_steps"][-1])
    assert np.abs(z[f"s{last}/MuOld"] - z[f"s{int(z['snap_steps'][0])}/MuOld"]).max() > 0
    assert z[f"s{last}/NfOld"].max() > 0
    if "et_fields" in z.files and "EvapWetCanopy" in z["et_fields"]:   # daytime fixture: evaporation on dry ground, then from a wet canopy
        first = int(z["snap_steps"][0])
        assert z[f"s{last}/Rain"].max() > 0 and (z[f"s{last}/NetPrecipitation"] < z[f"s{last}/Rain"]).any()
        assert z[f"s{first}/EvapSoil"].max() > 0 and z[f"s{first}/PotEvap"].max() > 0 and z[f"s{last}/EvapWetCanopy"].max() > 0
    return z
