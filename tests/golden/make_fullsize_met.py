"""Full-size golden vectors for BASELINE.json configs[2], made by the REFERENCE ITSELF: the 10M-cell V-valley
(build_basin(3163, 3163, seed=1)) under rain gauges + a weather station with Penman-Monteith ET and Rutter
interception (scenario "metday" of tests/refrun.py: the deck "met" beginning at 09:00 with the wet spell at noon; the
reference's own .sdf/.mdf forcing files, METSTEP 60 min).

Like make_fullsize.py: the basin travels to the reference as OPTMESHINPUT 9 files, the unmodified reference
(oracle/_ref/tribs_ref_harness) runs it (32 GB, ~4 min set-up, ~20 s per step on one core), and the C oracle
(unsaturated / saturated zone, routing, ET + interception) has to reproduce the reference's FULL arrays bit for
bit at both snapshots before the fixture is written. The fixture keeps, per snapshot (steps 40 and 96: the gauges' storm
begins at hour 3 = step 48, so the window has daytime ET on dry ground, then the canopy filling under rain, wet-canopy
evaporation, wetting fronts; six ET sweeps, twelve groundwater steps): a sample of cells with the state and the outputs of
kernel (1), 1000-cell block sums of every state field, the hydrograph
arrays; plus what turns build_basin()'s dict into the reference's exact inputs (sparse patches, and the
forcing / land / ET keys whole: they are tiny or compress to nothing).

    python tests/golden/make_fullsize_met.py [side]     # side 3163: ~75 min, 40 GB; writes tests/golden/fullsize_met_10m.npz"""
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

from make_fullsize import BLOCK, apply_patches, block_sums, make_patches  # noqa: E402,F401

SIDE, SEED, RUNTIME_H = 3163, 1, 24.0
SCENARIO = os.environ.get("TRIBS_FULLSIZE_SCENARIO", "metday")     # tests/refrun.py SCENARIOS: "met" or "metday"
SNAP_STEPS = {"metday": (40, 96), "snowplain": (96, 160)}.get(SCENARIO, (64, 96))
STATE = {"NwtOld": "NwtNew", "MuOld": "MuNew", "MiOld": "MiNew", "NtOld": "NtNew", "NfOld": "NfNew", "RuOld": "RuNew",
         "RiOld": "RiNew", "Qpout": "Qpout", "cumsrf": "cumsrf", "intstorm": "intstorm", "NetPrecipitation": "NetPrecipitation",
         "EvapSoil": "EvapSoil", "EvapDryCanopy": "EvapDryCanopy", "Rain": "Rain"}
# outputs of kernel (1) kept beside the water-balance state (device: the ET field block, tribs_gpu_et_sync)
ET_STATE = ["PotEvap", "ActEvap", "EvapoTrans", "EvapWetCanopy", "CanStorage", "CumIntercept", "SurfTemp", "NetRad", "LFlux", "HFlux",
            "GFlux"]
if SCENARIO == "snowplain":     # BASELINE configs[4]: the snow pack replaces both ET sweeps (tSnowPack::callSnowPack)
    ET_STATE = ["LiqWE", "IceWE", "SnTempC", "LiqRouted", "SnSub", "SnEvap", "CumMelt", "PeakSWE", "DU", "Unode", "SnLHF", "SnSHF", "SnGHF",
                "SnPHF", "SnRLin", "SnRLout", "SnRSin", "CrustAge", "DensityAge", "CumHrsSnow", "AirTemp", "NetRad"]
RESULTS = ["mhydro", "HsrfRout", "SbsrfRout", "SatsrfRout"]


def sample_cells(basin):
    """every stream cell + a strided sample of about 10 000 hillslope cells"""
    n = int(basin["n_active"][0])
    stream = np.nonzero(np.asarray(basin["boundary"]) == 3)[0]
    strided = np.arange(0, n, max(251, n // 10000) | 1)
    return np.unique(np.concatenate([stream, strided])).astype(np.int64)


def reference_run(side, workdir, snap_steps):
    import json
    import re
    import subprocess
    from refrun import HARNESS, write_meshb_deck
    from tribs_b200.trb import read_trb
    basin = write_meshb_deck(side, workdir, seed=SEED, runtime_h=RUNTIME_H, scenario=SCENARIO)
    steps = sorted(snap_steps)
    every = steps[1] - steps[0]
    cmd = [HARNESS, "basin.in", "-K", "--out", workdir, "--snap-from", str(steps[0]), "--snap-to", str(steps[-1]),
           "--snap-every", str(every), "--phases", "rm", "--max-steps", str(steps[-1])]
    r = subprocess.run(cmd, cwd=workdir, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference harness failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    m = re.search(r"ref_timing (\{.*\})", r.stdout)
    return basin, read_trb(os.path.join(workdir, "basin.trb")), read_trb(os.path.join(workdir, "steps.trb")), \
        (json.loads(m.group(1)) if m else {})


def main(side=SIDE, out=None):
    from common import OracleStepper
    from refrun import schedule_for
    t0 = time.time()
    work = tempfile.mkdtemp(prefix="tribs_fullsize_met_")
    basin, ref_basin, snaps, timing = reference_run(side, work, SNAP_STEPS)
    print(f"reference run: {time.time() - t0:.0f} s  {timing}", flush=True)
    from refrun import EXTRA_SCENARIOS, SCENARIOS
    z = {"side": np.array([side]), "seed": np.array([SEED]), "snap_steps": np.array(SNAP_STEPS),
         "gw_depth_mm": np.array([float({**SCENARIOS, **EXTRA_SCENARIOS}[SCENARIO]["gw_depth_mm"])]),       # build_basin argument of the deck
         "n_active": np.array([int(basin["n_active"][0])]), "ref_cell_steps_per_s": np.array([timing.get("cell_steps_per_s", 0.0)]),
         "ref_setup_s": np.array([timing.get("setup_s", 0.0)]), "ref_loop_s": np.array([timing.get("loop_s", 0.0)])}
    z.update(make_patches(basin, ref_basin))
    # forcing, land and ET keys build_basin does not make; mesh bookkeeping the device path never reads (edge ids,
    # complementary-edge ids ...) and the initial storages of tWaterBalance (not compared here) stay out: they are the
    # only large ones
    skip = {"spoke_compl_id", "spoke_edge_id", "edge_id", "edge_dest", "edge_org", "flow_edge_id", "init_TTime",
            "init_UnSaturatedStorage", "init_SaturatedStorage"}
    for k, a in ref_basin.items():
        if k not in basin and k not in skip:
            z[f"extra/{k}"] = np.asarray(a)
    idx = sample_cells(ref_basin)
    z["cells"] = idx
    z["et_fields"] = np.array(ET_STATE)
    for st in SNAP_STEPS:
        for f, src in STATE.items():
            a = snaps[f"s{st}_r_{src}"]
            z[f"s{st}/{f}"] = a[idx].copy()
            z[f"s{st}/blk/{f}"] = block_sums(a)
        for f in ET_STATE:
            z[f"s{st}/{f}"] = snaps[f"s{st}_m_{f}"][idx].copy()       # kernel (1) runs first in a step; nothing later writes these
        for r in RESULTS:
            z[f"s{st}/res/{r}"] = np.array(snaps[f"s{st}_r_{r}"]).copy()
    del basin
    o = OracleStepper(ref_basin, schedule_for(ref_basin))
    bitexact = 1
    for st in range(1, max(SNAP_STEPS) + 1):
        o.advance(); o.unsat()
        if o.gw_due():
            o.sat()
        o.route(); o.balance()
        if st in SNAP_STEPS:
            got = o.get(list(STATE.values()) + ET_STATE)
            for f, src in list(STATE.items()) + [(f, f) for f in ET_STATE]:
                ref_a = snaps[f"s{st}_r_{src}" if f in STATE else f"s{st}_m_{src}"]
                same = np.array_equal(got[src], ref_a)
                bitexact &= int(same)
                if not same:
                    e = np.abs(got[src] - ref_a)
                    print(f"  step {st} {src}: oracle differs from the reference in {int((e > 0).sum())} cells, max {e.max():.3e}")
            for r in RESULTS:
                bitexact &= int(np.array_equal(np.asarray(o.result(r)), snaps[f"s{st}_r_{r}"]))
            print(f"step {st}: oracle == reference bit for bit: {bool(bitexact)}  ({time.time() - t0:.0f} s)", flush=True)
        o.reset()
    assert bitexact == 1, "the oracle does not reproduce the reference on this basin"
    z["oracle_bitexact"] = np.array([bitexact])
    name = f"fullsize_met_{round(side * side / 1e6)}m.npz" if side >= 1000 else f"fullsize_met_{side}.npz"
    out = out or os.path.join(HERE, name)
    np.savez_compressed(out, **z)
    import shutil
    shutil.rmtree(work, ignore_errors=True)
    print("wrote", out, os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else SIDE, sys.argv[2] if len(sys.argv) > 2 else None)
