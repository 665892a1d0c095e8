"""Full-size golden vectors made by the REFERENCE ITSELF: BASELINE.json configs[1] - the basin and
storm bench.py times (build_basin(1000, 1000, seed=1): 1M Voronoi cells, 20 mm/h for 6 h).

The reference cannot build a 1M-node mesh from points (its flow-network set-up is super-linear:
141 s for 89k nodes), but it can READ one: synthbasin.build_basin(mesh_dir=...) writes the basin as
OPTMESHINPUT 9 files (tribs_b200/meshb.py) and the unmodified reference (oracle/_ref/tribs_ref_harness)
runs it: set-up ~30 s, 128 steps ~6 min on one core. The storm ends at step 96, so the window
contains the storm -> interstorm regime change.

What the reference derives differently from build_basin's closed forms (Voronoi areas of the cells
on the hull, which tResample clips; -1 as "stream node" of stream cells; hydrograph length) is kept in
the fixture as sparse PATCHES of the basin dict, so a test builds the reference's exact inputs from
build_basin() + apply_patches() without a 1 GB dump. The generator then checks that the C oracle on
those inputs reproduces the reference's full arrays BIT FOR BIT at both snapshots (the oracle is
pinned at full size, not only on the small decks) before it writes the fixture.

Kept per snapshot (steps 64 and 128): a fixed sample of cells (every stream cell + a strided sample
of the hillslopes) of the state fields, sums over blocks of 1000 consecutive cells of every field
(full coverage at block resolution) and the hydrograph arrays.

    python tests/golden/make_fullsize.py            # ~12 min, writes tests/golden/fullsize_1m.npz"""
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

SIDE, SEED, SNAP_STEPS, RUNTIME_H = 1000, 1, (64, 128), 24.0   # RUNTIME only sizes the hydrograph arrays
# fixture name -> name in the reference's post-routing dump (state before Reset = Old state after it)
STATE = {"NwtOld": "NwtNew", "MuOld": "MuNew", "MiOld": "MiNew", "NtOld": "NtNew", "NfOld": "NfNew", "RuOld": "RuNew",
         "RiOld": "RiNew", "Qpout": "Qpout", "cumsrf": "cumsrf", "intstorm": "intstorm"}
RESULTS = ["mhydro", "HsrfRout", "SbsrfRout", "SatsrfRout"]
BLOCK = 1000
NOT_PATCHED = ("aspect",)        # kernel (1) input; this deck runs without the evaporation scheme


def sample_cells(basin):
    n = int(basin["n_active"][0])
    stream = np.nonzero(np.asarray(basin["boundary"]) == 3)[0]
    strided = np.arange(0, n, 251)
    return np.unique(np.concatenate([stream, strided])).astype(np.int64)


def block_sums(a):
    n = (a.size // BLOCK) * BLOCK
    return a[:n].reshape(-1, BLOCK).sum(axis=1)


def reference_run(side, workdir, snap_steps, seed=SEED, runtime_h=RUNTIME_H):
    """build_basin -> OPTMESHINPUT 9 deck -> the reference. Returns (build_basin dict, the reference's
    own flattened basin, its snapshots, its timing line)."""
    import json
    from refrun import HARNESS, write_meshb_deck
    from tribs_b200.trb import read_trb
    basin = write_meshb_deck(side, workdir, seed=seed, runtime_h=runtime_h)
    steps = sorted(snap_steps)
    every = steps[1] - steps[0] if len(steps) > 1 else 1
    assert all(b - a == every for a, b in zip(steps, steps[1:]))
    cmd = [HARNESS, "basin.in", "-K", "--out", workdir, "--snap-from", str(steps[0]), "--snap-to", str(steps[-1]),
           "--snap-every", str(every), "--phases", "r", "--max-steps", str(steps[-1])]
    r = subprocess.run(cmd, cwd=workdir, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference harness failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    m = re.search(r"ref_timing (\{.*\})", r.stdout)
    return basin, read_trb(os.path.join(workdir, "basin.trb")), read_trb(os.path.join(workdir, "steps.trb")), \
        (json.loads(m.group(1)) if m else {})


def make_patches(basin, ref):
    """Where the reference's flattened basin differs from build_basin's dict: small arrays whole,
    large ones as (index, value)."""
    z = {}
    for k, a in basin.items():
        if k in NOT_PATCHED:
            continue
        a, b = np.asarray(a), np.asarray(ref[k])
        assert a.shape == b.shape, (k, a.shape, b.shape)
        if np.array_equal(a, b):
            continue
        if a.size <= 16:
            z[f"patch/full/{k}"] = b.copy()
        else:
            idx = np.flatnonzero(a != b)
            assert idx.size <= max(0.02 * a.size, 8 * a.size ** 0.5), (k, idx.size, a.size)   # sparse: hull / stream cells only
            z[f"patch/idx/{k}"], z[f"patch/val/{k}"] = idx.astype(np.int64), b[idx].copy()
    return z


def apply_patches(basin, z):
    out = dict(basin)
    for key in z.files if hasattr(z, "files") else z:
        if key.startswith("patch/full/"):
            out[key[11:]] = np.array(z[key])
        elif key.startswith("patch/idx/"):
            k = key[10:]
            a = np.array(out[k])
            a[z[key]] = z[f"patch/val/{k}"]
            out[k] = a
    return out


def main(side=SIDE, out=None, check_oracle=True):
    from bench import storm_rain
    from common import OracleStepper
    t0 = time.time()
    work = tempfile.mkdtemp(prefix="tribs_fullsize_")
    basin, ref_basin, snaps, timing = reference_run(side, work, SNAP_STEPS)
    print(f"reference run: {time.time() - t0:.0f} s  {timing}", flush=True)
    z = {"side": np.array([side]), "seed": np.array([SEED]), "snap_steps": np.array(SNAP_STEPS),
         "n_active": np.array([int(basin["n_active"][0])]), "ref_cell_steps_per_s": np.array([timing.get("cell_steps_per_s", 0.0)]),
         "ref_setup_s": np.array([timing.get("setup_s", 0.0)]), "ref_loop_s": np.array([timing.get("loop_s", 0.0)])}
    z.update(make_patches(basin, ref_basin))
    patched = apply_patches(basin, z)
    for k in basin:
        if k not in NOT_PATCHED:
            assert np.array_equal(np.asarray(patched[k]), np.asarray(ref_basin[k])), k
    idx = sample_cells(patched)
    z["cells"] = idx
    for st in SNAP_STEPS:
        for f, src in STATE.items():
            a = snaps[f"s{st}_r_{src}"]
            z[f"s{st}/{f}"] = a[idx].copy()
            z[f"s{st}/blk/{f}"] = block_sums(a)
        for r in RESULTS:
            z[f"s{st}/res/{r}"] = np.array(snaps[f"s{st}_r_{r}"]).copy()
    bitexact = -1
    if check_oracle:      # the oracle on the patched inputs against the reference's FULL arrays
        o = OracleStepper(patched, storm_rain)
        bitexact = 1
        for st in range(1, max(SNAP_STEPS) + 1):
            o.advance(); o.unsat()
            if o.gw_due():
                o.sat()
            o.route(); o.balance()
            if st in SNAP_STEPS:
                got = o.get(list(STATE.values()))
                for f, src in STATE.items():
                    same = np.array_equal(got[src], snaps[f"s{st}_r_{src}"])
                    bitexact &= int(same)
                    if not same:
                        e = np.abs(got[src] - snaps[f"s{st}_r_{src}"])
                        print(f"  step {st} {src}: oracle differs from the reference in {int((e > 0).sum())} cells, max {e.max():.3e}")
                for r in RESULTS:
                    bitexact &= int(np.array_equal(np.asarray(o.result(r)), snaps[f"s{st}_r_{r}"]))
                print(f"step {st}: oracle == reference bit for bit: {bool(bitexact)}  ({time.time() - t0:.0f} s)", flush=True)
            o.reset()
        assert bitexact == 1, "the oracle does not reproduce the reference on this basin"
    z["oracle_bitexact"] = np.array([bitexact])
    out = out or os.path.join(HERE, f"fullsize_{side // 1000}m.npz" if side >= 1000 else f"fullsize_{side}.npz")
    np.savez_compressed(out, **z)
    import shutil
    shutil.rmtree(work, ignore_errors=True)
    print("wrote", out, os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else SIDE)
