"""Generates tests/golden/*.npz from the reference itself (run in the build container, where
/root/reference exists and oracle/_ref/tribs_ref_harness has been built by `make -C oracle ref`).

    python tests/golden/make_golden.py

Each fixture holds the flattened basin ("b/..."), raw FP64 snapshots of the reference's tCNode
fields after selected phases/steps ("s/s<step>_<phase>_<field>") and the known-answer vectors of
the reference's scalar helpers ("k/...")."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from refrun import SCENARIOS, run_reference, spec_for  # noqa: E402

N = 10                      # 10x10 points -> 64 active cells
KEEP_STEPS_PER_H = (1, 2, 8, 9)   # steps within selected hours
DROP = ("NetPrecipitation", "EvapSoil", "EvapDryCanopy", "EvapoTrans", "UnSaturatedStorage", "SaturatedStorage",
        "Hlevel", "Qstrm")


def main():
    for name in SCENARIOS:
        spec = spec_for(name, N)
        nsteps = int(round(spec.runtime_h * 60.0 / spec.timestep_min))
        basin, snaps, timing, kat = run_reference(spec, snap_from=1, snap_to=nsteps, phases="mugrb" if spec.forcing == "met" else "ugrb", kat=True)
        keep = set()
        for h in list(range(0, int(spec.runtime_h), 8)) + [int(spec.runtime_h) - 1]:
            for k in KEEP_STEPS_PER_H:
                keep.add(h * 16 + k)
        keep.add(nsteps)
        keep = sorted(s for s in keep if 1 <= s <= nsteps)
        out = {"b/" + k: v for k, v in basin.items()}
        for k, v in snaps.items():
            if k == "snap_steps":
                continue
            step = int(k.split("_")[0][1:])
            field = k.split("_", 2)[2]
            phase = k.split("_")[1]
            if phase == "g" or (phase == "u" and step % 8 != 0):
                continue   # 'r' carries the post-step state; 'u' is kept on GW steps to separate the zones
            if step in keep and (phase in ("m", "b") or field not in DROP):
                out["s/" + k] = v
        out["s/snap_steps"] = np.array(keep, np.int32)
        out["s/n_steps"] = np.array([nsteps], np.int32)
        for k, v in kat.items():
            out["k/" + k] = v
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "steps kept", len(keep), "size", os.path.getsize(path) // 1024, "KiB", timing)


if __name__ == "__main__":
    main()
