"""Golden vectors of the kinematic-wave channel router (tKinemat::RunRoutingModel, src/tFlowNet/tKinemat.cpp:803-897),
made by the REFERENCE ITSELF (oracle/_ref/tribs_ref_harness):

    python tests/golden/make_channel_golden.py

Two basins under the "shallow" storm deck (saturation-excess runoff reaches the channels within the first hour):
one reach of 11 nodes, and a dendritic net of 13 reaches (tributaries every 4 rows; confluences are the heads of the
next reaches). Each fixture holds the flattened basin ("b/...": enough to run the whole model) and, for EVERY step,
what the channel router left: outlet discharge tKinemat::Qout, water level / discharge / velocity of every stream node,
tKinemat::OutletHlev per reach and the basin outlet node ("c/...")."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from refrun import run_reference, spec_for  # noqa: E402

CASES = {"channel_single": dict(n=12, steps=400, tributary_every=0),
         "channel_dendritic": dict(n=20, steps=600, tributary_every=4)}


def main():
    for name, case in CASES.items():
        steps = case["steps"]
        spec = spec_for("shallow", case["n"], runtime_h=steps * 0.0625 + 1, tributary_every=case["tributary_every"])
        basin, snaps, timing, _ = run_reference(spec, snap_from=1, snap_to=steps, snap_every=1, phases="r")
        stream = np.nonzero(np.asarray(basin["boundary"]) == 3)[0]
        out = {"b/" + k: v for k, v in basin.items()}
        out["c/n_steps"] = np.array([steps], np.int32)
        out["c/stream_cells"] = stream.astype(np.int32)
        out["c/qout"] = np.array([snaps[f"s{s}_r_globals"][6] for s in range(1, steps + 1)])
        for f in ("Hlevel", "Qstrm", "FlowVelocity"):
            out["c/" + f] = np.stack([snaps[f"s{s}_r_{f}"][stream] for s in range(1, steps + 1)])
        out["c/OutletHlev"] = np.stack([snaps[f"s{s}_r_OutletHlev"] for s in range(1, steps + 1)])
        for f in ("outlet_Hlevel", "outlet_Qstrm"):
            out["c/" + f] = np.array([snaps[f"s{s}_r_{f}"][0] for s in range(1, steps + 1)])
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "reaches", len(basin["reach_head"]), "peak Qout", out["c/qout"].max(), "size", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
