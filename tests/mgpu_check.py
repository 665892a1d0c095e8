"""Run under torchrun with 2+ ranks on one node: partitioned run == single-GPU run, bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from refrun import load_golden, schedule_for  # noqa: E402
from tribs_b200.model import Simulator  # noqa: E402
from tribs_b200.parallel import PartitionedSimulator  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    name, steps = sys.argv[1], int(sys.argv[2])
    basin, _, _ = load_golden(name)
    ps = PartitionedSimulator(basin, rank, world, schedule_for(basin))
    for _ in range(steps):
        ps.one_step()
    fields = ["NwtOld", "MuOld", "MiOld", "NtOld", "NfOld", "RuOld", "RiOld", "Qpout", "cumsrf", "srf_hr", "RechDisch",
              "Transmissivity", "AvSoilMoisture", "esrf", "UnSaturatedStorage", "SaturatedStorage"]
    if name == "met":
        fields += ["PotEvap", "SurfTemp", "SoilTemp", "CanStorage", "CumIntercept", "EvapSoil", "EvapDryCanopy",
                   "NetPrecipitation", "CumTotEvap", "CanopyStorVol"]
    got = ps.gather(fields)
    qeff = ps.dev.retrieve_qeff(0.0)
    ok = True
    if rank == 0:
        ref = Simulator(basin, schedule_for(basin))
        for _ in range(steps):
            ref.one_step()
        from tribs_b200.gpu import EF
        want = ref.dev.sync([f for f in fields if f not in EF])
        if any(f in EF for f in fields):
            want.update(ref.dev.et_sync([f for f in fields if f in EF]))
        for f in fields:
            if not np.array_equal(want[f], got[f]):
                k = int(np.argmax(np.abs(want[f] - got[f])))
                print(f"MISMATCH {f} cell {k}: 1-GPU {want[f][k]!r} vs {world}-GPU {got[f][k]!r}", flush=True)
                ok = False
        print(f"mgpu_check {name} steps={steps} world={world}: {'BITWISE EQUAL' if ok else 'DIFFERENT'}", flush=True)
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.broadcast(t, src=0)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
