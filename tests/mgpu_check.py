"""Run under torchrun with 2+ ranks on one node: the ranks advance one basin together (reach partitions,
peer-memory halos, pipelined batches) and the result must equal the one-GPU run bit for bit - every owned
cell value, the hydrograph arrays, the basin accumulators and the per-step lateral inflows. The check itself
is bench.self_check, the one bench.py --gpus N runs before it times anything."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    side, steps = int(sys.argv[1]), int(sys.argv[2])
    info = bench.self_check(rank, world, steps=steps, side=side)      # exits non-zero on any difference
    if rank == 0:
        print(f"mgpu_check side={side} steps={steps} world={world}: "
              f"{'BITWISE EQUAL' if info['bitwise_equal_to_one_gpu'] else 'DIFFERENT'}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
