"""Reach partitions advanced together in pipelined batches, boundary values moved by the kernels
themselves through peer memory (tribs_part_t with nranks > 1; see include/tribs_gpu.h).

Two hosts for the same device protocol:
  PeerSimulator   one process per GPU (torchrun): the 64-byte memory handles of the ranks are gathered with
                  torch.distributed, which is all the host ever exchanges; what the reference's tParallel does
                  with MPI_Allgather.
  LocalPeers      all ranks in one process (several handles, one or more devices): the same kernels, the same
                  remote stores and system-scope atomics; used by the tests that have to run on a one-GPU box.
Both give results bit-identical to one GPU running the same partition layout (GpuBasin(part=dict(nranks=1, ...)))."""
from __future__ import annotations

import math

import numpy as np

from . import gpu
from .model import tRunTimer
from .partition import cell_owner


def _step_dicts(timer: tRunTimer, k: int, gw_on: bool, rain_of_time):
    out = []
    for _ in range(k):
        timer.Advance(timer.getTimeStep())
        now = timer.getCurrentTime()
        out.append(dict(current_time=now, elapsed_steps=timer.getElapsedSteps(now),
                        hourly_reset=int(not math.fmod(now - timer.getTimeStep(), timer.etistep)),
                        sat=bool(gw_on and not math.fmod(now, timer.getGWTimeStep())),
                        rain=rain_of_time(now) if rain_of_time is not None else None))
    return out


def _timer(b):
    return tRunTimer(float(b["tstep"][0]), float(b["gwstep"][0]), float(b["etistep"][0]), float(b["endtime"][0]))


class PeerSimulator:
    """One rank of an nranks-GPU run (one process per GPU)."""

    def __init__(self, basin: dict, rank: int, world: int, rain_schedule=None, owner=None, max_batch_steps: int = 64,
                 cell_gauge=None):
        import torch.distributed as dist
        self.rank, self.world = rank, world
        self.owner = cell_owner(basin, world) if owner is None else np.asarray(owner, np.int32)
        self.dev = gpu.GpuBasin(basin, part=dict(rank=rank, nranks=world, cell_owner=self.owner,
                                                 max_batch_steps=max_batch_steps))
        if cell_gauge is not None:
            self.dev.set_rain_gauges(cell_gauge)
        handles = [None] * world
        dist.all_gather_object(handles, self.dev.ipc_export())
        self.dev.attach_peers(handles)
        self.timer = _timer(basin)
        self.rain_schedule = rain_schedule
        self.gw_on = bool(int(basin["GW_model_label"][0]))

    def run_batch(self, k: int, rain_of_time=None):
        steps = _step_dicts(self.timer, k, self.gw_on, rain_of_time or self.rain_schedule)
        self.dev.run(steps)
        return steps

    def owned(self):
        return self.owner == self.rank

    def gather(self, names):
        """Owned values of the requested fields from every rank, assembled on every rank."""
        import torch
        import torch.distributed as dist
        local = self.dev.sync(names)
        out = {}
        for nm in names:
            a = torch.as_tensor(np.where(self.owned(), local[nm], 0.0), device="cuda")
            dist.all_reduce(a)     # exactly one rank contributes a non-zero entry per cell
            out[nm] = a.cpu().numpy()
        return out


class LocalPeers:
    """All ranks of an nranks-partition run as handles of this process."""

    def __init__(self, basin: dict, world: int, rain_schedule=None, owner=None, max_batch_steps: int = 64, cell_gauge=None,
                 devices=None):
        self.world = world
        self.owner = cell_owner(basin, world) if owner is None else np.asarray(owner, np.int32)
        self.devs = []
        for r in range(world):
            if devices is not None:
                import torch
                torch.cuda.set_device(devices[r])
            d = gpu.GpuBasin(basin, part=dict(rank=r, nranks=world, cell_owner=self.owner, max_batch_steps=max_batch_steps))
            if cell_gauge is not None:
                d.set_rain_gauges(cell_gauge)
            self.devs.append(d)
        gpu.GpuBasin.attach_local(self.devs)
        self.timer = _timer(basin)
        self.rain_schedule = rain_schedule
        self.gw_on = bool(int(basin["GW_model_label"][0]))

    def run_batch(self, k: int, rain_of_time=None):
        steps = _step_dicts(self.timer, k, self.gw_on, rain_of_time or self.rain_schedule)
        for d in self.devs:      # every call returns without waiting: the ranks meet on the device
            d.run(steps)
        for d in self.devs:
            d.wait()
        return steps

    def gather(self, names):
        out = {nm: np.zeros(self.devs[0].n) for nm in names}
        for r, d in enumerate(self.devs):
            local = d.sync(names)
            m = self.owner == r
            for nm in names:
                out[nm][m] = local[nm][m]
        return out

    def close(self):
        for d in self.devs:
            d.close()
