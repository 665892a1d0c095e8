"""Reach partitions advanced together in pipelined batches, boundary values moved by the kernels
themselves through peer memory (tribs_part_t with nranks > 1; see include/tribs_gpu.h).

Two hosts for the same device protocol:
  PeerSimulator   one process per GPU (torchrun): the 64-byte memory handles of the ranks are gathered with
                  torch.distributed, which is all the host ever exchanges; what the reference's tParallel does
                  with MPI_Allgather.
  LocalPeers      all ranks in one process (several handles, one or more devices): the same kernels, the same
                  remote stores and system-scope atomics; used by the tests that have to run on a one-GPU box.
Both give results bit-identical to one GPU running the same partition layout (Simulator(part=dict(nranks=1, ...))).
Each rank is a model.Simulator on its own handle: same timer arithmetic, same forcing and ET scheduling."""
from __future__ import annotations

import copy

import numpy as np

from . import gpu
from .model import Simulator
from .partition import cell_owner


class PeerSimulator(Simulator):
    """One rank of an nranks-GPU run (one process per GPU)."""

    def __init__(self, basin: dict, rank: int, world: int, rain_schedule=None, owner=None, max_batch_steps: int = 64,
                 cell_gauge=None):
        import torch.distributed as dist
        self.rank, self.world = rank, world
        self.owner = cell_owner(basin, world) if owner is None else np.asarray(owner, np.int32)
        super().__init__(basin, rain_schedule, part=dict(rank=rank, nranks=world, cell_owner=self.owner,
                                                         max_batch_steps=max_batch_steps))
        if cell_gauge is not None:
            self.dev.set_rain_gauges(cell_gauge)
        handles = [None] * world
        dist.all_gather_object(handles, self.dev.ipc_export())
        self.dev.attach_peers(handles)

    def owned(self):
        return self.owner == self.rank

    def gather(self, names):
        """Owned values of the requested fields from every rank, assembled on every rank."""
        import torch
        import torch.distributed as dist
        main = [nm for nm in names if nm in gpu.F]
        local = self.dev.sync(main) if main else {}
        et = [nm for nm in names if nm in gpu.EF and nm not in gpu.F]
        if et:
            local.update(self.dev.et_sync(et))
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
        self.sims = []
        for r in range(world):
            if devices is not None:
                import torch
                torch.cuda.set_device(devices[r])
            # a schedule may carry state (met.GaugeRain's position in its series): every rank steps its own copy
            s = Simulator(basin, copy.deepcopy(rain_schedule), part=dict(rank=r, nranks=world, cell_owner=self.owner,
                                                                         max_batch_steps=max_batch_steps))
            if cell_gauge is not None:
                s.dev.set_rain_gauges(cell_gauge)
            s.flow.defer = True       # channel routing waits until every rank of this process has launched its batch
            self.sims.append(s)
        self.devs = [s.dev for s in self.sims]
        gpu.GpuBasin.attach_local(self.devs)

    def run_batch(self, k: int, rain_of_time=None):
        for s in self.sims:      # every call returns without waiting: the ranks meet on the device
            s.run_batch(k, rain_of_time=rain_of_time)
        for d in self.devs:
            d.wait()
        for s in self.sims:      # channel routing (if on): every rank routes the whole net from the combined inflows
            pending, s.flow.pending_steps = s.flow.pending_steps, []
            if len(pending) > 1:
                raise RuntimeError("LocalPeers: a batch that was cut into several launches cannot route its channels afterwards")
            for k in pending:
                s.flow.channels_after_batch(k, deferred=True)

    def gather(self, names):
        out = {nm: np.zeros(self.devs[0].n) for nm in names}
        for r, d in enumerate(self.devs):
            main = [nm for nm in names if nm in gpu.F]
            local = d.sync(main) if main else {}
            et = [nm for nm in names if nm in gpu.EF and nm not in gpu.F]
            if et:
                local.update(d.et_sync(et))
            m = self.owner == r
            for nm in names:
                out[nm][m] = local[nm][m]
        return out

    def close(self):
        for d in self.devs:
            d.close()
