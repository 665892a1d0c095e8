"""One reach partition per GPU: host-side orchestration of the halo exchange (torch.distributed,
NCCL on GPUs; the same code runs over gloo with CPU tensors in the tests of the index logic).

Per step and rank r (reference: the MPI call sites in tHydroModel.cpp:765-776, 2081-2095,
2194-2199, 3001-3014):
    sweep pass LOCAL                      cells whose upslope inputs are all here
    recv qpout halos from ranks < r       (the river runs from lower to higher ranks)
    sweep pass REMOTE                     cells downstream of another rank's cell
    send qpout halos to ranks > r
    [GW step] exchange nwt halos both ways, broadcast psrf_last, SAT
    ROUTE (owned cells only; basin arrays are per-rank partial sums)
    exchange state halos (downslope cells read next step), RESET
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import gpu
from .model import tRunTimer
from .partition import cell_owner, halo_plan


class _DevPtr:
    """Wraps raw device memory for torch via __cuda_array_interface__."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


class PartitionedSimulator:
    def __init__(self, basin: dict, rank: int, world: int, rain_schedule=None, owner=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = rank, world
        self.basin = basin
        self.owner = cell_owner(basin, world) if owner is None else np.asarray(owner, np.int32)
        self.plan = halo_plan(basin, self.owner, rank)
        self.dev = gpu.GpuBasin(basin)
        L = self.dev.L
        L.tribs_gpu_set_partition.argtypes = [C.c_void_p, gpu.PI, C.c_int]
        L.tribs_gpu_halo_pack.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.tribs_gpu_halo_unpack.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.tribs_gpu_set_global.argtypes = [C.c_void_p, C.c_int, C.c_double]
        own = np.ascontiguousarray(self.owner, np.int32)
        gpu._chk(L.tribs_gpu_set_partition(self.dev.h, own.ctypes.data_as(gpu.PI), rank))
        # everything (library kernels, pack/unpack, NCCL) is ordered on the library's stream
        self.stream = torch.cuda.ExternalStream(self.dev.stream_handle())
        torch.cuda.set_stream(self.stream)
        dmap = np.empty(self.dev.n, np.int32)
        gpu._chk(L.tribs_gpu_device_index(self.dev.h, dmap.ctypes.data_as(gpu.PI)))
        self.dmap = dmap
        self.idx, self.buf = {}, {}
        width = {"qpout": 1, "state": 7, "nwt": 1}
        for key, cells in ((k, v) for k in self.plan if k != "owned" for v in [self.plan[k]]):
            kind = key.split("_")[0]
            for peer, c in cells.items():
                self.idx[(key, peer)] = torch.as_tensor(dmap[c].astype(np.int32), device="cuda")
                self.buf[(key, peer)] = torch.empty(width[kind] * c.size, dtype=torch.float64, device="cuda")
        b = basin
        self.timer = tRunTimer(float(b["tstep"][0]), float(b["gwstep"][0]), float(b["etistep"][0]), float(b["endtime"][0]))
        self.rain_schedule = rain_schedule
        self.gw_on = bool(int(b["GW_model_label"][0]))
        # kernel (1) is cell-local: every rank runs it on the replicated arrays, owned cells are exact
        from .model import tEvapoTrans, tIntercept
        from . import met
        self.intercept = tIntercept(b)
        self.evapotrans = tEvapoTrans(self.dev, self.timer, b)
        self.met_clock = met.MetClock(self.timer.tstep, float(b["metstep"][0]) if "metstep" in b else self.timer.etistep,
                                      self.timer.etistep)
        last_owner = int(self.owner[-1])          # owner of the last swept cell (psrf_last)
        self.psrf_owner = last_owner

    KIND = {"qpout": 0, "state": 1, "nwt": 2}

    def _pack(self, key, peer):
        kind = self.KIND[key.split("_")[0]]
        idx, buf = self.idx[(key, peer)], self.buf[(key, peer)]
        gpu._chk(self.dev.L.tribs_gpu_halo_pack(self.dev.h, kind, idx.data_ptr(), idx.numel(), buf.data_ptr()))
        return buf

    def _unpack(self, key, peer):
        kind = self.KIND[key.split("_")[0]]
        idx, buf = self.idx[(key, peer)], self.buf[(key, peer)]
        gpu._chk(self.dev.L.tribs_gpu_halo_unpack(self.dev.h, kind, idx.data_ptr(), idx.numel(), buf.data_ptr()))

    def _exchange_both_ways(self, what):
        dist = self.dist
        ops = []
        for peer in sorted(self.plan[what + "_send"]):
            ops.append(dist.P2POp(dist.isend, self._pack(what + "_send", peer), peer))
        for peer in sorted(self.plan[what + "_recv"]):
            ops.append(dist.P2POp(dist.irecv, self.buf[(what + "_recv", peer)], peer))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for peer in sorted(self.plan[what + "_recv"]):
            self._unpack(what + "_recv", peer)

    def one_step(self, rain_cells=None, fetch=False):
        """rain_cells: optional host array [n] (per-cell rain of this step, sweep order);
        fetch=True also reads back what the host model consumes each step."""
        t, dev, dist = self.timer, self.dev, self.dist
        t.Advance(t.getTimeStep())
        now = t.getCurrentTime()
        el = t.getElapsedSteps(now)
        hr = int(not math.fmod(now - t.getTimeStep(), t.etistep))
        rain = self.rain_schedule(now) if self.rain_schedule is not None else None
        if rain_cells is not None:
            rain_cells.fill(0.0 if rain is None else rain)
            rain = rain_cells
        if self.evapotrans.getEToption() != 0:                         # Simulator::SurfaceHydroProcesses
            met_due, eti_due = self.met_clock.due(now)
            if met_due or eti_due:
                if rain is not None:
                    dev.step(0, now, el, 0, rain=rain)
                    rain = None
                if met_due:
                    self.evapotrans.callEvapoPotential()
                if eti_due:
                    self.evapotrans.callEvapoTrans(self.intercept, 1 if self.intercept.getIoption() != 0 else 0)
                self.evapotrans.flush()
        dev.step(16, now, el, hr, rain=rain)                          # TRIBS_PHASE_UNSAT_LOCAL
        for peer in sorted(self.plan["qpout_recv"]):                   # ranks upstream of me
            dist.recv(self.buf[("qpout_recv", peer)], src=peer)
            self._unpack("qpout_recv", peer)
        dev.step(32, now, el, hr)                                      # TRIBS_PHASE_UNSAT_REMOTE
        for peer in sorted(self.plan["qpout_send"]):
            dist.send(self._pack("qpout_send", peer), dst=peer)
        did_gw = self.gw_on and not math.fmod(now, t.getGWTimeStep())
        if did_gw:
            self._exchange_both_ways("nwt")
            v = self.torch.zeros(1, dtype=self.torch.float64, device="cuda")
            if self.rank == self.psrf_owner:
                v[0] = dev.globals()["psrf_last"]
            dist.broadcast(v, src=self.psrf_owner)
            gpu._chk(dev.L.tribs_gpu_set_global(dev.h, gpu.GLOBALS.index("psrf_last"), float(v.item())))
            dev.step(gpu.PHASE_SAT, now, el, hr)
        dev.step(gpu.PHASE_ROUTE, now, el, 0)
        out = (dev.retrieve_qeff(now), dev.globals()) if fetch else None
        self._exchange_both_ways("state")
        dev.step(gpu.PHASE_RESET, now, el, hr)
        return out if fetch else did_gw

    def gather(self, names):
        """Owned values of the requested fields from every rank, assembled on every rank."""
        main = [nm for nm in names if nm not in gpu.EF]
        local = self.dev.sync(main) if main else {}
        et = [nm for nm in names if nm in gpu.EF]
        if et:
            local.update(self.dev.et_sync(et))
        out = {}
        for nm in names:
            a = self.torch.as_tensor(np.where(self.owner == self.rank, local[nm], 0.0), device="cuda")
            self.dist.all_reduce(a)     # exactly one rank contributes a non-zero entry per cell
            out[nm] = a.cpu().numpy()
        return out
