"""CPU tests of the N>1 path: reach partition, halo index sets, and the exchange protocol run
over gloo with world_size 2 (CPU tensors carry global cell ids instead of physics)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from refrun import load_golden
from tribs_b200.partition import cell_owner, cell_reaches, default_partition, halo_plan


def test_default_partition_matches_reference_arithmetic():
    # tGraph::createDefaultPartition (src/tGraph/tGraph.cpp:588-608): ceil-sized leading blocks
    assert default_partition(10, 3).tolist() == [0, 0, 0, 0, 1, 1, 1, 2, 2, 2]
    assert default_partition(7, 4).tolist() == [0, 0, 1, 1, 2, 2, 3]
    assert default_partition(4, 4).tolist() == [0, 1, 2, 3]
    with pytest.raises(ValueError):
        default_partition(2, 3)


@pytest.mark.parametrize("name", ["deep", "shallow"])
@pytest.mark.parametrize("world", [2, 3])
def test_plans_are_closed_and_symmetric(name, world):
    b, _, _ = load_golden(name)
    owner = cell_owner(b, world)
    bnd, sn, fd = b["boundary"], b["stream_node"], b["flow_dest"]
    hill = np.flatnonzero((bnd == 0) & (sn >= 0))
    assert np.array_equal(owner[hill], owner[sn[hill]]), "a hillslope cell must live with its stream node"
    # hillslope sweeps are partition-local: only stream cells cross a cut along the flow
    cut = np.flatnonzero((fd >= 0) & (owner != owner[np.maximum(fd, 0)]))
    assert np.all(bnd[cut] == 3) and np.all(bnd[fd[cut]] == 3)
    plans = [halo_plan(b, owner, r) for r in range(world)]
    for a in range(world):
        for kind in ("qpout", "state", "nwt"):
            for peer, cells in plans[a][kind + "_send"].items():
                assert np.array_equal(cells, plans[peer][kind + "_recv"][a]), (kind, a, peer)
                assert np.all(owner[cells] == a)
    assert sum(p["owned"].size for p in plans) == owner.size


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _bench_like_basin(world):
    """The basin shape bench.py uses for N GPUs: a valley N times as long as it is wide."""
    from tribs_b200.synthbasin import build_basin
    return build_basin(24, 24 * world, seed=1, runtime_h=1.0)


@pytest.mark.parametrize("world", [4, 8])
def test_plans_of_the_multi_gpu_bench_basin(world):
    b = _bench_like_basin(world)
    owner = cell_owner(b, world)
    assert set(np.unique(owner)) == set(range(world))
    plans = [halo_plan(b, owner, r) for r in range(world)]
    for a in range(world):
        for kind in ("qpout", "state", "nwt"):
            for peer, cells in plans[a][kind + "_send"].items():
                assert np.array_equal(cells, plans[peer][kind + "_recv"][a]), (kind, a, peer)
        # the river runs through the ranks in order: outflow only goes to higher ranks
        assert all(peer > a for peer in plans[a]["qpout_send"])
    assert sum(p["owned"].size for p in plans) == owner.size


def _worker(rank, world, port, name, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = _bench_like_basin(world) if name == "bench" else load_golden(name)[0]
    owner = cell_owner(b, world)
    plan = halo_plan(b, owner, rank)
    n = owner.size
    # every rank knows only the "values" (here: 1000*step + cell id) of the cells it owns
    field = np.where(owner == rank, np.arange(n, dtype=np.float64), np.nan)
    ok = True
    for step in range(3):
        vals = field + 1000.0 * step
        # river exchange: receive from lower ranks, then send to higher ranks (PartitionedSimulator.one_step)
        for peer in sorted(plan["qpout_recv"]):
            buf = torch.empty(plan["qpout_recv"][peer].size, dtype=torch.float64)
            dist.recv(buf, src=peer)
            ok &= np.array_equal(buf.numpy(), plan["qpout_recv"][peer] + 1000.0 * step)
        for peer in sorted(plan["qpout_send"]):
            dist.send(torch.from_numpy(vals[plan["qpout_send"][peer]].copy()), dst=peer)
        for kind in ("nwt", "state"):
            ops, rbuf = [], {}
            for peer in sorted(plan[kind + "_send"]):
                ops.append(dist.P2POp(dist.isend, torch.from_numpy(vals[plan[kind + "_send"][peer]].copy()), peer))
            for peer in sorted(plan[kind + "_recv"]):
                rbuf[peer] = torch.empty(plan[kind + "_recv"][peer].size, dtype=torch.float64)
                ops.append(dist.P2POp(dist.irecv, rbuf[peer], peer))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            for peer, buf in rbuf.items():
                ok &= np.array_equal(buf.numpy(), plan[kind + "_recv"][peer] + 1000.0 * step)
    t = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(float(t.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,name", [(2, "deep"), (4, "bench")])
def test_halo_exchange_protocol_over_gloo(world, name):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 1.0


def test_single_river_basin_is_cut_into_reaches():
    b, _, _ = load_golden("deep")
    r = cell_reaches(b, 2)
    assert np.unique(r).size >= 2


def test_partition_equals_the_references_tgraph_on_a_multi_reach_basin():
    """CSR / partition indices bit-exact (north_star): the reference itself (tGraph::connectivity +
    tGraph::createDefaultPartition, src/tGraph/tGraph.cpp:297-343, 588-608, called by the harness on its own
    tKinemat reach lists) partitions a basin with 13 reaches into 2..8 parts; partition.py must give the same
    reach -> partition table and, through tCNode::getReach(), the same owner for every cell."""
    from refrun import have_harness, run_reference, spec_for
    if not have_harness():
        pytest.skip("oracle/_ref/tribs_ref_harness not built")
    spec = spec_for("deep", 40, runtime_h=1.0, seed=3, tributary_every=8)
    basin, _, _, _ = run_reference(spec, snap_from=1, snap_to=0)
    n_reach = int(basin["graph_n_reach"][0])
    assert n_reach == len(basin["reach_head"]) == np.unique(basin["reach"]).size >= 8
    for world in range(2, 9):
        ref_r2p = np.asarray(basin[f"graph_r2p_np{world}"])
        assert np.array_equal(default_partition(n_reach, world), ref_r2p), world
        assert np.array_equal(cell_owner(basin, world), ref_r2p[np.asarray(basin["reach"])]), world
    # and a partition never separates a hillslope cell from its stream node
    own = cell_owner(basin, 4)
    hill = np.flatnonzero((basin["boundary"] == 0) & (basin["stream_node"] >= 0))
    assert np.array_equal(own[hill], own[basin["stream_node"][hill]])
