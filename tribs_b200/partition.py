"""Reach partition and halo index sets for one-partition-per-GPU runs (pure numpy, no GPU).

Mirrors the reference's domain decomposition (src/tGraph/tGraph.cpp): reaches are assigned to
partitions in contiguous blocks (`createDefaultPartition`, 588-608); every hillslope cell
follows the reach of its stream node (tGraph::inWhichReach, 1771-1799); the cells on either
side of a cut are exchanged in ascending node-id order on both sides (the std::set<..> of
tGraph.h:43-48), which is what makes the two ends of a message agree without any handshake.

Halo kinds (the reference's MPI messages they replace):
  qpout  new lateral outflow of a cell whose downslope cell lives on another rank
         (sendQpin/receiveQpin, tGraph.cpp:2101-2155) -- needed *inside* the sweep;
  state  New state of a cell that a remote cell reads as its downslope cell
         (sendOverlap/receiveOverlap, 2035-2093: NwtOld/NfOld of remote reach heads);
  nwt    water table of Voronoi neighbours across the cut (sendNwt/receiveNwt, 2431-2479).
Because every rank holds the static data of the whole basin, the reference's second
groundwater exchange (un-summed flux lists, 2364-2423) is not needed: both directions of a cut
edge are evaluated locally from the exchanged water tables.
"""
from __future__ import annotations

import numpy as np


def default_partition(n_reach: int, n_parts: int) -> np.ndarray:
    """reach -> partition, contiguous blocks, as tGraph::createDefaultPartition."""
    if n_reach < n_parts:
        raise ValueError(f"{n_reach} reaches cannot be split over {n_parts} partitions")
    r2p = np.empty(n_reach, np.int32)
    rend, rp, nreach = -1, n_parts, n_reach
    for i in range(n_parts):
        rstart = rend + 1
        rend = n_reach - 1 if i == n_parts - 1 else rstart + (nreach + rp - 1) // rp - 1
        nreach -= rend - rstart + 1
        rp -= 1
        r2p[rstart:rend + 1] = i
    return r2p


def cell_reaches(basin: dict, min_reaches: int = 1) -> np.ndarray:
    """Reach id of every cell. Uses the reference's reach ids when there are enough of them;
    otherwise (e.g. a single-river synthetic basin) the stream cells are cut, in sweep order
    (upstream first), into 4*min_reaches stretches, which is what a reach is."""
    bnd = np.asarray(basin["boundary"])
    sn = np.asarray(basin["stream_node"])
    n = bnd.size
    reach = np.asarray(basin.get("reach", np.zeros(n, np.int32))).astype(np.int64)
    stream = np.flatnonzero(bnd == 3)
    if np.unique(reach[stream]).size < min_reaches:
        # Fewer reaches than partitions (a single-river synthetic basin is ONE reach in the reference, which then cannot
        # be partitioned at all): cut the river, in sweep order, into 4 * min_reaches stretches that drain about the same
        # number of cells each, so that blocks of stretches are balanced partitions.
        k = max(min_reaches * 4, 1)
        slot = np.full(n, -1, np.int64)
        slot[stream] = np.arange(stream.size)
        drains_to = np.where(bnd == 3, slot, np.where(sn >= 0, slot[np.maximum(sn, 0)], stream.size - 1))
        cells_of = np.bincount(drains_to[drains_to >= 0], minlength=stream.size)
        upto = np.cumsum(cells_of)
        reach = np.zeros(n, np.int64)
        reach[stream] = np.minimum((upto - cells_of) * k // max(int(upto[-1]), 1), k - 1)
    else:   # compact the ids in order of first appearance along the sweep (upstream first)
        _, first = np.unique(reach[stream], return_index=True)
        order = reach[stream][np.sort(first)]
        remap = {int(r): k for k, r in enumerate(order)}
        reach = np.array([remap.get(int(r), 0) for r in reach], np.int64)
    out = reach.copy()
    hill = bnd != 3
    has = hill & (sn >= 0)
    out[has] = reach[sn[has]]
    out[hill & (sn < 0)] = reach[stream].max() if stream.size else 0   # drains straight to the outlet
    return out


def cell_owner(basin: dict, n_ranks: int) -> np.ndarray:
    reach = cell_reaches(basin, n_ranks)
    r2p = default_partition(int(reach.max()) + 1, n_ranks)
    return r2p[reach].astype(np.int32)


def halo_plan(basin: dict, owner: np.ndarray, rank: int) -> dict:
    """Index sets (sweep indices, ascending) per peer for the three halo kinds."""
    owner = np.asarray(owner)
    fd = np.asarray(basin["flow_dest"])
    n = owner.size
    mine = owner == rank
    plan = {k: {} for k in ("qpout_send", "qpout_recv", "state_send", "state_recv", "nwt_send", "nwt_recv")}
    has = fd >= 0
    src = np.flatnonzero(has)
    dst = fd[src]
    cut = owner[src] != owner[dst]
    src, dst = src[cut], dst[cut]

    def group(cells, peers, key):
        for p in np.unique(peers):
            plan[key][int(p)] = np.unique(cells[peers == p]).astype(np.int32)

    m = owner[src] == rank          # my cell flows into a remote cell
    group(src[m], owner[dst[m]], "qpout_send")
    group(dst[m], owner[dst[m]], "state_recv")
    m = owner[dst] == rank          # a remote cell flows into mine
    group(src[m], owner[src[m]], "qpout_recv")
    group(dst[m], owner[src[m]], "state_send")
    rp = np.asarray(basin["spoke_rowptr"])
    col = np.asarray(basin["spoke_nbr"])
    row = np.repeat(np.arange(n), np.diff(rp))
    ok = col >= 0
    a, b = row[ok], col[ok]
    cutn = owner[a] != owner[b]
    a, b = a[cutn], b[cutn]
    m = owner[a] == rank
    group(b[m], owner[b[m]], "nwt_recv")
    group(a[m], owner[b[m]], "nwt_send")
    # the river must run from lower to higher ranks for the one-round exchange of the sweep
    bad = [p for p in plan["qpout_send"] if p < rank] + [p for p in plan["qpout_recv"] if p > rank]
    if bad:
        raise ValueError("partition is not ordered along the river (a higher rank drains into a lower one)")
    plan["owned"] = np.flatnonzero(mine).astype(np.int32)
    return plan
