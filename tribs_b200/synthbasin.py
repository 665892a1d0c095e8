"""Large synthetic basins, built directly in flattened form (no reference run needed).

The reference's own mesh/flow-network setup is super-linear and is out of scope
for the hot path, so for the 1M-50M-cell benchmark basins this module builds what that setup
would hand to `tribs_gpu_init`: a jittered-grid TIN of a V-shaped valley with its Voronoi
geometry, steepest-descent flow edges, stream column, sweep order (upslope cells first,
hillslope before stream: the contract of tFlowNet::SortNodesByNetOrder, tFlowNet.cpp:429-498),
hillslope/stream path lengths and per-cell parameters — the same dict layout the reference
harness dumps (basin.trb), so everything downstream is shared.

Triangulation: structured (each grid square split by the same diagonal), which equals the
Delaunay triangulation for the small jitter used; Voronoi vertices are the triangle
circumcentres, `vedglen` the distance between the two circumcentres flanking an edge and
`varea` the shoelace area of the cell polygon (reference: tMesh.cpp:3336-3404,
meshElements.cpp:858-867).
"""
from __future__ import annotations

import math

import numpy as np

# neighbour offsets (di, dj) counter-clockwise, consistent-diagonal triangulation
_NB = [(1, 0), (1, 1), (0, 1), (-1, 0), (-1, -1), (0, -1)]


def _circumcentre(ax, ay, bx, by, cx, cy):
    d = 2.0 * (ax * (by - cy) + bx * (cy - ay) + cx * (ay - by))
    a2, b2, c2 = ax * ax + ay * ay, bx * bx + by * by, cx * cx + cy * cy
    ux = (a2 * (by - cy) + b2 * (cy - ay) + c2 * (ay - by)) / d
    uy = (a2 * (cx - bx) + b2 * (ax - cx) + c2 * (bx - ax)) / d
    return ux, uy


def build_basin(nx: int, ny: int, spacing: float = 30.0, jitter: float = 3.0, seed: int = 1,
                slope_x: float = 0.05, slope_y: float = 0.02, stream_every: int = 0,
                gw_depth_mm: float = 3000.0, bedrock_mm: float = 15000.0,
                soil=(10.0, 0.45, 0.05, 0.4, -200.0, 0.001, 100.0, 100.0),
                tstep_h: float = 0.0625, gwstep_h: float = 0.5, runtime_h: float = 48.0,
                n_soil_classes: int = 1, mesh_dir: str | None = None) -> dict:
    """nx x ny active cells. One stream column along the valley axis (plus, with
    stream_every=k>0, a tributary stream row every k rows draining to it). Returns the
    flattened-basin dict in sweep order. With mesh_dir, the same basin is also written there as the
    reference's OPTMESHINPUT 9 files (meshb.py), so the reference itself can run it."""
    rng = np.random.default_rng(seed)
    # node grid with a one-node closed-boundary ring: indices 0..nx+1, 0..ny+1
    NX, NY = nx + 2, ny + 2
    ii, jj = np.meshgrid(np.arange(NX), np.arange(NY), indexing="xy")   # [NY, NX]
    x = ii * spacing
    y = jj * spacing
    interior = (ii >= 1) & (ii <= nx) & (jj >= 1) & (jj <= ny)
    ic = NX // 2
    stream = interior & (ii == ic)
    if stream_every:
        stream |= interior & (jj % stream_every == 0)
    jit = interior & ~stream
    x = x + np.where(jit, rng.uniform(-jitter, jitter, x.shape), 0.0)
    y = y + np.where(jit, rng.uniform(-jitter, jitter, y.shape), 0.0)
    xc = ic * spacing
    z = 100.0 + slope_y * y + slope_x * np.abs(x - xc)
    if stream_every:
        jm = jj % stream_every
        # local valleys towards the tributary rows; the main stem keeps its monotone profile
        z = z + np.where(ii == ic, 0.0, 0.03 * np.minimum(jm, stream_every - jm) * spacing)
    # outlet: boundary node below the bottom stream cell
    outlet_ij = (0, ic)

    def sh(a, di, dj):
        """a shifted so that result[j, i] = a[j+dj, i+di] (edges clamped; only interior is used)."""
        return np.roll(np.roll(a, -dj, axis=0), -di, axis=1)

    # circumcentres of the 6 triangles around every node: triangle k = (P, N_k, N_{k+1})
    nbx = [sh(x, di, dj) for di, dj in _NB]
    nby = [sh(y, di, dj) for di, dj in _NB]
    ccx, ccy = [], []
    for k in range(6):
        k2 = (k + 1) % 6
        ux, uy = _circumcentre(x, y, nbx[k], nby[k], nbx[k2], nby[k2])
        ccx.append(ux); ccy.append(uy)
    # Voronoi edge dual to spoke k joins circumcentres of triangles k-1 and k
    vedg = [np.hypot(ccx[k] - ccx[k - 1], ccy[k] - ccy[k - 1]) for k in range(6)]
    elen = [np.hypot(nbx[k] - x, nby[k] - y) for k in range(6)]
    area = np.zeros_like(x)
    for k in range(6):
        k2 = (k + 1) % 6
        area += ccx[k] * ccy[k2] - ccx[k2] * ccy[k]
    area = 0.5 * np.abs(area)

    # flow edge: steepest descent among the 6 neighbours; stream cells follow the stream
    nbz = [sh(z, di, dj) for di, dj in _NB]
    nb_stream = [sh(stream, di, dj) for di, dj in _NB]
    nb_inter = [sh(interior, di, dj) for di, dj in _NB]
    is_outlet = np.zeros_like(interior)
    is_outlet[outlet_ij] = True
    nb_outlet = [sh(is_outlet, di, dj) for di, dj in _NB]
    slopes = np.stack([(z - nbz[k]) / elen[k] for k in range(6)])            # [6, NY, NX]
    ok_hill = np.stack([nb_inter[k] for k in range(6)])
    ok_strm = np.stack([nb_stream[k] | nb_outlet[k] for k in range(6)])
    s_h = np.where(ok_hill, slopes, -np.inf)
    s_s = np.where(ok_strm, slopes, -np.inf)
    kbest = np.where(stream, np.argmax(s_s, axis=0), np.argmax(s_h, axis=0))
    best_slope = np.take_along_axis(np.where(stream, s_s, s_h), kbest[None], 0)[0]
    if not np.all(np.isfinite(best_slope[interior])):
        raise RuntimeError("synthetic basin: a cell has no admissible downslope neighbour")

    node_id = -np.ones((NY, NX), np.int64)
    flat_int = np.flatnonzero(interior.ravel())
    node_id.ravel()[flat_int] = np.arange(flat_int.size)
    n = flat_int.size
    di = np.array([d[0] for d in _NB]); dj = np.array([d[1] for d in _NB])
    I = ii.ravel()[flat_int]; J = jj.ravel()[flat_int]
    kb = kbest.ravel()[flat_int]
    dest_node = node_id[J + dj[kb], I + di[kb]]          # -1 = outlet
    is_stream = stream.ravel()[flat_int]

    flow_len = np.stack(elen)[kb, J, I]
    # Hops to the stream (hillslope cells) / to the outlet (stream cells) by pointer doubling: exact integers
    # in O(n log depth). The path lengths are then accumulated level by level, from the stream outwards,
    # in the association len_i + (len_{i+1} + ...) - the loops below visit every cell once.
    def hops_to_root(nxt):
        """nxt[i] = next cell of the same kind on the path, or -1; returns 1 + number of such successors."""
        d = np.where(nxt >= -1, 1, 0).astype(np.int64)
        p = nxt.copy()
        # pointer doubling reaches every root within ceil(log2 n) + 1 rounds; more means a cycle
        for _ in range(int(np.ceil(np.log2(max(nxt.size, 2)))) + 2):
            m = np.flatnonzero(p >= 0)
            if m.size == 0:
                return d
            d[m] += d[p[m]]
            p[m] = p[p[m]]
        raise RuntimeError("synthetic basin: flow graph has a cycle")

    def levels_of(depth_of, members):
        """members sorted by depth (stable), as a list of index arrays, one per level 1, 2, ..."""
        o = members[np.argsort(depth_of[members], kind="stable")]
        dv = depth_of[o]
        cuts = np.flatnonzero(np.diff(dv)) + 1
        return np.split(o, cuts)

    dn = np.maximum(dest_node, 0)
    hill = ~is_stream
    nxt_h = np.where(hill & (dest_node >= 0) & hill[dn], dest_node, -1)
    depth = np.where(hill, hops_to_root(nxt_h), 0)
    hillpath = np.zeros(n)
    stream_node = np.where(is_stream, np.arange(n), -1)
    for r in levels_of(depth, np.flatnonzero(hill)):
        dr = dest_node[r]
        to_outlet = dr < 0
        hillpath[r] = flow_len[r] + np.where(to_outlet, 0.0, hillpath[np.maximum(dr, 0)])
        stream_node[r] = np.where(to_outlet, -2, stream_node[np.maximum(dr, 0)])
    if (stream_node[hill] == -1).any():
        raise RuntimeError("synthetic basin: flow graph has a cycle")
    stream_node = np.where(stream_node == -2, -1, stream_node)
    # stream cells: distance to outlet along the stream, and order upstream -> downstream
    nxt_s = np.where(is_stream & (dest_node >= 0) & is_stream[dn], dest_node, -1)
    sdepth = np.where(is_stream, hops_to_root(nxt_s), 0)
    streampath_s = np.zeros(n)
    for r in levels_of(sdepth, np.flatnonzero(is_stream)):
        dr = dest_node[r]
        streampath_s[r] = flow_len[r] + np.where(dr < 0, 0.0, streampath_s[np.maximum(dr, 0)])
    streampath = np.where(is_stream, streampath_s, streampath_s[np.maximum(stream_node, 0)] * (stream_node >= 0))
    hillpath = np.where(is_stream, 0.0, hillpath)

    # sweep order: hillslope cells far from the stream first, then stream cells upstream first
    key = np.where(is_stream, 0, depth)
    skey = np.where(is_stream, sdepth, 0)
    order = np.lexsort((np.arange(n), -skey, -key, is_stream.astype(np.int64)))
    pos = np.empty(n, np.int64)
    pos[order] = np.arange(n)

    def remap(idx):
        return np.where(idx >= 0, pos[np.maximum(idx, 0)], -1).astype(np.int32)

    def take(a2d):
        return a2d.ravel()[flat_int][order]

    # contributing area of every cell: accumulate along the sweep order (upslope first)
    varea = take(area)
    fdest = remap(dest_node)[order]
    contr = varea.copy()
    # level-synchronous accumulation (levels of the sweep DAG)
    lvl = np.zeros(n, np.int64)
    src = np.arange(n)
    has = fdest >= 0
    # a cell's level = 1 + max level of its contributors; iterate in sweep order chunks by depth
    dd = np.where(is_stream[order], 10 ** 9 - sdepth[order], depth[order].max() + 1 - depth[order])
    o = np.argsort(dd, kind="stable")
    o = o[has[o]]
    for sel in np.split(o, np.flatnonzero(np.diff(dd[o])) + 1):
        np.add.at(contr, fdest[sel], contr[sel])

    # spokes CSR (CCW from offset 0), columns remapped to sweep order
    rowptr = np.arange(0, 6 * n + 1, 6, dtype=np.int32)
    In, Jn = I[order], J[order]
    nbr = np.empty((n, 6), np.int32)
    slen = np.empty((n, 6)); svl = np.empty((n, 6)); slen_in = np.empty((n, 6)); svl_in = np.empty((n, 6))
    for k in range(6):
        nid = node_id[Jn + dj[k], In + di[k]]
        nbr[:, k] = remap(nid)
        slen[:, k] = elen[k][Jn, In]
        svl[:, k] = vedg[k][Jn, In]
        # complementary edge = spoke (k+3)%6 of the neighbour
        k3 = (k + 3) % 6
        jn2 = np.clip(Jn + dj[k], 0, NY - 1); in2 = np.clip(In + di[k], 0, NX - 1)
        slen_in[:, k] = elen[k3][jn2, in2]
        svl_in[:, k] = vedg[k3][jn2, in2]
    # directed-edge ids and the order of the reference's edge list (tMesh::MakeMeshFromMeshBuilder keeps
    # file order, complementary edges adjacent): a pair belongs to its endpoint with the lower node id
    # (active cells: sweep position; then the outlet; then the closed ring), pairs are written owner by
    # owner, spoke by spoke; edges with a closed end, or between two open ends, are inactive
    gid = np.full(NX * NY, -1, np.int64)                  # node id per grid position
    gid[flat_int[order]] = np.arange(n)
    out_flat = outlet_ij[0] * NX + outlet_ij[1]
    gid[out_flat] = n
    ring = np.flatnonzero(gid < 0)
    gid[ring] = n + 1 + np.arange(ring.size)
    gI, gJ = ii.ravel(), jj.ravel()
    g_of_id = np.empty(NX * NY, np.int64)
    g_of_id[gid] = np.arange(NX * NY)
    exists = np.stack([(gI + di[k] >= 0) & (gI + di[k] < NX) & (gJ + dj[k] >= 0) & (gJ + dj[k] < NY) for k in range(6)], 1)
    nb_flat = np.stack([np.clip(gJ + dj[k], 0, NY - 1) * NX + np.clip(gI + di[k], 0, NX - 1) for k in range(6)], 1)
    nb_id = np.where(exists, gid[nb_flat], -1)
    owned = exists & (gid[:, None] < nb_id)                # [grid, k]
    owned_by_id = owned[g_of_id]                           # [id, k]
    pair_rank = np.cumsum(owned_by_id.ravel()) - 1         # file index of the pair, for owned (id, k)
    eid = np.full((NX * NY, 6), -1, np.int64)              # directed edge id per (grid, k)
    oid, ok_ = np.nonzero(owned_by_id)
    og = g_of_id[oid]
    pr = pair_rank.reshape(-1, 6)[oid, ok_]
    eid[og, ok_] = 2 * pr
    eid[nb_flat[og, ok_], (ok_ + 3) % 6] = 2 * pr + 1
    closed = np.ones(NX * NY, bool)
    closed[flat_int] = False
    closed[out_flat] = False
    is_open = np.zeros(NX * NY, bool)
    is_open[out_flat] = True
    pair_active = ~closed[og] & ~closed[nb_flat[og, ok_]] & ~(is_open[og] & is_open[nb_flat[og, ok_]])
    act_rank = np.cumsum(pair_active) - 1
    epos = np.full(2 * pr.size, -1, np.int64)              # position in the active edge list per edge id
    epos[2 * pr[pair_active]] = 2 * act_rank[pair_active]
    epos[2 * pr[pair_active] + 1] = 2 * act_rank[pair_active] + 1
    cell_g = flat_int[order]                               # grid position of every active cell, sweep order
    listpos = epos[eid[cell_g]].ravel().astype(np.int32)

    kbo = kb[order]
    flow_slope = best_slope.ravel()[flat_int][order]
    flow_vedglen = svl[np.arange(n), kbo]
    velcoef, velratio = 0.5, 5.0
    hillvel, streamvel = velcoef / velratio, velcoef
    hp, sp = hillpath[order], streampath[order]
    ttime = (hp / hillvel + sp / streamvel) / 3600.0
    bas_area = float(varea.sum())
    out_interval = 1.0
    maxtt = float(ttime.max())
    res_limit = int(np.ceil((runtime_h + maxtt) / out_interval)) + 2

    # aspect of a cell = direction of its flow edge, clockwise from north (tEvapoTrans::DeriveAspect, tEvapoTrans.cpp:2723-2756)
    xg, yg = x.ravel(), y.ravel()
    dest_g = ((J + dj[kb]) * NX + (I + di[kb]))[order]
    fx, fy, fl = xg[dest_g] - xg[cell_g], yg[dest_g] - yg[cell_g], slen[np.arange(n), kbo]
    aspect = np.arccos((fx * 0.0 + fy * fl) / (fl * fl))
    aspect = np.where(xg[dest_g] < xg[cell_g], 8.0 * math.atan(1.0) - aspect, aspect)

    ks, ths, thr, lam, psib, f, ar, uar = soil
    full = lambda v: np.full(n, float(v))
    if n_soil_classes > 1:
        cls = ((In // 64) + (Jn // 64)) % n_soil_classes
        ks_a = ks * (1.0 + 0.25 * cls)
        lam_a = lam * (1.0 + 0.05 * cls)
    else:
        ks_a, lam_a = full(ks), full(lam)
    psib_abs = abs(psib)
    nwt0 = full(gw_depth_mm)
    # initial moisture = profile integral of the initial water table (tHydroModel.cpp:357-372)
    dth = ths - thr
    mi0 = thr * (nwt0 + psib) - ths * psib - dth / (lam_a - 1.0) * (psib + nwt0 * np.power(psib_abs / nwt0, lam_a))
    basin = {
        "n_active": np.array([n], np.int32), "n_nodes_total": np.array([NX * NY], np.int32),
        "n_active_edges": np.array([2 * int(pair_active.sum())], np.int32),
        "id": np.arange(n, dtype=np.int32),
        "boundary": np.where(is_stream[order], 3, 0).astype(np.int32),
        "x": take(x), "y": take(y), "z": take(z), "varea": varea, "bedrock": full(bedrock_mm),
        "flow_slope": flow_slope, "flow_len": slen[np.arange(n), kbo], "flow_vedglen": flow_vedglen,
        "flow_dest": fdest, "stream_node": remap(stream_node)[order],
        "hillpath": hp, "streampath": sp, "ttime": ttime, "contr_area": contr,
        "reach": np.zeros(n, np.int32), "aspect": aspect,
        "Ks": ks_a, "ThetaS": full(ths), "ThetaR": full(thr), "PoreSize": lam_a, "AirEBubPres": full(psib),
        "DecayF": full(f), "SatAnRatio": full(ar), "UnsatAnRatio": full(uar),
        "spoke_rowptr": rowptr, "spoke_nbr": nbr.ravel(), "spoke_edge_listpos": listpos,
        "spoke_len": slen.ravel(), "spoke_vedglen": svl.ravel(),
        "spoke_compl_len": slen_in.ravel(), "spoke_compl_vedglen": svl_in.ravel(),
        "hillvel": np.array([hillvel]), "streamvel": np.array([streamvel]), "velratio": np.array([velratio]),
        "velcoef": np.array([velcoef]), "flowexp": np.array([0.0]), "baseflow": np.array([0.01]),
        "kincoef": np.array([0.1]), "dtReff": np.array([0.5]), "BasArea": np.array([bas_area]),
        "BasArea_flow": np.array([bas_area]), "maxttime": np.array([maxtt]),
        "outlet_stream_node": np.array([-1], np.int32), "outlet_contr_area": np.array([bas_area]),
        "res_limit": np.array([res_limit], np.int32),
        "IntStormMAX": np.array([10000.0]), "surfaceSoilDepth": np.array([100.0]), "rootZoneDepth": np.array([1000.0]),
        "EToption": np.array([0], np.int32), "Ioption": np.array([0], np.int32), "SnOpt": np.array([0], np.int32),
        "RunOnoption": np.array([0], np.int32), "RdstrOption": np.array([0], np.int32),
        "GW_model_label": np.array([1], np.int32),
        "tstep": np.array([tstep_h]), "gwstep": np.array([gwstep_h]), "metstep": np.array([1.0]),
        "etistep": np.array([1.0]), "endtime": np.array([runtime_h]), "output_interval": np.array([out_interval]),
        "init_NwtOld": nwt0, "init_NwtNew": nwt0.copy(), "init_MuOld": mi0, "init_MuNew": mi0.copy(),
        "init_MiOld": mi0.copy(), "init_MiNew": mi0.copy(),
    }
    # the channel network as tKinemat sets it up for this deck: one reach from the head of the river to the basin outlet,
    # widths from the geomorphic relation CHANNELWIDTHCOEFF * (contributing area in km2) ** CHANNELWIDTHEXPNT
    # (tKinemat::AssignChannelWidths, tKinemat.cpp:404-446; synth.py writes 1.0 and 0.3), CHANNELROUGHNESS 0.15
    stream_mask = basin["boundary"] == 3
    river = np.flatnonzero(stream_mask)
    river_head = river[~np.isin(river, fdest[river])]
    chan_width = np.zeros(n)
    chan_width[river] = [1.0 * math.pow(a * 1.0E-6, 0.3) for a in contr[river].tolist()]     # libm's pow, as the reference
    basin.update({
        "reach_head": river_head[:1].astype(np.int32), "reach_outlet": np.array([-1], np.int32),
        "reach_nnodes": np.array([river.size + 1], np.int32),
        "chan_width": chan_width, "outlet_chan_width": np.array([1.0 * math.pow(bas_area * 1.0E-6, 0.3)]),
        "chan_roughness": np.array([0.15]), "chan_width_uniform": np.array([-999.0]),
        "chan_percolation": np.array([0], np.int32), "chan_optres": np.array([0], np.int32),
    })
    if mesh_dir is not None:
        from .meshb import EDGE_DTYPE, NODE_DTYPE, write_meshb
        G = NX * NY
        nodes = np.zeros(G, NODE_DTYPE)                    # rows in id order = file order
        g = g_of_id
        nodes["id"] = np.arange(G)
        nodes["boundary"] = 1                               # closed ring ...
        nodes["boundary"][:n] = basin["boundary"]           # ... active cells (0 hillslope, 3 stream) ...
        nodes["boundary"][n] = 2                            # ... and the outlet
        nodes["x"], nodes["y"], nodes["z"] = x.ravel()[g], y.ravel()[g], z.ravel()[g]
        nodes["varea"][:n] = varea
        first_k = np.argmax(exists[g], axis=1)
        nodes["first_edge"] = eid[g, first_k]
        nodes["tracer"] = (np.arange(G) <= n).astype(np.int32)
        nodes["hillpath"][:n], nodes["traveltime"][:n], nodes["streampath"][:n] = hp, ttime, sp
        nodes["contr_area"][:n] = contr
        nodes["contr_area"][n] = bas_area
        nodes["reach"] = np.where(np.arange(G) <= n, 0, -1)
        nodes["flow_edge"] = -1
        nodes["flow_edge"][:n] = eid[cell_g, kbo]
        nodes["stream_node"] = -1
        sn = basin["stream_node"]
        nodes["stream_node"][:n] = np.where(basin["boundary"] == 3, -1, sn)
        ne = 2 * pr.size
        edges = np.zeros(ne, EDGE_DTYPE)
        edges["id"] = np.arange(ne)
        ccx_a, ccy_a = np.stack([c.ravel() for c in ccx], 1), np.stack([c.ravel() for c in ccy], 1)
        elen_a, vedg_a = np.stack([e.ravel() for e in elen], 1), np.stack([v.ravel() for v in vedg], 1)
        zf, xf, yf = z.ravel(), x.ravel(), y.ravel()
        for side in (0, 1):                                # the owner's edge, then its complement
            og_s = og if side == 0 else nb_flat[og, ok_]
            k_s = ok_ if side == 0 else (ok_ + 3) % 6
            dg_s = nb_flat[og_s, k_s]
            e = edges[side::2]
            inner = ~closed[og_s] | is_open[og_s]
            mx, my = 0.5 * (xf[og_s] + xf[dg_s]), 0.5 * (yf[og_s] + yf[dg_s])
            with np.errstate(all="ignore"):
                rvx, rvy = ccx_a[og_s, (k_s - 1) % 6], ccy_a[og_s, (k_s - 1) % 6]
                ve = vedg_a[og_s, k_s]
            okv = inner & np.isfinite(rvx) & np.isfinite(rvy) & ~is_open[og_s]
            e["rvx"], e["rvy"] = np.where(okv, rvx, mx), np.where(okv, rvy, my)
            ln = np.hypot(xf[dg_s] - xf[og_s], yf[dg_s] - yf[og_s])
            ln = np.where(~closed[og_s] & ~is_open[og_s], elen_a[og_s, k_s], ln)
            e["length"] = ln
            e["slope"] = (zf[og_s] - zf[dg_s]) / ln
            e["vedglen"] = np.where(np.isfinite(ve), ve, 0.0)
            e["orig"], e["dest"] = gid[og_s], gid[dg_s]
            nk = (k_s + 1) % 6                             # next existing spoke counter-clockwise
            for _ in range(5):
                miss = ~exists[og_s, nk]
                if not miss.any():
                    break
                nk = np.where(miss, (nk + 1) % 6, nk)
            e["ccw"] = eid[og_s, nk]
            edges[side::2] = e
        heads = np.flatnonzero((basin["boundary"] == 3) & ~np.isin(np.arange(n), fdest[basin["boundary"] == 3]))
        flow = dict(hillvel=hillvel, streamvel=streamvel, velratio=velratio, velcoef=velcoef, flowexp=0.0, baseflow=0.01,
                    flowout=0.01, maxttime=maxtt, dist_hill_max=float(hp.max()), dist_stream_max=float(sp.max()),
                    BasArea=bas_area, outlet=n, heads=heads, reach_heads=heads[:1], reach_outlets=[n],
                    reach_sizes=[int((basin["boundary"] == 3).sum()) + 1])
        write_meshb(mesh_dir, nodes, edges, flow)
    return basin


def add_met_forcing(basin: dict, hours: int = 48, et_option: int = 1, i_option: int = 2, winter: bool = False,
                    snow: bool = False) -> dict:
    """Adds what kernel (1) needs to a synthetic basin: one land class, per-cell thermal soil
    properties and aspect, one weather station with the hourly series of synth.met_series.
    Keys follow the reference-harness dump so GpuBasin.et_init / met.EvapoTransHost read them."""
    from .synth import BasinSpec, met_series
    n = int(basin["n_active"][0])
    spec = BasinSpec(runtime_h=float(hours), forcing="met", pmean=12.0, stdur=5.0, winter=winter)
    _, met = met_series(spec)
    nt = len(met)
    cols = np.array(met)                      # PA RH XC US TA IS
    xc = 0.5 * (basin["x"].min() + basin["x"].max())
    land = np.zeros((2, 15))
    land[1, 1:] = spec.land
    nd = np.full(nt, 9999.99)
    basin.update({
        "VolHeatCond": np.full(n, spec.soil[9]), "SoilHeatCap": np.full(n, spec.soil[10]),
        "soil_cutoff": basin["ThetaR"] + 1e-3, "root_cutoff": basin["ThetaR"] + 1e-3,
        "land_class": np.ones(n, np.int32), "land_table": land.ravel(),
        "et_option": np.array([et_option], np.int32), "et_Ioption": np.array([i_option], np.int32),
        "et_metdataOption": np.array([1], np.int32), "et_gFluxOption": np.array([2], np.int32),
        "et_luOption": np.array([0], np.int32), "et_snowOption": np.array([int(snow)], np.int32),
        "sn_minSnTemp": np.array([-20.0]), "sn_snliqfrac": np.array([0.06]), "sn_hillAlbedoOption": np.array([0], np.int32),
        "SnOpt": np.array([int(snow)], np.int32),
        "et_shelterOption": np.array([0], np.int32), "et_timeStep": np.array([60.0]),
        "et_Tlinke": np.array([2.5]), "et_tempLapseRate": np.array([-0.0065]),
        "icp_maxInterStormPeriod": np.array([5000.0]), "icp_option": np.array([i_option], np.int32),
        "rain_type": np.array([3], np.int32), "rain_optStorm": np.array([0], np.int32),
        "met_n_stations": np.array([1], np.int32), "met_station_id": np.array([1], np.int32),
        "met_gmt": np.array([-7], np.int32), "met_n_times": np.array([nt], np.int32),
        "met_n_params": np.array([11], np.int32), "met_lat": np.array([34.0]), "met_long": np.array([-107.0]),
        "met_other": np.array([float(basin["z"].mean())]),
        "met_airTemp": cols[:, 4].copy(), "met_dewTemp": nd.copy(), "met_surfTemp": nd.copy(),
        "met_rHumidity": cols[:, 1].copy(), "met_vPress": nd.copy(), "met_atmPress": cols[:, 0].copy(),
        "met_windSpeed": cols[:, 3].copy(), "met_skyCover": cols[:, 2].copy(), "met_radGlobal": cols[:, 5].copy(),
        "met_calendar": np.array([[2000, 6, 1 + h // 24, h % 24] for h in range(nt)], np.int32).ravel(),
        "met_assigned_station": np.ones(n, np.int32), "met_tsOption": np.array([2], np.int32),
        "met_vapOption": np.array([1], np.int32), "timer_start": np.array([2000, 6, 1, 0, 0], np.int32),
        "EToption": np.array([et_option], np.int32), "Ioption": np.array([i_option], np.int32),
    })
    return basin
