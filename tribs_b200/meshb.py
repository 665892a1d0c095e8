"""OPTMESHINPUT 9 ("MeshBuilder") files of the reference: reader and writer.

Formats as defined by the reference's own readers (native endianness, no padding):
  nodes.meshb  tMesh::MakeMeshFromMeshBuilder, src/tMesh/tMesh.cpp:1089-1158 (192 B per node)
  edges.meshb  same function, 1173-1222 (56 B per directed edge; ccw pass 1284-1306)
  flow.meshb   tFlowNet::ReadFlowNetFromMeshBuilder, src/tFlowNet/tFlowNet.cpp:117-175
With these files the reference skips its mesh construction and flow-network set-up - the part that
is super-linear in the number of nodes - so it can run the large synthetic
basins of synthbasin.py itself: full-size parity fixtures and a CPU baseline on the same workload.
Checked against the reference in tests/test_meshb.py (round trip through the reference harness's
--dump-meshb, and a run of the reference on the files written here)."""
from __future__ import annotations

import os

import numpy as np

NODE_DTYPE = np.dtype([
    ("id", "<i4"), ("boundary", "<i4"), ("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("varea", "<f8"),
    ("first_edge", "<i4"), ("flood", "<i4"), ("tracer", "<i4"), ("hillpath", "<f8"), ("traveltime", "<f8"),
    ("srf", "<f8"), ("hsrf", "<f8"), ("psrf", "<f8"), ("satsrf", "<f8"), ("sbsrf", "<f8"), ("evapotrans", "<f8"),
    ("soil_moisture_sc", "<f8"), ("soil_moisture_unsc", "<f8"), ("root_moisture_sc", "<f8"), ("contr_area", "<f8"),
    ("nwt_new", "<f8"), ("rain", "<f8"), ("curvature", "<f8"), ("streampath", "<f8"), ("reach", "<i4"),
    ("flow_edge", "<i4"), ("stream_node", "<i4")])
EDGE_DTYPE = np.dtype([("id", "<i4"), ("rvx", "<f8"), ("rvy", "<f8"), ("length", "<f8"), ("slope", "<f8"),
                       ("vedglen", "<f8"), ("orig", "<i4"), ("dest", "<i4"), ("ccw", "<i4")])
FLOW_SCALARS = ["hillvel", "streamvel", "velratio", "velcoef", "flowexp", "baseflow", "flowout", "maxttime",
                "dist_hill_max", "dist_stream_max", "BasArea"]
assert NODE_DTYPE.itemsize == 192 and EDGE_DTYPE.itemsize == 56


def read_meshb(d: str) -> dict:
    out = {}
    for name, dt in (("nodes", NODE_DTYPE), ("edges", EDGE_DTYPE)):
        raw = open(os.path.join(d, name + ".meshb"), "rb").read()
        n = int(np.frombuffer(raw, "<i4", 1)[0])
        out[name] = np.frombuffer(raw, dt, n, offset=4).copy()
    raw = open(os.path.join(d, "flow.meshb"), "rb").read()
    sc = np.frombuffer(raw, "<f8", len(FLOW_SCALARS))
    flow = dict(zip(FLOW_SCALARS, (float(v) for v in sc)))
    ints = np.frombuffer(raw, "<i4", offset=8 * len(FLOW_SCALARS))
    flow["outlet"] = int(ints[0])
    p = 1
    for key in ("heads", "reach_heads", "reach_outlets", "reach_sizes"):
        k = int(ints[p])
        flow[key] = ints[p + 1:p + 1 + k].copy()
        p += 1 + k
    assert p == ints.size
    out["flow"] = flow
    return out


def write_meshb(d: str, nodes: np.ndarray, edges: np.ndarray, flow: dict) -> None:
    assert nodes.dtype == NODE_DTYPE and edges.dtype == EDGE_DTYPE
    for name, a in (("nodes", nodes), ("edges", edges)):
        with open(os.path.join(d, name + ".meshb"), "wb") as f:
            f.write(np.array([a.size], "<i4").tobytes())
            f.write(a.tobytes())
    with open(os.path.join(d, "flow.meshb"), "wb") as f:
        f.write(np.array([flow[k] for k in FLOW_SCALARS], "<f8").tobytes())
        f.write(np.array([flow["outlet"]], "<i4").tobytes())
        for key in ("heads", "reach_heads", "reach_outlets", "reach_sizes"):
            a = np.asarray(flow[key], "<i4")
            f.write(np.array([a.size], "<i4").tobytes())
            f.write(a.tobytes())
