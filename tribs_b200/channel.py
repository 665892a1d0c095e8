"""The channel network as tKinemat holds it (src/tFlowNet/tKinemat.cpp:826-830: NodesLstH / NodesLstO / NNodes),
flattened for tribs_gpu_channel_init."""
from __future__ import annotations

import numpy as np


def channel_network(basin: dict) -> dict:
    """Per reach the chain of cell indices from its head along the flow edges to its outlet (index n_cells = the basin
    outlet node, which is not an active cell), with each entry's channel width and the length / slope of its flow edge.
    Keys as the reference harness flattens them (reach_head, reach_nnodes, flow_dest, chan_width, outlet_chan_width,
    flow_len, flow_slope, chan_roughness, chan_width_uniform)."""
    n = int(basin["n_active"][0])
    dest = np.asarray(basin["flow_dest"])
    off, node = [0], []
    for h, k in zip(np.asarray(basin["reach_head"]), np.asarray(basin["reach_nnodes"])):
        c = int(h)
        for _ in range(int(k)):
            node.append(c if 0 <= c < n else n)
            c = int(dest[c]) if 0 <= c < n else n
        off.append(len(node))
    node = np.array(node, np.int32)
    inside = node < n
    safe = np.where(inside, node, 0)
    width = np.where(inside, np.asarray(basin["chan_width"], float)[safe], float(basin["outlet_chan_width"][0]))
    return {"n_cells": n, "off": np.array(off, np.int32), "node": node, "width": np.ascontiguousarray(width),
            "length": np.ascontiguousarray(np.where(inside, np.asarray(basin["flow_len"], float)[safe], 0.0)),
            "slope": np.ascontiguousarray(np.where(inside, np.asarray(basin["flow_slope"], float)[safe], 0.0)),
            "roughness": float(basin["chan_roughness"][0]), "uniform_width": float(basin["chan_width_uniform"][0])}
