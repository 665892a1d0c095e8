"""ctypes binding of libtribs_gpu.so — the C ABI declared in include/tribs_gpu.h.

This module is plumbing only: it marshals numpy arrays into the C structs and calls the
library. There is no CPU fallback: importing works anywhere, but `load()` raises if the CUDA
library is missing and `tribs_gpu_init` fails loudly without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB_PATH = os.environ.get("TRIBS_GPU_LIB", os.path.join(PKG, "libtribs_gpu.so"))
HEADER = os.path.join(ROOT, "include", "tribs_gpu.h")


def _fields_from_header() -> list[str]:
    src = open(HEADER).read()
    m = re.search(r"#define TRIBS_FIELDS\(X\) \\\n((?:.*\\\n)*.*)\n", src)
    return re.findall(r"X\((\w+)\)", m.group(1))


def _enum_from_header(prefix: str) -> list[str]:
    src = open(HEADER).read()
    names = re.findall(r"\b" + prefix + r"(\w+)\b", src.split("typedef struct tribs_ctx")[1])
    out = []
    for nm in names:
        if nm not in out:
            out.append(nm)
    return out


def _et_fields_from_header() -> list[str]:
    src = open(HEADER).read()
    m = re.search(r"#define TRIBS_ET_FIELDS\(X\) \\\n((?:.*\\\n)*.*)\n", src)
    return re.findall(r"X\((\w+)\)", m.group(1))


FIELDS = _fields_from_header()
F = {nm: k for k, nm in enumerate(FIELDS)}
def _snow_fields_from_header() -> list[str]:
    src = open(HEADER).read()
    m = re.search(r"#define TRIBS_SNOW_FIELDS\(X\) \\\n((?:.*\\\n)*.*)\n", src)
    return re.findall(r"X\((\w+)\)", m.group(1))


ET_FIELDS = _et_fields_from_header()
EF = {nm: k for k, nm in enumerate(ET_FIELDS)}
SNOW_FIELDS = _snow_fields_from_header()
SNF = {nm: k for k, nm in enumerate(SNOW_FIELDS)}
RESULTS = ["mhydro", "HsrfRout", "SbsrfRout", "PsrfRout", "SatsrfRout", "crr", "rmax", "rmin", "frac",
           "msm", "msmRt", "msmU", "mgw", "sat", "qunsat"]
GLOBALS = ["Stok", "TotRain", "TotGWchange", "TotMoist", "psrf_last", "hillvel_last", "BasinRain", "BasinEvap",
           "BasinRunoff"]
PHASE_UNSAT, PHASE_SAT, PHASE_ROUTE, PHASE_RESET = 1, 2, 4, 8
RAIN_KEEP, RAIN_UNIFORM, RAIN_PER_CELL, RAIN_GAUGES = 0, 1, 2, 3
IPC_HANDLE_BYTES = 64


class GaugeRain:
    """Rain of one step given per rain gauge (tribs_gpu_set_rain_gauges holds the cell -> gauge map)."""

    def __init__(self, values):
        self.values = np.ascontiguousarray(values, np.float64)

PD = C.POINTER(C.c_double)
PI = C.POINTER(C.c_int32)


class Mesh(C.Structure):
    _fields_ = [("n_cells", C.c_int32), ("n_edges", C.c_int32), ("boundary", PI), ("z", PD), ("varea", PD),
                ("flow_slope", PD), ("flow_vedglen", PD), ("flow_dest", PI), ("stream_node", PI),
                ("hillpath", PD), ("streampath", PD), ("contr_area", PD), ("ttime", PD), ("reach", PI),
                ("nbr_rowptr", PI), ("nbr_col", PI), ("nbr_len", PD), ("nbr_vedglen", PD),
                ("nbr_len_in", PD), ("nbr_vedglen_in", PD), ("nbr_listpos", PI)]


class Params(C.Structure):
    _fields_ = [("Ks", PD), ("ThetaS", PD), ("ThetaR", PD), ("PoreSize", PD), ("AirEBubPres", PD),
                ("DecayF", PD), ("SatAnRatio", PD), ("UnsatAnRatio", PD), ("bedrock", PD),
                ("BasArea", C.c_double), ("BasArea_flow", C.c_double), ("IntStormMAX", C.c_double),
                ("surfaceSoilDepth", C.c_double), ("rootZoneDepth", C.c_double),
                ("EToption", C.c_int32), ("Ioption", C.c_int32), ("SnOpt", C.c_int32), ("RdstrOption", C.c_int32),
                ("velcoef", C.c_double), ("velratio", C.c_double), ("flowexp", C.c_double), ("baseflow", C.c_double),
                ("kincoef", C.c_double), ("dtReff", C.c_double), ("outlet_contr_area", C.c_double),
                ("tstep", C.c_double), ("gwstep", C.c_double), ("etistep", C.c_double),
                ("output_interval", C.c_double), ("res_limit", C.c_int32), ("reserved", C.c_int32)]


class StateT(C.Structure):
    _fields_ = [("f", PD * len(FIELDS))]


class Part(C.Structure):
    _fields_ = [("rank", C.c_int32), ("nranks", C.c_int32), ("n_parts", C.c_int32), ("max_batch_steps", C.c_int32),
                ("cell_owner", PI)]


class Channel(C.Structure):
    _fields_ = [("n_reach", C.c_int32), ("percolation_option", C.c_int32), ("reservoir_option", C.c_int32), ("pad_", C.c_int32),
                ("reach_off", PI), ("reach_node", PI), ("width", PD), ("length", PD), ("slope", PD),
                ("roughness", C.c_double), ("uniform_width", C.c_double), ("dt_s", C.c_double),
                ("hlevel0", PD), ("qstrm0", PD), ("outlet_hlev0", PD)]


class Step(C.Structure):
    _fields_ = [("current_time", C.c_double), ("dt", C.c_double), ("dt_gw", C.c_double),
                ("elapsed_steps", C.c_int32), ("hourly_reset", C.c_int32), ("rain_mode", C.c_int32),
                ("reserved", C.c_int32), ("rain_uniform", C.c_double), ("rain", PD), ("qstrm", PD),
                ("qstrm_outlet", C.c_double)]


class EtParams(C.Structure):
    _fields_ = [("flow_slope", PD), ("aspect", PD), ("vol_heat_cond", PD), ("soil_heat_cap", PD),
                ("soil_cutoff", PD), ("root_cutoff", PD), ("land_class", PI), ("n_land_classes", C.c_int32),
                ("land_table", PD), ("et_option", C.c_int32), ("i_option", C.c_int32), ("gflux_option", C.c_int32),
                ("shelter_option", C.c_int32), ("lu_option", C.c_int32), ("snow_option", C.c_int32),
                ("metstep_min", C.c_double), ("etistep_h", C.c_double), ("tlinke", C.c_double),
                ("temp_lapse", C.c_double), ("max_interstorm", C.c_double), ("init", C.POINTER(PD)),
                ("hor_angle_0000", PD)]


class MetT(C.Structure):
    _fields_ = [("n_records", C.c_int32), ("cell_record", PI), ("air_temp", PD), ("dew_temp", PD), ("rel_humid", PD),
                ("vap_press", PD), ("atm_press", PD), ("wind_speed", PD), ("sky_cover", PD), ("rad_global", PD),
                ("elev", PD)]


class EtStep(C.Structure):
    _fields_ = [("do_potential", C.c_int32), ("do_components", C.c_int32), ("intercept_flag", C.c_int32),
                ("alphaD", C.c_double), ("sinAlpha", C.c_double), ("sunaz", C.c_double), ("Io", C.c_double),
                ("hour", C.c_int32), ("first_call", C.c_int32), ("dew_hum_flag", C.c_int32), ("ts_option", C.c_int32),
                ("sky_flag_from", C.c_int32), ("te_met", C.c_double), ("te_eti", C.c_double),
                ("met_potential", C.POINTER(MetT)), ("met_components", C.POINTER(MetT))]


class SnowParams(C.Structure):
    _fields_ = [("min_sn_temp", C.c_double), ("snliqfrac", C.c_double), ("hill_albedo_option", C.c_int32),
                ("reserved", C.c_int32), ("init", C.POINTER(PD))]


EXPORTS = ["tribs_gpu_init", "tribs_gpu_step", "tribs_gpu_run", "tribs_gpu_run_et", "tribs_gpu_run_snow", "tribs_gpu_retrieve_qeff_steps", "tribs_gpu_sync", "tribs_gpu_push", "tribs_gpu_retrieve_qeff",
           "tribs_gpu_pending_qeff",
           "tribs_gpu_n_stream", "tribs_gpu_stream_cells", "tribs_gpu_get_results", "tribs_gpu_get_globals",
           "tribs_gpu_sweep_levels", "tribs_gpu_device_index", "tribs_gpu_launch_count", "tribs_gpu_set_timing",
           "tribs_gpu_phase_ms", "tribs_gpu_wait", "tribs_gpu_stream", "tribs_gpu_destroy",
           "tribs_gpu_set_rain_gauges", "tribs_gpu_ipc_export", "tribs_gpu_attach_peers", "tribs_gpu_attach_local",
           "tribs_gpu_owned_range", "tribs_gpu_snapshot", "tribs_gpu_rollback", "tribs_gpu_save_state",
           "tribs_gpu_load_state", "tribs_gpu_reserve_batch", "tribs_gpu_set_partition", "tribs_gpu_halo_pack", "tribs_gpu_halo_unpack", "tribs_gpu_set_global",
           "tribs_gpu_padded_cells", "tribs_gpu_et_init", "tribs_gpu_et_step", "tribs_gpu_et_sync", "tribs_gpu_et_push",
           "tribs_gpu_snow_init", "tribs_gpu_snow_step", "tribs_gpu_snow_sync", "tribs_gpu_snow_members",
           "tribs_gpu_snow_push", "tribs_gpu_snow_set_members",
           "tribs_gpu_channel_init", "tribs_gpu_channel_route", "tribs_gpu_channel_sync",
           "tribs_gpu_last_error", "tribs_gpu_abi_version"]

_lib = None


def load():
    """Load libtribs_gpu.so (built by __graft_entry__.build()). Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                           "tribs_b200 has no CPU path.")
    L = C.CDLL(LIB_PATH)
    want = int(re.search(r"#define\s+TRIBS_GPU_ABI_VERSION\s+(\d+)", open(HEADER).read()).group(1))
    if L.tribs_gpu_abi_version() != want:      # a library built from an older header: the struct layouts differ
        raise RuntimeError(f"{LIB_PATH} has ABI version {L.tribs_gpu_abi_version()}, include/tribs_gpu.h has {want}: rebuild "
                           "(python -c 'import __graft_entry__ as g; g.build()')")
    H = C.c_void_p
    L.tribs_gpu_init.argtypes = [C.POINTER(Mesh), C.POINTER(Params), C.POINTER(StateT), C.POINTER(Part), C.POINTER(H)]
    L.tribs_gpu_step.argtypes = [H, C.c_uint32, C.POINTER(Step)]
    L.tribs_gpu_run.argtypes = [H, C.c_int, C.POINTER(Step), C.POINTER(C.c_uint32)]
    L.tribs_gpu_run_et.argtypes = [H, C.c_int, C.POINTER(Step), C.POINTER(C.c_uint32), C.POINTER(C.POINTER(EtStep))]
    L.tribs_gpu_run_snow.argtypes = [H, C.c_int, C.POINTER(Step), C.POINTER(C.c_uint32), C.POINTER(C.POINTER(EtStep)), PI]
    L.tribs_gpu_retrieve_qeff_steps.argtypes = [H, C.c_int, PD]
    L.tribs_gpu_sync.argtypes = [H, C.c_uint64, C.POINTER(StateT)]
    L.tribs_gpu_push.argtypes = [H, C.c_uint64, C.POINTER(StateT)]
    L.tribs_gpu_retrieve_qeff.argtypes = [H, C.c_double, PD]
    L.tribs_gpu_pending_qeff.argtypes = [H, C.c_double, C.c_int, PD]
    L.tribs_gpu_n_stream.argtypes = [H]
    L.tribs_gpu_stream_cells.argtypes = [H, PI]
    L.tribs_gpu_get_results.argtypes = [H, C.c_int, PD, C.c_int]
    L.tribs_gpu_get_globals.argtypes = [H, PD]
    L.tribs_gpu_sweep_levels.argtypes = [H]
    L.tribs_gpu_device_index.argtypes = [H, PI]
    L.tribs_gpu_launch_count.argtypes = [H]
    L.tribs_gpu_launch_count.restype = C.c_int64
    L.tribs_gpu_set_timing.argtypes = [H, C.c_int]
    L.tribs_gpu_phase_ms.argtypes = [H, PD, C.c_int]
    L.tribs_gpu_wait.argtypes = [H]
    L.tribs_gpu_stream.argtypes = [H]
    L.tribs_gpu_stream.restype = C.c_void_p
    L.tribs_gpu_destroy.argtypes = [H]
    L.tribs_gpu_et_init.argtypes = [H, C.POINTER(EtParams)]
    L.tribs_gpu_et_step.argtypes = [H, C.POINTER(EtStep)]
    L.tribs_gpu_et_sync.argtypes = [H, C.c_uint64, C.POINTER(PD)]
    L.tribs_gpu_et_push.argtypes = [H, C.c_uint64, C.POINTER(PD)]
    L.tribs_gpu_snow_init.argtypes = [H, C.POINTER(SnowParams)]
    L.tribs_gpu_snow_step.argtypes = [H, C.POINTER(EtStep), C.c_int32]
    L.tribs_gpu_snow_sync.argtypes = [H, C.c_uint64, C.POINTER(PD)]
    L.tribs_gpu_snow_members.argtypes = [H, PD, C.POINTER(C.c_int32)]
    L.tribs_gpu_snow_push.argtypes = [H, C.c_uint64, C.POINTER(PD)]
    L.tribs_gpu_snow_set_members.argtypes = [H, PD]
    L.tribs_gpu_set_rain_gauges.argtypes = [H, C.c_int32, PI]
    L.tribs_gpu_ipc_export.argtypes = [H, C.c_void_p]
    L.tribs_gpu_attach_peers.argtypes = [H, C.c_void_p]
    L.tribs_gpu_attach_local.argtypes = [C.POINTER(H), C.c_int]
    L.tribs_gpu_owned_range.argtypes = [H, PI, PI]
    L.tribs_gpu_snapshot.argtypes = [H]
    L.tribs_gpu_rollback.argtypes = [H]
    L.tribs_gpu_save_state.argtypes = [H, C.c_char_p]
    L.tribs_gpu_load_state.argtypes = [H, C.c_char_p]
    L.tribs_gpu_reserve_batch.argtypes = [H, C.c_int]
    L.tribs_gpu_channel_init.argtypes = [H, C.POINTER(Channel)]
    L.tribs_gpu_channel_route.argtypes = [H, C.c_int32, PD]
    L.tribs_gpu_channel_sync.argtypes = [H, PD, PD, PD, PD, C.POINTER(C.c_int64)]
    L.tribs_gpu_last_error.restype = C.c_char_p
    _lib = L
    return L


class TribsGpuError(RuntimeError):
    pass


def _chk(rc: int):
    if rc != 0:
        raise TribsGpuError(load().tribs_gpu_last_error().decode())


def _d(a):
    return a.ctypes.data_as(PD)


def _i(a):
    return a.ctypes.data_as(PI)


# names in a flattened-basin dict (basin.trb / tribs_b200.synthbasin) -> tribs_mesh_t members
_MESH_F64 = {"z": "z", "varea": "varea", "flow_slope": "flow_slope", "flow_vedglen": "flow_vedglen",
             "hillpath": "hillpath", "streampath": "streampath", "contr_area": "contr_area", "ttime": "ttime",
             "nbr_len": "spoke_len", "nbr_vedglen": "spoke_vedglen", "nbr_len_in": "spoke_compl_len",
             "nbr_vedglen_in": "spoke_compl_vedglen"}
_MESH_I32 = {"boundary": "boundary", "flow_dest": "flow_dest", "stream_node": "stream_node", "reach": "reach",
             "nbr_rowptr": "spoke_rowptr", "nbr_col": "spoke_nbr", "nbr_listpos": "spoke_edge_listpos"}
_PRM_F64 = ["Ks", "ThetaS", "ThetaR", "PoreSize", "AirEBubPres", "DecayF", "SatAnRatio", "UnsatAnRatio", "bedrock"]
_PRM_SC = ["BasArea", "BasArea_flow", "IntStormMAX", "surfaceSoilDepth", "rootZoneDepth", "velcoef", "velratio",
           "flowexp", "baseflow", "kincoef", "dtReff", "outlet_contr_area", "tstep", "gwstep", "etistep",
           "output_interval"]
_PRM_INT = ["EToption", "Ioption", "SnOpt", "RdstrOption"]


class GpuBasin:
    """Owns one tribs_handle_t. `basin` is a dict of flattened arrays in sweep order."""

    def __init__(self, basin: dict, part: dict | None = None):
        """part (optional): rank, nranks, cell_owner [n] and, for nranks > 1, max_batch_steps: the reach
        partition of tribs_part_t. nranks == 1 with several partitions in cell_owner lays the basin out
        and sums it like the multi-GPU run, on one GPU."""
        self.L = load()
        b = basin
        self.n = int(b["n_active"][0])
        self._keep = []
        mesh = Mesh()
        mesh.n_cells = self.n
        mesh.n_edges = int(b["spoke_rowptr"][-1])
        for member, key in _MESH_F64.items():
            a = np.ascontiguousarray(b[key], np.float64); self._keep.append(a); setattr(mesh, member, _d(a))
        for member, key in _MESH_I32.items():
            a = np.ascontiguousarray(b[key], np.int32); self._keep.append(a); setattr(mesh, member, _i(a))
        prm = Params()
        for key in _PRM_F64:
            a = np.ascontiguousarray(b[key], np.float64); self._keep.append(a); setattr(prm, key, _d(a))
        for key in _PRM_SC:
            setattr(prm, key, float(b[key][0]))
        for key in _PRM_INT:
            setattr(prm, key, int(b[key][0]))
        prm.res_limit = int(b["res_limit"][0])
        st = StateT()
        for k, nm in enumerate(FIELDS):
            src = b.get("init_" + nm)
            if src is not None:
                a = np.ascontiguousarray(src, np.float64); self._keep.append(a); st.f[k] = _d(a)
        self.h = C.c_void_p()
        pt = None
        self.rank, self.nranks = 0, 1
        if part is not None:
            own = np.ascontiguousarray(part["cell_owner"], np.int32); self._keep.append(own)
            if own.size != self.n:
                raise ValueError("cell_owner must have one entry per active cell")
            pt = Part()
            pt.rank, pt.nranks = int(part.get("rank", 0)), int(part.get("nranks", 1))
            pt.n_parts = int(part.get("n_parts", own.max() + 1 if pt.nranks == 1 else pt.nranks))
            pt.max_batch_steps = int(part.get("max_batch_steps", 64))
            pt.cell_owner = _i(own)
            self.rank, self.nranks = pt.rank, pt.nranks
        _chk(self.L.tribs_gpu_init(C.byref(mesh), C.byref(prm), C.byref(st), C.byref(pt) if pt is not None else None,
                                   C.byref(self.h)))
        self.tstep = prm.tstep
        self.gwstep = prm.gwstep
        self.etistep = prm.etistep
        self.n_stream = self.L.tribs_gpu_n_stream(self.h)
        self._keep.clear()

    def step(self, phases: int, current_time: float, elapsed_steps: int, hourly_reset: int,
             rain=None, qstrm=None, qstrm_outlet: float = 0.0):
        s = Step()
        s.current_time = current_time
        s.dt = self.tstep
        s.dt_gw = self.gwstep
        s.elapsed_steps = elapsed_steps
        s.hourly_reset = hourly_reset
        if rain is None:
            s.rain_mode = RAIN_KEEP
        elif isinstance(rain, GaugeRain):
            s.rain_mode = RAIN_GAUGES
            s.rain = _d(rain.values)
        elif np.isscalar(rain):
            s.rain_mode = RAIN_UNIFORM
            s.rain_uniform = float(rain)
        else:
            rain = np.ascontiguousarray(rain, np.float64)      # held until the call returns
            if rain.size != self.n:
                raise ValueError(f"rain has {rain.size} entries, the basin has {self.n} cells")
            s.rain_mode = RAIN_PER_CELL
            s.rain = _d(rain)
        if qstrm is not None:
            qstrm = np.ascontiguousarray(qstrm, np.float64)
            if qstrm.size != self.n:
                raise ValueError(f"qstrm has {qstrm.size} entries, the basin has {self.n} cells")
            s.qstrm = _d(qstrm)
        s.qstrm_outlet = qstrm_outlet
        _chk(self.L.tribs_gpu_step(self.h, phases, C.byref(s)))

    def run(self, steps: list[dict]):
        """Pipelined batch. Each dict: current_time, elapsed_steps, hourly_reset, sat (bool), rain
        (None | scalar | array), qstrm_outlet (optional)."""
        k = len(steps)
        arr = (Step * k)()
        masks = (C.c_uint32 * k)()
        keep = []
        for q, sd in enumerate(steps):
            s = arr[q]
            s.current_time = sd["current_time"]
            s.dt = self.tstep
            s.dt_gw = self.gwstep
            s.elapsed_steps = sd["elapsed_steps"]
            s.hourly_reset = sd["hourly_reset"]
            rain = sd.get("rain")
            if rain is None:
                s.rain_mode = RAIN_KEEP
            elif isinstance(rain, GaugeRain):
                keep.append(rain.values)
                s.rain_mode = RAIN_GAUGES
                s.rain = _d(rain.values)
            elif np.isscalar(rain):
                s.rain_mode = RAIN_UNIFORM
                s.rain_uniform = float(rain)
            else:
                a = np.ascontiguousarray(rain, np.float64); keep.append(a)
                if a.size != self.n:
                    raise ValueError(f"rain has {a.size} entries, the basin has {self.n} cells")
                s.rain_mode = RAIN_PER_CELL
                s.rain = _d(a)
            s.qstrm_outlet = sd.get("qstrm_outlet", 0.0)
            masks[q] = PHASE_UNSAT | PHASE_ROUTE | PHASE_RESET | (PHASE_SAT if sd.get("sat") else 0)
        if any(sd.get("snow") is not None for sd in steps):      # (step, rec, hourly, cell_record, elev) as for snow_step
            ets = (C.POINTER(EtStep) * k)()
            hourly = (C.c_int32 * k)()
            for q, sd in enumerate(steps):
                if sd.get("snow") is not None:
                    step, rec, hr, cell_record, elev = sd["snow"]
                    es = self._snow_step_struct(step, rec, cell_record, elev, keep)
                    keep.append(es)
                    ets[q] = C.pointer(es)
                    hourly[q] = int(hr)
            _chk(self.L.tribs_gpu_run_snow(self.h, k, arr, masks, ets, hourly))
        elif any(sd.get("et") is not None for sd in steps):
            ets = (C.POINTER(EtStep) * k)()
            for q, sd in enumerate(steps):
                if sd.get("et") is not None:
                    es = self._et_step_struct(*sd["et"], keep=keep)
                    keep.append(es)
                    ets[q] = C.pointer(es)
            _chk(self.L.tribs_gpu_run_et(self.h, k, arr, masks, ets))
        else:
            _chk(self.L.tribs_gpu_run(self.h, k, arr, masks))

    # -- rain gauges and peer-memory partitions --------------------------------------------------
    def set_rain_gauges(self, cell_gauge, n_gauges: int | None = None):
        g = np.ascontiguousarray(cell_gauge, np.int32)
        _chk(self.L.tribs_gpu_set_rain_gauges(self.h, int(g.max()) + 1 if n_gauges is None else int(n_gauges), _i(g)))

    def ipc_export(self) -> bytes:
        buf = C.create_string_buffer(IPC_HANDLE_BYTES)
        _chk(self.L.tribs_gpu_ipc_export(self.h, buf))
        return buf.raw

    def attach_peers(self, handles: list[bytes]):
        blob = b"".join(handles)
        if len(blob) != IPC_HANDLE_BYTES * self.nranks:
            raise ValueError("one 64-byte handle per rank, in rank order")
        _chk(self.L.tribs_gpu_attach_peers(self.h, C.create_string_buffer(blob, len(blob))))

    @staticmethod
    def attach_local(devs: list["GpuBasin"]):
        arr = (C.c_void_p * len(devs))(*[d.h for d in devs])
        _chk(load().tribs_gpu_attach_local(arr, len(devs)))

    def snapshot(self):
        _chk(self.L.tribs_gpu_snapshot(self.h))

    def rollback(self):
        _chk(self.L.tribs_gpu_rollback(self.h))

    def save_state(self, path: str):
        _chk(self.L.tribs_gpu_save_state(self.h, path.encode()))

    def load_state(self, path: str):
        _chk(self.L.tribs_gpu_load_state(self.h, path.encode()))

    def reserve_batch(self, n_steps: int):
        _chk(self.L.tribs_gpu_reserve_batch(self.h, int(n_steps)))

    # ---- channel routing (tKinemat::RunRoutingModel's kinematic-wave half, tKinemat.cpp:803-897)
    def channel_init(self, net: dict, dt_s: float, hlevel0=None, qstrm0=None, outlet_hlev0=None,
                     percolation_option: int = 0, reservoir_option: int = 0):
        """net: tKinemat's reach lists as flat arrays (tribs_b200.channel.channel_network)."""
        keep = {k: np.ascontiguousarray(net[k], np.int32) for k in ("off", "node")}
        keep.update({k: np.ascontiguousarray(net[k], float) for k in ("width", "length", "slope")})
        opt = {nm: (None if a is None else np.ascontiguousarray(a, float))
               for nm, a in (("hlevel0", hlevel0), ("qstrm0", qstrm0), ("outlet_hlev0", outlet_hlev0))}
        for nm in ("hlevel0", "qstrm0"):
            if opt[nm] is not None and opt[nm].size != self.n + 1:
                raise ValueError(f"{nm} must hold n_cells + 1 values (the last one the basin outlet)")
        ch = Channel(len(keep["off"]) - 1, int(percolation_option), int(reservoir_option), 0, _i(keep["off"]), _i(keep["node"]),
                     _d(keep["width"]), _d(keep["length"]), _d(keep["slope"]), float(net["roughness"]),
                     float(net["uniform_width"]), float(dt_s),
                     *[(_d(opt[nm]) if opt[nm] is not None else None) for nm in ("hlevel0", "qstrm0", "outlet_hlev0")])
        _chk(self.L.tribs_gpu_channel_init(self.h, C.byref(ch)))
        self.n_reach = len(keep["off"]) - 1

    def channel_route(self, n_steps: int) -> np.ndarray:
        """Outlet discharge (tKinemat::Qout) after each step of the last batch (n_steps >= 1) or of the last single
        step (n_steps == 0)."""
        out = np.empty(max(int(n_steps), 1))
        _chk(self.L.tribs_gpu_channel_route(self.h, int(n_steps), _d(out)))
        return out

    def channel_sync(self) -> dict:
        """Hlevel / Qstrm / FlowVelocity per stream slot (+ the basin outlet), OutletHlev per reach, solver counters."""
        slots = self.n_stream + 1
        out = {"Hlevel": np.empty(slots), "Qstrm": np.empty(slots), "FlowVelocity": np.empty(slots),
               "OutletHlev": np.empty(self.n_reach)}
        cnt = (C.c_int64 * 7)()
        _chk(self.L.tribs_gpu_channel_sync(self.h, _d(out["Hlevel"]), _d(out["Qstrm"]), _d(out["FlowVelocity"]),
                                           _d(out["OutletHlev"]), cnt))
        out["newton_iterations"], out["backtracks"], out["maxits_exceeded"] = int(cnt[0]), int(cnt[1]), int(cnt[2])
        out["cycles"] = {"band_solve": int(cnt[3]), "line_search_sums": int(cnt[4]), "function_sums": int(cnt[5]), "kernel": int(cnt[6])}
        return out

    def owned_range(self):
        a, b = C.c_int32(0), C.c_int32(0)
        _chk(self.L.tribs_gpu_owned_range(self.h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def device_index(self) -> np.ndarray:
        out = np.empty(self.n, np.int32)
        _chk(self.L.tribs_gpu_device_index(self.h, _i(out)))
        return out

    def retrieve_qeff_steps(self, k: int) -> np.ndarray:
        out = np.empty((k, self.n_stream + 1))
        _chk(self.L.tribs_gpu_retrieve_qeff_steps(self.h, k, _d(out)))
        return out

    def sync(self, names, out: dict | None = None) -> dict:
        out = {} if out is None else out
        st = StateT()
        mask = 0
        for nm in names:
            if nm not in out:
                out[nm] = np.empty(self.n)
            st.f[F[nm]] = _d(out[nm])
            mask |= 1 << F[nm]
        _chk(self.L.tribs_gpu_sync(self.h, mask, C.byref(st)))
        return out

    def push(self, arrays: dict):
        st = StateT()
        mask = 0
        keep = []
        for nm, a in arrays.items():
            a = np.ascontiguousarray(a, np.float64); keep.append(a)
            st.f[F[nm]] = _d(a)
            mask |= 1 << F[nm]
        _chk(self.L.tribs_gpu_push(self.h, mask, C.byref(st)))

    # -- kernel (1): ET + interception --------------------------------------------------------
    _MET_KEYS = (("air_temp", "airTemp"), ("dew_temp", "dewTemp"), ("rel_humid", "rHumidity"), ("vap_press", "vPress"),
                 ("atm_press", "atmPress"), ("wind_speed", "windSpeed"), ("sky_cover", "skyCover"),
                 ("rad_global", "radGlobal"))

    def et_init(self, basin: dict):
        """tribs_gpu_et_init from a flattened basin (keys as dumped by the reference harness)."""
        b, keep = basin, []
        p = EtParams()
        for member, key in (("flow_slope", "flow_slope"), ("aspect", "aspect"), ("vol_heat_cond", "VolHeatCond"),
                            ("soil_heat_cap", "SoilHeatCap"), ("soil_cutoff", "soil_cutoff"), ("root_cutoff", "root_cutoff")):
            a = np.ascontiguousarray(b[key], np.float64); keep.append(a); setattr(p, member, _d(a))
        lc = np.ascontiguousarray(b["land_class"], np.int32); keep.append(lc); p.land_class = _i(lc)
        lt = np.ascontiguousarray(b["land_table"], np.float64); keep.append(lt); p.land_table = _d(lt)
        p.n_land_classes = lt.size // 15
        p.et_option = int(b["et_option"][0]); p.i_option = int(b["et_Ioption"][0])
        p.gflux_option = int(b["et_gFluxOption"][0]); p.shelter_option = int(b["et_shelterOption"][0])
        p.lu_option = int(b["et_luOption"][0]); p.snow_option = int(b["et_snowOption"][0])
        p.metstep_min = float(b["et_timeStep"][0]); p.etistep_h = float(b["etistep"][0])
        p.tlinke = float(b["et_Tlinke"][0]); p.temp_lapse = float(b["et_tempLapseRate"][0])
        p.max_interstorm = float(b["icp_maxInterStormPeriod"][0])
        init = (PD * len(ET_FIELDS))()
        for k, nm in enumerate(ET_FIELDS):
            src = b.get("init_" + nm)
            if src is not None:
                a = np.ascontiguousarray(src, np.float64); keep.append(a); init[k] = _d(a)
        p.init = C.cast(init, C.POINTER(PD))
        if "hor_angle_0000" in b and 1 <= p.shelter_option <= 3:      # tShelter's horizon angle at azimuth 0
            a = np.ascontiguousarray(b["hor_angle_0000"], np.float64); keep.append(a); p.hor_angle_0000 = _d(a)
        _chk(self.L.tribs_gpu_et_init(self.h, C.byref(p)))
        self._et_flag = int(p.i_option != 0)

    def _met(self, rec: dict, cell_record, elev, keep):
        m = MetT()
        m.n_records = len(elev)
        if cell_record is not None:
            cr = np.ascontiguousarray(cell_record, np.int32); keep.append(cr); m.cell_record = _i(cr)
        for member, key in self._MET_KEYS:
            a = np.ascontiguousarray(rec[key], np.float64); keep.append(a); setattr(m, member, _d(a))
        el = np.ascontiguousarray(elev, np.float64); keep.append(el); m.elev = _d(el)
        return m

    def et_step(self, potential=None, components=None, cell_record=None, elev=None):
        """One call of kernel (1). *potential* / *components* are (step dict, record dict) pairs as
        produced by tribs_b200.met.EvapoTransHost, or None when that sweep is not due."""
        if potential is None and components is None:
            return
        keep = []
        s = self._et_step_struct(potential, components, cell_record, elev, keep)
        _chk(self.L.tribs_gpu_et_step(self.h, C.byref(s)))

    def _et_step_struct(self, potential, components, cell_record, elev, keep):
        base = (potential or components)[0]
        s = EtStep()
        for k in ("alphaD", "sinAlpha", "sunaz", "Io", "hour", "dew_hum_flag", "ts_option"):
            setattr(s, k, base[k])
        s.sky_flag_from = base["sky_flag_from"]
        s.intercept_flag = self._et_flag
        if potential is not None:
            st, rec = potential
            s.do_potential = 1
            s.first_call, s.te_met, s.sky_flag_from = st["first_call"], st["te_met"], st["sky_flag_from"]
            mp = self._met(rec, cell_record, elev, keep)
            keep.append(mp)
            s.met_potential = C.pointer(mp)
        if components is not None:
            st, rec = components
            s.do_components = 1
            s.te_eti = st["te_eti"]
            mc = self._met(rec, cell_record, elev, keep)
            keep.append(mc)
            s.met_components = C.pointer(mc)
        return s

    def snow_init(self, basin: dict):
        b, keep = basin, []
        p = SnowParams()
        p.min_sn_temp = float(b["sn_minSnTemp"][0]); p.snliqfrac = float(b["sn_snliqfrac"][0])
        p.hill_albedo_option = int(b["sn_hillAlbedoOption"][0])
        init = (PD * len(SNOW_FIELDS))()
        for k, nm in enumerate(SNOW_FIELDS):
            src = b.get("init_" + nm)
            if src is not None:
                a = np.ascontiguousarray(src, np.float64); keep.append(a); init[k] = _d(a)
        p.init = C.cast(init, C.POINTER(PD))
        _chk(self.L.tribs_gpu_snow_init(self.h, C.byref(p)))

    def snow_step(self, step: dict, rec: dict, hourly: int, cell_record=None, elev=None):
        """One tSnowPack::callSnowPack; *step*/*rec* as produced by met.EvapoTransHost.potential_call."""
        keep = []
        s = self._snow_step_struct(step, rec, cell_record, elev, keep)
        _chk(self.L.tribs_gpu_snow_step(self.h, C.byref(s), int(hourly)))

    def _snow_step_struct(self, step, rec, cell_record, elev, keep):
        s = EtStep()
        for k in ("alphaD", "sinAlpha", "sunaz", "Io", "hour", "dew_hum_flag", "ts_option", "first_call", "te_met",
                  "te_eti", "sky_flag_from"):
            setattr(s, k, step[k])
        s.do_potential = 1
        m = self._met(rec, cell_record, elev, keep)
        keep.append(m)
        s.met_potential = C.pointer(m)
        return s

    def snow_members(self):
        """(albedo, snCanWE) as the last swept cell left them, and the probe rounds of the last call."""
        m, r = np.zeros(2), C.c_int32(0)
        _chk(self.L.tribs_gpu_snow_members(self.h, _d(m), C.byref(r)))
        return m, int(r.value)

    def snow_push(self, arrays: dict, members=None):
        """Snow fields (and, optionally, the carried members) back onto the device: a restart."""
        ptrs = (PD * len(SNOW_FIELDS))()
        mask, keep = 0, []
        for nm, a in arrays.items():
            a = np.ascontiguousarray(a, np.float64); keep.append(a)
            ptrs[SNF[nm]] = _d(a); mask |= 1 << SNF[nm]
        if mask:
            _chk(self.L.tribs_gpu_snow_push(self.h, mask, C.cast(ptrs, C.POINTER(PD))))
        if members is not None:
            m = np.ascontiguousarray(members, np.float64)
            _chk(self.L.tribs_gpu_snow_set_members(self.h, _d(m)))

    def snow_sync(self, names) -> dict:
        out = {nm: np.empty(self.n) for nm in names}
        ptrs = (PD * len(SNOW_FIELDS))()
        mask = 0
        for nm in names:
            ptrs[SNF[nm]] = _d(out[nm]); mask |= 1 << SNF[nm]
        _chk(self.L.tribs_gpu_snow_sync(self.h, mask, C.cast(ptrs, C.POINTER(PD))))
        return out

    def et_sync(self, names) -> dict:
        out = {nm: np.empty(self.n) for nm in names}
        ptrs = (PD * len(ET_FIELDS))()
        mask = 0
        for nm in names:
            ptrs[EF[nm]] = _d(out[nm]); mask |= 1 << EF[nm]
        _chk(self.L.tribs_gpu_et_sync(self.h, mask, C.cast(ptrs, C.POINTER(PD))))
        return out

    def et_push(self, arrays: dict):
        ptrs = (PD * len(ET_FIELDS))()
        mask, keep = 0, []
        for nm, a in arrays.items():
            a = np.ascontiguousarray(a, np.float64); keep.append(a)
            ptrs[EF[nm]] = _d(a); mask |= 1 << EF[nm]
        _chk(self.L.tribs_gpu_et_push(self.h, mask, C.cast(ptrs, C.POINTER(PD))))

    def retrieve_qeff(self, current_time: float) -> np.ndarray:
        out = np.empty(self.n_stream + 1)
        _chk(self.L.tribs_gpu_retrieve_qeff(self.h, current_time, _d(out)))
        return out

    def pending_qeff(self, current_time: float, n_bins: int) -> np.ndarray:
        """[n_stream + 1, n_bins]: the dtReff bins from the current one on (the (TimeInd, Qeff) stacks)."""
        out = np.empty((self.n_stream + 1, n_bins))
        _chk(self.L.tribs_gpu_pending_qeff(self.h, current_time, n_bins, _d(out)))
        return out

    def stream_cells(self) -> np.ndarray:
        out = np.empty(self.n_stream, np.int32)
        _chk(self.L.tribs_gpu_stream_cells(self.h, _i(out)))
        return out

    def results(self, name: str, nbins: int) -> np.ndarray:
        out = np.empty(nbins)
        _chk(self.L.tribs_gpu_get_results(self.h, RESULTS.index(name), _d(out), nbins))
        return out

    def globals(self) -> dict:
        out = np.empty(len(GLOBALS))
        _chk(self.L.tribs_gpu_get_globals(self.h, _d(out)))
        return dict(zip(GLOBALS, out))

    def levels(self) -> int:
        return self.L.tribs_gpu_sweep_levels(self.h)

    def launch_count(self) -> int:
        return int(self.L.tribs_gpu_launch_count(self.h))

    def set_timing(self, on: bool):
        _chk(self.L.tribs_gpu_set_timing(self.h, int(on)))

    def phase_ms(self, reset=True):
        out = np.zeros(4)
        _chk(self.L.tribs_gpu_phase_ms(self.h, _d(out), int(reset)))
        return dict(zip(["unsat", "sat", "route", "reset"], out))

    def wait(self):
        _chk(self.L.tribs_gpu_wait(self.h))

    def stream_handle(self) -> int:
        return int(self.L.tribs_gpu_stream(self.h) or 0)

    def close(self):
        if self.h:
            self.L.tribs_gpu_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
