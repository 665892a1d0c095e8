/* tribs_gpu.h — C ABI of the B200 water-balance path (libtribs_gpu.so).
 *
 * This is the boundary a host model binds to. It replaces, for the per-timestep Voronoi-cell
 * water balance only, the bodies of these reference member functions (the host keeps their
 * signatures and forwards to the calls below; see INTEGRATION.md for the binding):
 *
 *   tHydroModel::UnSaturatedZone(double dt)      src/tHydro/tHydroModel.cpp:720   -> TRIBS_PHASE_UNSAT
 *   tHydroModel::ResetGW()                       src/tHydro/tHydroModel.cpp:2470  -> TRIBS_PHASE_SAT (first half)
 *   tHydroModel::SaturatedZone(double dtGW)      src/tHydro/tHydroModel.cpp:2998  -> TRIBS_PHASE_SAT
 *   tKinemat::RunHydrologicRouting()             src/tFlowNet/tKinemat.cpp:920    -> TRIBS_PHASE_ROUTE
 *   tFlowNet::SurfaceFlow()                      src/tFlowNet/tFlowNet.cpp:735    -> TRIBS_PHASE_ROUTE
 *   tHydroModel::Reset()                         src/tHydro/tHydroModel.cpp:2441  -> TRIBS_PHASE_RESET
 *   tRainfall::NewRain(double) / per-cell rain   src/tRasTin/tRainfall.cpp:281    -> tribs_step_t.rain*
 *
 * Conventions: plain pointers and sizes, no C++ or torch types. All per-cell arrays are in the
 * reference's sweep order (position in the active part of tMesh::nodeList, == tNode::getID()
 * after tFlowNet's sort). Indices are int32, values FP64. Every entry point returns 0 on
 * success, non-zero on failure with the message available from tribs_gpu_last_error(); the C++
 * shim turns that into the reference's `cout << msg; exit(n)` convention.
 * Single host thread per handle; one handle per GPU (one partition per GPU).
 */
#ifndef TRIBS_GPU_H
#define TRIBS_GPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRIBS_GPU_ABI_VERSION 4   /* 3: tribs_part_t (peer-memory partitions), rain gauges, tribs_gpu_pending_qeff; 4: tribs_channel_t, tribs_gpu_channel_* */

typedef struct tribs_ctx *tribs_handle_t;

/* ------------------------------------------------------------------ dynamic per-cell fields
 * Names are the tCNode members (src/tCNode/tCNode.h:683-920). Old/New pairs are listed once:
 * the "New" id is the current value; "Old" is the value at the start of the step. */
#define TRIBS_FIELDS(X) \
  X(NwtNew) X(MuNew) X(MiNew) X(NtNew) X(NfNew) X(RuNew) X(RiNew) \
  X(NwtOld) X(MuOld) X(MiOld) X(NtOld) X(NfOld) X(RuOld) X(RiOld) \
  X(Qpin) X(Qpout) X(intstorm) X(gwchange) \
  X(srf) X(hsrf) X(sbsrf) X(psrf) X(satsrf) X(esrf) X(srf_hr) X(cumsrf) \
  X(Rain) X(NetPrecipitation) X(EvapSoil) X(EvapDryCanopy) \
  X(SoilMoisture) X(SoilMoistureSC) X(SoilMoistureUNSC) X(RootMoisture) X(RootMoistureSC) \
  X(AvSoilMoisture) X(Recharge) X(UnSatFlowIn) X(UnSatFlowOut) X(Transmissivity) \
  X(hsrfOccur) X(sbsrfOccur) X(psrfOccur) X(satsrfOccur) X(satOccur) X(RechDisch) \
  X(FlowVelocity) X(TTime) X(Qstrm) X(LiqWE) X(IceWE) X(LiqRouted) X(UnSaturatedStorage) X(SaturatedStorage)

#define TRIBS_FIELD_ENUM(name) TRIBS_F_##name,
typedef enum { TRIBS_FIELDS(TRIBS_FIELD_ENUM) TRIBS_N_FIELDS } tribs_field_t;
#define TRIBS_MASK(field) (1ull << (field))
#define TRIBS_MASK_ALL ((TRIBS_N_FIELDS >= 64) ? ~0ull : ((1ull << TRIBS_N_FIELDS) - 1ull))

/* ------------------------------------------------------------------ mesh (tMesh/tFlowNet outputs) */
typedef struct {
  int32_t n_cells;              /* active cells (boundary 0 or 3), sweep order */
  int32_t n_edges;              /* directed Voronoi spokes in the CSR below */
  const int32_t *boundary;      /* [n] kNonBoundary=0 | kStream=3 (Definitions.h:42-45) */
  const double *z;              /* [n] m */
  const double *varea;          /* [n] m^2 */
  const double *flow_slope;     /* [n] slope of tCNode::getFlowEdg() */
  const double *flow_vedglen;   /* [n] Voronoi width of the flow edge, m */
  const int32_t *flow_dest;     /* [n] sweep index of the downslope cell, -1 = outlet node */
  const int32_t *stream_node;   /* [n] sweep index of tCNode::getStreamNode(), -1 = outlet node */
  const double *hillpath;       /* [n] m */
  const double *streampath;     /* [n] m */
  const double *contr_area;     /* [n] m^2 (used for stream cells) */
  const double *ttime;          /* [n] initial tCNode::traveltime, h */
  const int32_t *reach;         /* [n] tCNode::getReach(); may be NULL when not partitioned */
  /* Voronoi neighbour graph: row i lists the spokes of cell i counter-clockwise starting at
   * tNode::getEdg(). Each spoke is the directed edge i -> nbr_col[k]. */
  const int32_t *nbr_rowptr;    /* [n+1] */
  const int32_t *nbr_col;       /* [E] sweep index of the destination, -1 = not an active cell */
  const double *nbr_len;        /* [E] tEdge::getLength() of i->j */
  const double *nbr_vedglen;    /* [E] tEdge::getVEdgLen() of i->j */
  const double *nbr_len_in;     /* [E] getLength() of the complementary edge j->i */
  const double *nbr_vedglen_in; /* [E] getVEdgLen() of j->i (differs after FixVoronoiEdgeWidth) */
  const int32_t *nbr_listpos;   /* [E] position of i->j in the active edge list (flux loop order) */
} tribs_mesh_t;

/* ------------------------------------------------------------------ parameters */
typedef struct {
  /* per-cell soil (tCNode::getKs() ...; src/tHydro/tHydroModel.cpp:669-679) */
  const double *Ks, *ThetaS, *ThetaR, *PoreSize, *AirEBubPres, *DecayF, *SatAnRatio, *UnsatAnRatio;
  const double *bedrock;        /* [n] tCNode::getBedrockDepth(), mm */
  double BasArea;               /* tHydroModel::BasArea */
  double BasArea_flow;          /* tFlowNet::BasArea */
  double IntStormMAX;           /* h */
  double surfaceSoilDepth, rootZoneDepth; /* mm */
  int32_t EToption, Ioption, SnOpt, RdstrOption;
  /* hillslope routing (tFlowNet / tKinemat members) */
  double velcoef, velratio, flowexp, baseflow, kincoef, dtReff, outlet_contr_area;
  /* timer (hours) */
  double tstep, gwstep, etistep, output_interval;
  int32_t res_limit;            /* tFlowResults::limit: bins per hydrograph array */
  int32_t reserved;
} tribs_params_t;

/* host staging of dynamic fields; NULL entries are skipped */
typedef struct {
  double *f[TRIBS_N_FIELDS];
} tribs_state_t;

/* ------------------------------------------------------------------ multi-GPU partition (optional)
 * Replaces, for this path, tGraph's reach partition and tParallel's messages (src/tGraph/tGraph.cpp:401-608,
 * 2035-2155, 2431-2479; src/tParallel/tParallel.cpp:195-246). cell_owner[i] is the partition of sweep cell i
 * (tCNode::getReach() through reach2partition). Cells are laid out partition by partition; every basin sum is
 * formed per partition and the partitions are added in partition order, so a run over nranks GPUs is
 * bit-identical to a run of the same partition layout on one GPU (nranks == 1, n_parts > 1).
 * With nranks > 1 every rank is initialised with the whole basin and advances the cells of partition `rank`;
 * boundary values (lateral outflow of reach outlets, the state a remote upslope cell or Voronoi neighbour
 * reads) are stored straight into the peers' device memory by the kernels and dependent work on a peer is
 * released with system-scope atomics over NVLink: no host exchange inside tribs_gpu_run. The ranks make
 * their memory known to each other once, through tribs_gpu_ipc_export / tribs_gpu_attach_peers (one process
 * per GPU; the host moves the 64-byte handles with whatever it has - MPI_Allgather in the reference's
 * tParallel, torch.distributed here) or tribs_gpu_attach_local (several handles in one process). */
typedef struct {
  int32_t rank, nranks;         /* this handle's partition among nranks GPUs; nranks == 1: one GPU runs all partitions */
  int32_t n_parts;              /* partitions in cell_owner; must equal nranks when nranks > 1 */
  int32_t max_batch_steps;      /* nranks > 1: largest n_steps of tribs_gpu_run (its buffers are exported once) */
  const int32_t *cell_owner;    /* [n_cells] partition of every cell of the basin */
} tribs_part_t;

/* ------------------------------------------------------------------ stepping */
enum {
  TRIBS_PHASE_UNSAT = 1,   /* UnSaturatedZone(dt) */
  TRIBS_PHASE_SAT   = 2,   /* ResetGW(); SaturatedZone(dt_gw) */
  TRIBS_PHASE_ROUTE = 4,   /* RunHydrologicRouting(); tFlowNet::SurfaceFlow() accumulations */
  TRIBS_PHASE_RESET = 8,   /* tHydroModel::Reset() */
  /* partitioned handles: the sweep in two passes around the exchange of the river inflow */
  TRIBS_PHASE_UNSAT_LOCAL = 16,  /* cells whose upslope inputs all live on this rank */
  TRIBS_PHASE_UNSAT_REMOTE = 32  /* cells downstream of a cell owned by another rank */
};

/* GAUGES: tRainfall's gauge forcing (RAINSOURCE 3, src/tRasTin/tRainfall.cpp:305-391): `rain` holds one value per
 * gauge, every cell takes the value of its gauge (tribs_gpu_set_rain_gauges: the static Thiessen map) */
enum { TRIBS_RAIN_KEEP = 0, TRIBS_RAIN_UNIFORM = 1, TRIBS_RAIN_PER_CELL = 2, TRIBS_RAIN_GAUGES = 3 };

typedef struct {
  double current_time;      /* tRunTimer::getCurrentTime() after Advance(), hours */
  double dt;                /* tRunTimer::getTimeStep(), hours */
  double dt_gw;             /* tRunTimer::getGWTimeStep(), hours */
  int32_t elapsed_steps;    /* tRunTimer::getElapsedSteps(currentTime) */
  int32_t hourly_reset;     /* !(fmod(currentTime - tstep, etistep)), tHydroModel.cpp:700-702 */
  int32_t rain_mode;        /* TRIBS_RAIN_* : tRainfall::NewRain for this step */
  int32_t reserved;
  double rain_uniform;      /* mm/h, rain_mode == UNIFORM */
  const double *rain;       /* host [n] mm/h, rain_mode == PER_CELL; host [n_gauges], rain_mode == GAUGES */
  const double *qstrm;      /* host [n] tCNode::Qstrm (stream cells) or NULL: only read if flowexp>0 */
  double qstrm_outlet;      /* OutletNode->getQstrm() */
} tribs_step_t;

/* basin accumulators returned by tribs_gpu_get_globals (tHydroModel private sums) */
enum { TRIBS_G_Stok, TRIBS_G_TotRain, TRIBS_G_TotGWchange, TRIBS_G_TotMoist, TRIBS_G_psrf_last,
       TRIBS_G_hillvel_last,
       /* tWaterBalance::BasinStorages[0], [1], [5] (tWaterBalance.cpp:283-337): running volumes, m3 */
       TRIBS_G_BasinRain, TRIBS_G_BasinEvap, TRIBS_G_BasinRunoff, TRIBS_N_GLOBALS };

/* tFlowResults arrays kept on the device (src/tFlowNet/tFlowResults.h:74-96) */
enum { TRIBS_R_mhydro, TRIBS_R_HsrfRout, TRIBS_R_SbsrfRout, TRIBS_R_PsrfRout, TRIBS_R_SatsrfRout,
       TRIBS_R_crr, TRIBS_R_rmax, TRIBS_R_rmin, TRIBS_R_frac, TRIBS_R_msm, TRIBS_R_msmRt,
       TRIBS_R_msmU, TRIBS_R_mgw, TRIBS_R_sat, TRIBS_R_qunsat, TRIBS_N_RESULTS };

/* Build the device mirror: SoA FP64 state, CSR neighbour graph, sweep schedule. Copies
 * everything it needs; the host arrays may be freed afterwards. `part` may be NULL. */
int tribs_gpu_init(const tribs_mesh_t *mesh, const tribs_params_t *prm, const tribs_state_t *init,
                   const tribs_part_t *part, tribs_handle_t *out);

/* Run the phases in `phase_mask` (TRIBS_PHASE_* ORed, executed in the reference's order
 * UNSAT, SAT, ROUTE, RESET) for the step described by `s`. Asynchronous on the handle's
 * compute stream; tribs_gpu_sync / the getters synchronise. */
int tribs_gpu_step(tribs_handle_t h, uint32_t phase_mask, const tribs_step_t *s);

/* Pipelined batch: run `n_steps` whole reference time steps (step k executes the phases in
 * phase_masks[k]; every mask must contain UNSAT, ROUTE and RESET, SAT where the host's
 * fmod(currentTime, GWSTEP) test says so) in ONE persistent launch. Cells advance through
 * space-time as soon as their own inputs exist (upslope outflow of the same step, own and
 * downslope state of the previous phase, Voronoi neighbours around a saturated-zone phase), so
 * the long within-step dependency chain of one step overlaps with the next steps instead of
 * serialising them. Results are bit-identical to n_steps single tribs_gpu_step calls. Forcing
 * for all steps must be known up front (steps[k].rain_mode / rain_uniform / rain); requires
 * flowexp == 0 (no channel feedback into hillslope velocities inside a batch).
 * The per-step lateral inflows are kept for tribs_gpu_retrieve_qeff_steps. */
int tribs_gpu_run(tribs_handle_t h, int n_steps, const tribs_step_t *steps, const uint32_t *phase_masks);
/* out[k*(n_stream+1) + slot] for the steps of the last tribs_gpu_run */
int tribs_gpu_retrieve_qeff_steps(tribs_handle_t h, int n_steps, double *out);

/* Copy the fields selected by `field_mask` (TRIBS_MASK(...)) back into host SoA staging in
 * sweep order; the host shim scatters them into tCNode. Blocks until the data is on the host. */
int tribs_gpu_sync(tribs_handle_t h, uint64_t field_mask, tribs_state_t *host);

/* Overwrite device fields from host staging (restart / external modification of tCNode). */
int tribs_gpu_push(tribs_handle_t h, uint64_t field_mask, const tribs_state_t *host);

/* Lateral inflow the channel router sees this step: what tKinemat::RetrieveQeff
 * (tKinemat.cpp:1504) returns for every stream cell, in sweep order of the stream cells, and
 * for the outlet node in the last slot. `out` has tribs_gpu_n_stream()+1 entries. Consumes
 * the bin at the end of a dtReff interval exactly like the reference. */
int tribs_gpu_retrieve_qeff(tribs_handle_t h, double current_time, double *out);
/* The device's form of the per-stream-cell (TimeInd, Qeff) stacks (tKinemat.cpp:1023-1067, what a restart
 * has to carry for the channel router): out[slot * n_bins + k] = content of the dtReff bin
 * ceil(current_time / dtReff) + k of stream slot `slot` (slots as for tribs_gpu_retrieve_qeff) after the
 * ROUTE phase of the step that ended at current_time. */
int tribs_gpu_pending_qeff(tribs_handle_t h, double current_time, int n_bins, double *out);
int tribs_gpu_n_stream(tribs_handle_t h);
int tribs_gpu_stream_cells(tribs_handle_t h, int32_t *sweep_index_out);

/* tFlowResults arrays (`which` = TRIBS_R_*), `n` <= res_limit bins. */
int tribs_gpu_get_results(tribs_handle_t h, int which, double *out, int n);
int tribs_gpu_get_globals(tribs_handle_t h, double *out /* TRIBS_N_GLOBALS */);

/* The static part of gauge rainfall: cell_gauge[i] = index of the rain gauge whose Thiessen polygon holds
 * sweep cell i (tRainfall::setRainInput's nearest-station search / the .sdf station list). Afterwards a step
 * with rain_mode GAUGES uploads n_gauges values instead of n_cells. */
int tribs_gpu_set_rain_gauges(tribs_handle_t h, int32_t n_gauges, const int32_t *cell_gauge);

/* ---- partitioned runs, peer memory (tribs_part_t.nranks > 1). After tribs_gpu_init on every rank:
 * tribs_gpu_ipc_export fills the 64-byte handle of this rank's exported memory; the host gathers the handles
 * of all ranks (rank order) and passes them to tribs_gpu_attach_peers on every rank. tribs_gpu_attach_local
 * is the same for `nranks` handles that live in one process (they may share a device: each then uses
 * 1/n of its SMs' resident blocks). tribs_gpu_run then advances all partitions together: every rank calls it
 * with the same steps; it returns without waiting (peers are awaited on the device). The hydrograph arrays,
 * basin accumulators and per-step lateral inflows are complete on every rank after the call. */
#define TRIBS_IPC_HANDLE_BYTES 64
int tribs_gpu_ipc_export(tribs_handle_t h, void *handle_out /* TRIBS_IPC_HANDLE_BYTES */);
int tribs_gpu_attach_peers(tribs_handle_t h, const void *handles /* nranks x TRIBS_IPC_HANDLE_BYTES, rank order */);
int tribs_gpu_attach_local(const tribs_handle_t *handles, int nranks);
/* device-index range [begin, end) of the cells this handle advances (the whole basin unless nranks > 1) */
int tribs_gpu_owned_range(tribs_handle_t h, int32_t *begin, int32_t *end);

/* ---- partitioned runs, single steps with a host exchange (the reference's own message pattern: sendQpin /
 * sendOverlap / sendNwt become the three halo kinds below, moved by the host with NCCL between
 * pack on the owner and unpack on the reader). idx_dev / buf_dev are DEVICE pointers; idx holds device-order
 * indices (tribs_gpu_device_index). Kinds: QPOUT = new lateral outflow of a reach outlet (1 value,
 * unpack also marks it as published for the REMOTE pass of the sweep), STATE = the 7 New state
 * values of cells that others read as their downslope cell, NWT = NwtNew of Voronoi neighbours
 * across the cut (before a SAT phase). */
enum { TRIBS_HALO_QPOUT = 0, TRIBS_HALO_STATE = 1, TRIBS_HALO_NWT = 2 };
int tribs_gpu_set_partition(tribs_handle_t h, const int32_t *cell_owner, int rank);
int tribs_gpu_halo_pack(tribs_handle_t h, int kind, const int32_t *idx_dev, int n, double *buf_dev);
int tribs_gpu_halo_unpack(tribs_handle_t h, int kind, const int32_t *idx_dev, int n, const double *buf_dev);
int tribs_gpu_set_global(tribs_handle_t h, int which, double value);   /* e.g. psrf_last from its owner */
int tribs_gpu_padded_cells(tribs_handle_t h);

/* ---- state of a handle
 * tribs_gpu_snapshot keeps one device-side copy of everything the path evolves (all dynamic, ET, snow fields,
 * pending hillslope inflow bins, tFlowResults arrays, basin accumulators, and - after tribs_gpu_channel_init - the channel
 * router's levels, discharges and OutletHlev); tribs_gpu_rollback returns to it.
 * tribs_gpu_save_state / tribs_gpu_load_state write / read the same content as a binary checkpoint on a step
 * boundary (what tRestart / tCNode::writeRestart do for this path, src/tSimulator/tRestart.cpp; per-cell arrays
 * in sweep order, so the file does not depend on the partition layout).
 * tribs_gpu_reserve_batch sizes the buffers of tribs_gpu_run for n_steps ahead of the first batch. */
int tribs_gpu_snapshot(tribs_handle_t h);
int tribs_gpu_rollback(tribs_handle_t h);
int tribs_gpu_save_state(tribs_handle_t h, const char *path);
int tribs_gpu_load_state(tribs_handle_t h, const char *path);
int tribs_gpu_reserve_batch(tribs_handle_t h, int n_steps);

/* ---- channel routing: the kinematic-wave half of tKinemat::RunRoutingModel
 * (src/tFlowNet/tKinemat.cpp:803-897: InitializeStreamReach 1136, AssignLateralInflux 1296, AssignQin 1561,
 * KinematWave 1788 / SolveForTwoNodeReach 1736, ComputeQout 1582, UpdateStreamVars 1617) on the device, fed by the
 * lateral inflows the hillslope routing of the same handle produced (what RetrieveQeff 1504 hands the router).
 * The network is tKinemat's own lists: reaches in the order of NodesLstH / NodesLstO / NNodes, each the chain of cells
 * from its head to its outlet (cell index n_cells = the basin outlet node); per entry the node's tCNode::getChannelWidth(),
 * and getFlowEdg()->getLength() / getSlope() of its flow edge (ignored for the last entry of a reach).
 * OPTPERCOLATION and OPTRESERVOIR other than 0 are refused. On a partitioned handle every rank routes the whole
 * network from the combined inflows (a few thousand nodes), so all ranks hold the same channel state. */
typedef struct {
  int32_t n_reach;
  int32_t percolation_option;   /* OPTPERCOLATION, must be 0 */
  int32_t reservoir_option;     /* OPTRESERVOIR, must be 0 */
  int32_t pad_;
  const int32_t *reach_off;     /* [n_reach + 1] */
  const int32_t *reach_node;    /* [reach_off[n_reach]] */
  const double *width, *length, *slope;   /* per entry of reach_node */
  double roughness;             /* CHANNELROUGHNESS (tKinemat::Roughness) */
  double uniform_width;         /* tKinemat::Width: > 0 overrides the per-node widths */
  double dt_s;                  /* tRunTimer::getTimeStep() * 3600 */
  const double *hlevel0, *qstrm0;   /* [n_cells + 1] tCNode::getHlevel / getQstrm at the start, or NULL for zeros */
  const double *outlet_hlev0;       /* [n_reach] tKinemat::OutletHlev, or NULL */
} tribs_channel_t;
int tribs_gpu_channel_init(tribs_handle_t h, const tribs_channel_t *net);
/* n_steps >= 1: the steps of the last tribs_gpu_run / _run_et / _run_snow batch; n_steps == 0: the single step whose
 * TRIBS_PHASE_ROUTE ran last. qout[k] = tKinemat::Qout after step k (the value SurfaceFlow 740-757 writes to <base>_Outlet.qout). */
int tribs_gpu_channel_route(tribs_handle_t h, int32_t n_steps, double *qout);
/* Hlevel / Qstrm / FlowVelocity per stream slot (tribs_gpu_stream_cells order) + the basin outlet, OutletHlev per reach,
 * counters[7] = Newton iterations, line-search backtracks, reaches that exceeded MAXITS since init, then the device clock
 * cycles thread 0 spent in the band solve, the line-search sums, the function sums and the whole kernel. Pointers may be NULL. */
int tribs_gpu_channel_sync(tribs_handle_t h, double *hlevel, double *qstrm, double *velocity, double *outlet_hlev,
                           int64_t *counters);

/* Introspection used by tests and the bench. */
int tribs_gpu_sweep_levels(tribs_handle_t h);            /* depth of the Qpin dependency DAG */
int tribs_gpu_device_index(tribs_handle_t h, int32_t *dev_index_of_sweep /* [n_cells] */);
int64_t tribs_gpu_launch_count(tribs_handle_t h);        /* kernels launched so far */
/* device time (ms, CUDA events on the compute stream) accumulated per phase since the last call
 * with reset!=0: out[0]=unsat out[1]=sat out[2]=route out[3]=reset; enable with set_timing(1) */
int tribs_gpu_set_timing(tribs_handle_t h, int enabled);
int tribs_gpu_phase_ms(tribs_handle_t h, double *out4, int reset);

int tribs_gpu_wait(tribs_handle_t h);                    /* cudaStreamSynchronize */
void *tribs_gpu_stream(tribs_handle_t h);                /* cudaStream_t of the compute stream */
int tribs_gpu_destroy(tribs_handle_t h);
const char *tribs_gpu_last_error(void);
int tribs_gpu_abi_version(void);


/* ==================================================================== kernel (1): ET + interception
 * The per-cell loops of tEvapoTrans::callEvapoPotential (src/tHydro/tEvapoTrans.cpp:562-831:
 * met assembly, betaFunc/betaFuncT 2911-3010, energyBalance 2281-2357 with inLongWave 2004,
 * inShortWave/DirectDiffuse 1751-1993, aeroResist/stomResist 1442-1512, rtsafe_mod_energy /
 * FunctionAndDerivative / ForceRestore 2391-2716, setToNode 3039-3076) and of
 * tEvapoTrans::callEvapoTrans (980-1197: ComputeETComponents with tIntercept::InterceptGray /
 * InterceptRutter, src/tHydro/tIntercept.cpp:214-459). What the reference does once per call
 * stays with the host and arrives in tribs_et_step_t: calendar hour, sun geometry
 * (SetSunVariables 1567-1660), the weather-station records of the hour (newHydroMetData
 * 3962-4010) and the flags derived from record 0.
 * Covered options: OPTEVAPOTRANS 1-3, OPTINTERCEPT 0-2, GFLUXOPTION 1-2, OPTRADSHELT 0-4 (1-3 with
 * tShelter's per-cell values, see tribs_et_params_t), OPTLANDUSE 0, OPTSNOW 0/1 (1: tribs_gpu_snow_*
 * below), station or gridded met. Anything else fails in tribs_gpu_et_init. */
#define TRIBS_ET_FIELDS(X) \
  X(PotEvap) X(ActEvap) X(SurfTemp) X(SoilTemp) X(NetRad) X(GFlux) X(HFlux) X(LFlux) \
  X(LongRadIn) X(LongRadOut) X(ShortRadIn) X(ShortRadIn_dir) X(ShortRadIn_dif) X(ShortRadSlope) \
  X(ShortAbsbVeg) X(ShortAbsbSoi) X(AirTemp) X(DewTemp) X(RelHumid) X(VapPressure) X(SkyCover) \
  X(WindSpeed) X(AirPressure) X(CumHrsSun) X(AvEvapFract) X(SheltFact) X(CumRSin) \
  X(EvapWetCanopy) X(EvapoTrans) X(CumTotEvap) X(CumBarEvap) X(AvET) X(InterceptLoss) \
  X(CanStorage) X(CumIntercept) X(StormLength) X(CanopyStorVol)
#define TRIBS_ET_FIELD_ENUM(name) TRIBS_ET_##name,
typedef enum { TRIBS_ET_FIELDS(TRIBS_ET_FIELD_ENUM) TRIBS_N_ET_FIELDS } tribs_et_field_t;

typedef struct {
  const double *flow_slope;     /* [n] slope of tCNode::getFlowEdg() (tEvapoTrans.cpp:617) */
  const double *aspect;         /* [n] tCNode::getAspect(), radians */
  const double *vol_heat_cond;  /* [n] tCNode::getVolHeatCond() */
  const double *soil_heat_cap;  /* [n] tCNode::getSoilHeatCap() */
  const double *soil_cutoff;    /* [n] tCNode::getSoilCutoff() */
  const double *root_cutoff;    /* [n] tCNode::getRootCutoff() */
  const int32_t *land_class;    /* [n] row of land_table (tCNode::getLandUse()) */
  int32_t n_land_classes;
  const double *land_table;     /* [n_land_classes][15]: column k = tInvariant::getLandProp(k), k = 1..14 */
  int32_t et_option, i_option, gflux_option, shelter_option;   /* OPTEVAPOTRANS OPTINTERCEPT GFLUXOPTION OPTRADSHELT */
  int32_t lu_option, snow_option;                              /* OPTLANDUSE (must be 0), OPTSNOW (1: tribs_gpu_snow_init follows) */
  double metstep_min;           /* METSTEP (tEvapoTrans::timeStep) */
  double etistep_h;             /* tRunTimer::getEtIStep() */
  double tlinke, temp_lapse;    /* TLINKE, TEMPLAPSE */
  double max_interstorm;        /* tIntercept::maxInterStormPeriod (INTSTORMMAX/2, tIntercept.cpp:44-45) */
  const double *const *init;    /* NULL, or TRIBS_N_ET_FIELDS pointers (NULL entries = 0) */
  /* OPTRADSHELT 1-3 (tShelter): init[TRIBS_ET_SheltFact] carries the sky-view factor of every cell and
   * hor_angle_0000 [n] is tCNode::getHorAngle0000() in radians - the only horizon angle
   * tEvapoTrans::aboveHorizon (tEvapoTrans.cpp:2080-2270) ends up using (its sector test is always
   * true, so the first case decides). NULL for OPTRADSHELT 0 and >= 4. */
  const double *hor_angle_0000;
} tribs_et_params_t;

/* one record per weather station (METDATAOPTION 1) or per cell (gridded, METDATAOPTION 2) */
typedef struct {
  int32_t n_records;
  const int32_t *cell_record;   /* [n] record of each cell; NULL = as in the previous call */
  const double *air_temp, *dew_temp, *rel_humid, *vap_press, *atm_press, *wind_speed, *sky_cover;
  const double *rad_global;     /* tHydroMet::getRadGlobal */
  const double *elev;           /* tHydroMet::getOther(): station elevation for TEMPLAPSE */
} tribs_met_t;

typedef struct {
  int32_t do_potential;         /* callEvapoPotential is due (currentTime == met_hour) */
  int32_t do_components;        /* callEvapoTrans is due (currentTime == eti_hour) */
  int32_t intercept_flag;       /* 2nd argument of callEvapoTrans */
  double alphaD, sinAlpha, sunaz, Io;     /* tEvapoTrans members after SetSunVariables */
  int32_t hour;                 /* currentTime[3] */
  int32_t first_call;           /* hourlyTimeStep == 0: soil/surface temperature start (3989-3992) */
  int32_t dew_hum_flag;         /* 0 RH, 1 dew point, 2 vapour pressure observed (3994-4003) */
  int32_t ts_option;            /* tEvapoTrans::tsOption (> 1: global radiation observed) */
  int32_t sky_flag_from;        /* sweep index from which the computed sky cover applies (724-727) */
  double te_met, te_eti;        /* getElapsedMETSteps / getElapsedETISteps(currentTime) as double */
  const tribs_met_t *met_potential;   /* record read by callEvapoPotential */
  const tribs_met_t *met_components;  /* record read by callEvapoTrans (met clock already advanced) */
} tribs_et_step_t;

int tribs_gpu_et_init(tribs_handle_t h, const tribs_et_params_t *prm);
int tribs_gpu_et_step(tribs_handle_t h, const tribs_et_step_t *s);
/* host[k] (k < TRIBS_N_ET_FIELDS) receives / supplies field k in sweep order when bit k is set */
/* tribs_gpu_run with kernel (1) inside the batch: et_steps[k] (or NULL) is what tribs_gpu_et_step would be called
 * with before the sweep of step k (Simulator::SurfaceHydroProcesses, tSimul.cpp:419-461). The records of all ET
 * steps of the batch are copied before the call returns. OPTSNOW 0 only (the snow pack runs between batches). */
int tribs_gpu_run_et(tribs_handle_t h, int n_steps, const tribs_step_t *steps, const uint32_t *phase_masks,
                     const tribs_et_step_t *const *et_steps);
int tribs_gpu_et_sync(tribs_handle_t h, uint64_t et_field_mask, double *const *host);
int tribs_gpu_et_push(tribs_handle_t h, uint64_t et_field_mask, const double *const *host);

/* ------------------------------------------------------------------ snow pack (OPTSNOW 1)
 * The cell loop of tSnowPack::callSnowPack (src/tHydro/tSnowPack.cpp:249-757: pack mass and
 * energy balance 944-1365, 1940-2001, turbulent exchange 1373-1480, short wave on snow 1574-1748,
 * melt routing, peak/persistence trackers 1052-1135), which under OPTSNOW replaces both ET sweeps
 * at every METSTEP (tSimul.cpp:464-492); cells without snow take the ET path of kernel (1).
 * With a canopy (OPTINTERCEPT 1/2) callSnowIntercept (769-934) and computeSub (1492-1568) run
 * first. There the reference reads two class members before the current cell has written them -
 * `albedo` (computeSub) and, under HILLALBOPT 2, `snCanWE` (inShortWaveSn) - so they hold what the
 * last cell BEFORE it in sweep order (or the last of the previous call) left behind. The device
 * reproduces that order dependence with a last-writer scan over sweep positions: store-free probe
 * rounds of the cell update until the carried values are settled (normally two), then one commit.
 * The energy-balance closure error (Uerror, CumUerror: round-off noise that the reference also
 * carries from cell to cell) is written per cell from the cell's own balance. */
#define TRIBS_SNOW_FIELDS(X) \
  X(SnTempC) X(CrustAge) X(DensityAge) X(EvapoTransAge) X(SnSub) X(SnEvap) X(DU) X(Unode) X(Uerror) \
  X(SnLHF) X(SnSHF) X(SnGHF) X(SnPHF) X(SnRLout) X(SnRLin) X(SnRSin) \
  X(IntSWE) X(IntPrec) X(IntSnUnload) X(IntSub) X(CumIntSub) X(CumIntUnl) \
  X(CumLHF) X(CumMelt) X(CumSHF) X(CumSnSub) X(CumSnEvap) X(CumPHF) X(CumRLin) X(CumRLout) X(CumGHF) \
  X(CumUerror) X(CumHrsSnow) X(PersTimeMax) X(PersTime) X(PeakSWE) X(PeakSWETemp) X(InitPackTime) \
  X(InitPackTimeTemp) X(PeakPackTime)
#define TRIBS_SNOW_FIELD_ENUM(name) TRIBS_SN_##name,
typedef enum { TRIBS_SNOW_FIELDS(TRIBS_SNOW_FIELD_ENUM) TRIBS_N_SNOW_FIELDS } tribs_snow_field_t;

typedef struct {
  double min_sn_temp;           /* MINSNTEMP */
  double snliqfrac;             /* SNLIQFRAC */
  int32_t hill_albedo_option;   /* HILLALBOPT 0..2 */
  int32_t reserved;
  const double *const *init;    /* NULL, or TRIBS_N_SNOW_FIELDS pointers (NULL entries = 0) */
} tribs_snow_params_t;

/* tribs_gpu_run with the snow pack inside the batch (OPTSNOW 1, OPTINTERCEPT 0): snow_steps[k] (or NULL) and
 * hourly_time_step[k] are what tribs_gpu_snow_step would be called with before the sweep of step k. */
int tribs_gpu_run_snow(tribs_handle_t h, int n_steps, const tribs_step_t *steps, const uint32_t *phase_masks,
                       const tribs_et_step_t *const *snow_steps, const int32_t *hourly_time_step);
/* after tribs_gpu_et_init (with snow_option = 1 in its parameters) */
int tribs_gpu_snow_init(tribs_handle_t h, const tribs_snow_params_t *prm);
/* one callSnowPack: `s` as for tribs_gpu_et_step (do_potential set, met_potential = the record of
 * the hour, te_met and te_eti filled); hourly_time_step = tEvapoTrans::hourlyTimeStep of the call */
int tribs_gpu_snow_step(tribs_handle_t h, const tribs_et_step_t *s, int32_t hourly_time_step);
int tribs_gpu_snow_sync(tribs_handle_t h, uint64_t snow_field_mask, double *const *host);
/* the tSnowPack members that outlive a call (members[0] = albedo, members[1] = snCanWE, as the last
 * swept cell left them; what a restart has to carry) and the number of probe rounds of the last call
 * (0 without a canopy); either pointer may be NULL */
int tribs_gpu_snow_members(tribs_handle_t h, double *members, int32_t *probe_rounds);
/* restart: the snow fields and the two carried members back onto the device (tSnowPack::readRestart's part) */
int tribs_gpu_snow_push(tribs_handle_t h, uint64_t snow_field_mask, const double *const *host);
int tribs_gpu_snow_set_members(tribs_handle_t h, const double *members);

#ifdef __cplusplus
}
#endif
#endif
