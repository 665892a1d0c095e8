// tribs_gpu.cu — libtribs_gpu.so: the C ABI of include/tribs_gpu.h and the sm_100a kernels.
//
// Data layout in HBM
//   * every per-cell quantity is its own FP64 (or int32) array of length n_pad: structure of
//     arrays, so a warp touching 32 consecutive cells issues fully coalesced 256-byte loads;
//   * cells are stored in "device order": the reference's sweep order stably re-sorted by the
//     depth ("level") of the cell in the within-step lateral-inflow DAG (cell -> flow_dest),
//     each level padded to a multiple of 32. A cell's upslope contributors therefore always
//     live at smaller device indices, in earlier 32-cell tiles;
//   * the Voronoi edge lists are a CSR neighbour graph (rowptr/col + 4 edge arrays) with column
//     indices already in device order; the upslope contributors are a second CSR whose rows
//     are sorted by sweep position so that Qpin is summed in the reference's order.
//
// Kernels (all FP64, HBM/latency bound; no tensor-core work exists on this path)
//   unsat_sweep_kernel   UnSaturatedZone: one persistent cooperative launch; warps claim 32-cell
//                        tiles in device order and each thread waits (acquire-poll on a per-cell
//                        epoch flag) for its upslope cells' new Qpout before advancing its column.
//   sat_prepare_kernel   water-table elevation + transmissivity of every cell.
//   sat_cells_kernel     per-cell CSR gather of the Darcy fluxes over its Voronoi edges (both
//                        directions evaluated locally, no atomics), ascending sort + sequential sum
//                        (bit-reproducible, order independent), then the saturated state machine.
//   route_* kernels      hillslope runoff -> stream-node time bins (warp-shuffle segmented sums over
//                        contributors pre-sorted by hillslope path) and basin hydrograph bins.
//   reduce_sums_kernel   fixed-order reduction of the per-tile basin accumulators.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cub/device/device_scan.cuh>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <algorithm>
#include <chrono>
#include <numeric>
#include <string>
#include <unordered_map>
#include <vector>

#include "cell_update.cuh"
#include "et_update.cuh"
#include "snow_update.cuh"

namespace cg = cooperative_groups;
using namespace tribs;

// ------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(const std::string &m) { g_err = m; return 1; }
#define CK(call)                                                                            \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" +  \
                  std::to_string(__LINE__) + ")");                                          \
  } while (0)

constexpr int kTile = 32;
constexpr int kMaxRing = 4096;

// ------------------------------------------------------------------------------------ context
struct __align__(16) PubWord { double value; long long epoch; };   // see pub_store / pub_load
constexpr int kSegBlock = 1024;
struct SegBoundary { long long first_key, last_key; double first_val, last_val; };
struct RouteConst {
  double velcoef, velratio, flowexp, baseflow, kincoef, dt_reff, outlet_contr_area;
  double tstep, output_interval;
  int32_t res_limit, ring;
};

constexpr int kMaxParts = 16;
constexpr int kQueueClasses = 4;   // ready queues of a batch: hi, and lo by kind of phase (pipeline.cuh)
// Memory other ranks map (cudaIpc) or address directly (several handles in one process): one allocation per
// handle, carved by a bump allocator in the same order on every rank, so an address on a peer is
// peer_base + (my_address - my_base).
struct Slab { char *base = nullptr; size_t cap = 0, used = 0; };
// hillslope routing entries of one partition (see route_seg_kernel)
struct SegPart {
  int32_t *cell = nullptr, *slot = nullptr;
  double *path = nullptr, *area = nullptr;
  struct SegBoundary *bnd = nullptr, *bnd2 = nullptr;   // boundary records of the first / second reduce level
  int n_seg = 0, blocks = 0, blocks2 = 0;
};
struct PipeState;
struct ChannelState;
struct tribs_ctx {
  int device = 0;
  // partitions (tribs_part_t): cells are laid out partition by partition; sums are formed per partition
  int n_parts = 1, n_ranks = 1;
  std::vector<int> part_begin, part_end;     // device-index range of every partition (tile aligned)
  std::vector<int> my_parts;                 // partitions this handle computes
  int own_begin = 0, own_end = 0;            // device-index range this handle computes
  std::vector<int32_t> sweep_owner;          // partition of every sweep cell (empty: one partition)
  Slab slab{};
  std::vector<SegPart> seg_parts;
  int seg_level2_min = 64;                   // partitions with more 1024-entry blocks reduce their boundary runs in two levels
  int tile_bins = 1, rec_stride = 15;        // per-tile routing record: 14 + 5 * tile_bins doubles
  double *d_tile_rec = nullptr;              // [n_own_tiles][rec_stride] single-step routing records
  double *d_part_vec = nullptr, *d_block_part = nullptr, *d_sums_part = nullptr;
  struct StepBins *d_stepbins = nullptr;     // single-step routing: output bins of the step
  int32_t *d_one = nullptr;                  // device constant 1
  int32_t *d_gauge_of_cell = nullptr;        // device order, tribs_gpu_set_rain_gauges
  double *d_gauge_vals = nullptr;
  int n_gauges = 0;
  PipeState *pipe = nullptr;   // pipelined-batch state (tribs_gpu_run), created on first use
  ChannelState *chan = nullptr;   // channel network and state (tribs_gpu_channel_init)
  struct StateSnapshot *snap = nullptr;   // tribs_gpu_snapshot
  struct EtState *et = nullptr; // kernel (1) state (tribs_gpu_et_init)
  int n = 0, n_pad = 0, n_edges = 0, n_levels = 0, n_tiles = 0, n_stream = 0, n_late = 0;
  int max_degree = 0, segment_cells = 0;
  double gwstep_h = 0.0;
  cudaStream_t stream = nullptr;
  // host maps
  std::vector<int32_t> dev_of_sweep, sweep_of_dev, stream_cells_sweep;
  // device static
  StaticView sv{};
  DynView dv{};
  std::vector<void *> allocs;
  int32_t *d_up_rowptr = nullptr, *d_up_col = nullptr, *d_tile_counter = nullptr, *d_tile_order = nullptr;
  int32_t *d_tile_hi = nullptr;
  // partitioned runs (one reach partition per GPU)
  int32_t *d_owned = nullptr, *d_tiles_local = nullptr, *d_tiles_remote = nullptr;
  int n_tiles_local = 0, n_tiles_remote = 0, rank = 0;
  bool topo_order = true;     // ascending tile index is a topological order of the lateral-flow DAG
  bool partitioned = false;
  int n_hi_tiles = 0;
  PubWord *d_pub = nullptr;
  int32_t *d_abort = nullptr, *d_sweep_of_dev = nullptr, *d_dev_of_sweep = nullptr;
  int32_t *d_late_rowptr = nullptr, *d_late_col = nullptr;
  double *d_tile_sums = nullptr, *d_globals = nullptr, *d_wt_elev = nullptr, *d_transm = nullptr;
  double *d_stage = nullptr;
  int32_t *d_flags = nullptr;  // [0] gather overflow
  // routing
  int32_t *d_sn_dev = nullptr;
  double *d_ring = nullptr, *d_qeff_now = nullptr, *d_results = nullptr;
  int route_window = 0;
  RouteConst rc{};
  ModelConst mc{};
  int last_sweep_dev = -1;   // device index of the last cell in sweep order
  int last_hill_dev = -1;    // device index of the last hillslope cell in sweep order
  long long epoch = 0;
  int unsat_blocks = 0, unsat_threads = 128;
  bool gw_dirty = false;
  bool new_valid = true;     // false between Reset() and the next zone update: New == Old there
  bool at_boundary = true;   // no zone update since init / the last Reset(): Old holds the current state
  int64_t launches = 0;
  // timing
  bool timing = false;
  double phase_ms[4] = {0, 0, 0, 0};
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> pending;
  std::vector<cudaEvent_t> event_pool;
};

template <class T>
static int dev_alloc(tribs_ctx *c, T **p, size_t count) {
  void *q = nullptr;
  CK(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
  c->allocs.push_back(q);
  *p = (T *)q;
  return 0;
}
// memory peers address: from the slab when the handle has one (nranks > 1), else an ordinary allocation
template <class T>
static int shared_alloc(tribs_ctx *c, T **p, size_t count) {
  if (!c->slab.base) return dev_alloc(c, p, count);
  const size_t bytes = ((std::max<size_t>(count, 1) * sizeof(T)) + 255) & ~(size_t)255;
  if (c->slab.used + bytes > c->slab.cap) { g_err = "internal: exported memory slab too small"; return 1; }
  *p = (T *)(c->slab.base + c->slab.used);
  c->slab.used += bytes;
  return 0;
}
template <class T>
static int dev_upload(tribs_ctx *c, T **p, const std::vector<T> &h) {
  if (dev_alloc(c, p, h.size())) return 1;
  if (!h.empty()) CK(cudaMemcpy(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

// ------------------------------------------------------------------------------------ device helpers
// Publication word of the chained sweep: the new Qpout of a cell and the epoch (sweep number)
// it belongs to travel in ONE 16-byte transaction, so a consumer that sees the epoch also holds
// the value: no fence, no L1 invalidation, one L2 round trip per poll (the scheme of
// decoupled look-back scans).

__device__ __forceinline__ void pub_store(PubWord *p, double v, long long epoch) {
  asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(__double_as_longlong(v)), "l"(epoch) : "memory");
}
__device__ __forceinline__ bool pub_load(const PubWord *p, long long epoch, double &v) {
  long long a, b;
  asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
  v = __longlong_as_double(a);
  return b == epoch;
}

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// A consumer that has waited this long (wall time on the device, not a poll count) for an input that
// never arrives raises the handle's abort flag; the host turns it into an error at the next getter.
constexpr unsigned long long kSpinTimeoutNs = 60ull * 1000000000ull;
__device__ __forceinline__ double warp_sum_fixed(double v) {
  // butterfly in a fixed order: the result depends only on the 32 inputs
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------ UNSAT
// Persistent, cooperative (co-resident) grid. Tiles are claimed in increasing device order, so
// the lowest unfinished tile is always held by a running warp whose inputs are complete: the
// poll below cannot deadlock. A bounded spin plus a global abort flag guarantees termination
// even if that invariant were ever broken.
__global__ void __launch_bounds__(128)
unsat_sweep_kernel(StaticView s, DynView d, ModelConst m, StepConst t, const int32_t *__restrict__ up_rowptr,
                   const int32_t *__restrict__ up_col, PubWord *pub, long long epoch, int n_tiles,
                   const int32_t *__restrict__ tile_order, int32_t *tile_counter, int32_t *abort_flag, double *tile_sums, double *globals,
                   int last_sweep_dev) {
  const int lane = threadIdx.x & 31;
  for (;;) {
    int tile = 0;
    if (lane == 0) tile = atomicAdd(tile_counter, 1);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= n_tiles) break;
    tile = tile_order[tile];   // levels in ascending order: inputs always belong to tiles claimed earlier
    const int i = tile * kTile + lane;
    UnsatSums sums{0.0, 0.0, 0.0, 0.0};
    if (s.boundary[i] >= 0 && s.owned[i]) {
      UnsatCellInputs u;
      unsat_cell_load(u, s, d, i, m, t);
      // lateral inflow: new Qpout of the upslope cells swept earlier, summed in sweep order
      double qpin = 0.0;
      const int k1 = up_rowptr[i + 1];
      for (int k = up_rowptr[i]; k < k1; k++) {
        const PubWord *src = pub + up_col[k];
        double v;
        unsigned spins = 0, ns = 32;
        unsigned long long t_first = 0;
        while (!pub_load(src, epoch, v)) {
          __nanosleep(ns);
          if (ns < 256) ns += ns;
          if (((++spins) & 255u) == 0u) {
            if (*(volatile int32_t *)abort_flag) { v = 0.0; break; }
            const unsigned long long t_now = global_ns();
            if (!t_first) t_first = t_now;
            else if (t_now - t_first > kSpinTimeoutNs) { atomicExch(abort_flag, 1); v = 0.0; break; }
          }
        }
        qpin += v;
      }
      int state;
      const double qout = unsat_cell_advance(u, qpin, &state);
      pub_store(pub + i, qout, epoch);   // the downslope cell is waiting for this
      unsat_cell_store(u, d, i, qpin, m, t, sums);
      if (i == last_sweep_dev) globals[TRIBS_G_psrf_last] = u.c.psrf;
    }
    // per-tile partial sums of the basin accumulators, reduced later in tile order
    double a = warp_sum_fixed(sums.stok), b = warp_sum_fixed(sums.tot_rain);
    double c2 = warp_sum_fixed(sums.tot_gw), e = warp_sum_fixed(sums.tot_moist);
    if (lane == 0) {
      tile_sums[4 * tile + 0] = a;
      tile_sums[4 * tile + 1] = b;
      tile_sums[4 * tile + 2] = c2;
      tile_sums[4 * tile + 3] = e;
    }
  }
}

// contributions scattered by cells that are swept *after* their target (non-topological pairs):
// visible in Qpin until Reset(), never used by the physics (tHydroModel.cpp:2078-2079, 2454)
__global__ void unsat_late_qpin_kernel(DynView d, const int32_t *late_rowptr, const int32_t *late_col, int n_pad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  double q = d.f[TRIBS_F_Qpin][i];
  for (int k = late_rowptr[i]; k < late_rowptr[i + 1]; k++) q += d.f[TRIBS_F_Qpout][late_col[k]];
  if (late_rowptr[i + 1] > late_rowptr[i]) d.f[TRIBS_F_Qpin][i] = q;
}

// ------------------------------------------------------------------------------------ SAT
__global__ void __launch_bounds__(256)
sat_prepare_kernel(StaticView s, DynView d, double *wt_elev, double *transm, int n_pad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad || s.boundary[i] < 0) return;
  sat_prepare_cell(s, d.f[TRIBS_F_NwtOld], i, wt_elev, transm);
}

// 8 blocks of 128 threads per SM = 64 registers per thread (124 without the bound): the kernel waits for its neighbour
// gathers, more resident warps hide more of that. 1M-cell valley, per groundwater step (prepare + cells): 1.01 ms unbounded,
// 0.88 ms at 6 blocks (80 registers), 0.83 ms at 8 (tools/sat_ab.py).
#ifndef TRIBS_SAT_MIN_BLOCKS
#define TRIBS_SAT_MIN_BLOCKS 8
#endif
template <int MAXC>
__global__ void __launch_bounds__(128, (MAXC <= 16 ? TRIBS_SAT_MIN_BLOCKS : 1))   // the wide-degree variants would spill
sat_cells_kernel(StaticView s, DynView d, ModelConst m, StepConst t, const double *__restrict__ wt_elev,
                 const double *__restrict__ transm, int n_pad, double *tile_sums, int32_t *flags,
                 const double *__restrict__ globals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  SatSums sums{0.0, 0.0, 0.0};
  t.psrf_last = globals[TRIBS_G_psrf_last];
  if (i < n_pad && s.boundary[i] >= 0 && s.owned[i]) {
    int overflow = 0;
    double gw = sat_gather<MAXC>(s, d, i, wt_elev, transm, &overflow);
    if (overflow) flags[0] = 1;
    sat_cell(s, d, i, gw, m, t, sums);
  }
  double a = warp_sum_fixed(sums.stok), b = warp_sum_fixed(sums.tot_gw), c2 = warp_sum_fixed(sums.tot_moist);
  if ((threadIdx.x & 31) == 0 && i < n_pad) {
    int tile = i / kTile;
    tile_sums[4 * tile + 0] = a;
    tile_sums[4 * tile + 1] = 0.0;
    tile_sums[4 * tile + 2] = b;
    tile_sums[4 * tile + 3] = c2;
  }
}

// ------------------------------------------------------------------------------------ ROUTE
// (1) per-cell part of tKinemat::RunHydrologicRouting: FlowVelocity and the travel time of
//     runoff-producing hillslope cells (tKinemat.cpp:958-970, 1084).
__device__ __forceinline__ double hill_velocity(const StaticView &s, const DynView &d, const RouteConst &rc,
                                                int sn, double qstrm_outlet) {
  if (!(rc.flowexp > 0.0)) return rc.velcoef / rc.velratio;
  double q, carea;
  if (sn < 0) { q = qstrm_outlet; carea = rc.outlet_contr_area; }
  else {
    int dn = s.flow_dest[sn];
    carea = s.contr_area[sn];
    q = (d.f[TRIBS_F_Qstrm][sn] + (dn >= 0 ? d.f[TRIBS_F_Qstrm][dn] : qstrm_outlet)) / 2.;
  }
  double flowout = (!q) ? rc.baseflow * carea / rc.outlet_contr_area : q;
  return rc.kincoef * pow(flowout / carea, rc.flowexp);
}

// (2) hillslope runoff -> (stream node, dtReff bin), tKinemat.cpp:941-1067.
//     Static "segment order": every stream slot (stream cell, or the outlet node) owns a
//     contiguous run of entries = the stream cell itself (hillslope path 0) followed by its
//     hillslope contributors sorted by hillslope path. The dtReff bin is monotone in the path, so
//     (slot, bin) keys are non-decreasing along the whole array and every key is ONE contiguous
//     run: a streaming reduce-by-key. Each 1024-entry block does a warp-shuffle segmented scan,
//     adds the runs that lie strictly inside it straight into the slot's ring of bins (unique
//     owner, no atomics) and hands its first and last (possibly cut) runs to a sequential fix-up.

__device__ __forceinline__ double slot_velocity(const StaticView &s, const DynView &d, const RouteConst &rc,
                                                int node, double qstrm_outlet) {
  return hill_velocity(s, d, rc, node, qstrm_outlet);
}

// One 1024-entry block of a streaming reduce-by-key: keys are non-decreasing along the entries, every key is ONE
// contiguous run. Warp-shuffle segmented scan, carries across warps; the runs that lie strictly inside the block go
// straight into the slot's ring of bins (unique owner, no atomics), the first and the last (possibly cut) run are handed
// to the next level as the block's boundary record.
constexpr long long kSegPadKey = 0x7fffffffffffffffLL;   // padding entries sort last and carry nothing
__device__ __forceinline__ void seg_block_reduce(long long key, double v, const RouteConst &rc, double *ring, SegBoundary *bnd) {
  __shared__ long long sh_key_last[32], sh_key_first[32];
  __shared__ double sh_acc_last[32], sh_carry[32];
  __shared__ int sh_single[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // warp-level segmented inclusive scan
  long long prev_key = __shfl_up_sync(0xffffffffu, key, 1);
  const bool head = (lane == 0) || (prev_key != key);
  const unsigned heads = __ballot_sync(0xffffffffu, head);
  double acc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double up = __shfl_up_sync(0xffffffffu, acc, o);
    // lane-o is in my run iff no head lies in (lane-o, lane]
    unsigned between = (lane >= o) ? ((heads >> (lane - o + 1)) & ((1u << o) - 1u)) : 1u;
    if (lane >= o && between == 0u) acc += up;
  }
  if (lane == 31) { sh_key_last[warp] = key; sh_acc_last[warp] = acc; }
  if (lane == 0) { sh_key_first[warp] = key; sh_single[warp] = (heads == 1u); }
  __syncthreads();
  if (threadIdx.x == 0) {
    // carry_in[w]: what the run that starts warp w already collected in earlier warps
    sh_carry[0] = 0.0;
    for (int w = 1; w < 32; w++) {
      const double run_end = sh_acc_last[w - 1] + (sh_single[w - 1] ? sh_carry[w - 1] : 0.0);
      sh_carry[w] = (sh_key_first[w] == sh_key_last[w - 1]) ? run_end : 0.0;
    }
    bnd[blockIdx.x] = SegBoundary{-1, -1, 0.0, 0.0};
  }
  __syncthreads();
  // lanes of the first run of their warp inherit the carry
  const bool in_first_run = ((heads & ((2u << lane) - 1u)) == 1u);
  if (in_first_run) acc += sh_carry[warp];
  // run tails
  long long next_key = __shfl_down_sync(0xffffffffu, key, 1);
  if (lane == 31) next_key = (warp < 31) ? sh_key_first[warp + 1] : ~key;  // block end always closes the run
  const bool tail = (next_key != key);
  const long long block_first_key = sh_key_first[0];
  const long long block_last_key = sh_key_last[31];
  if (tail && key != kSegPadKey) {
    const bool is_first_run = (key == block_first_key);
    const bool is_last_run = (key == block_last_key);
    if (is_first_run) { bnd[blockIdx.x].first_key = key; bnd[blockIdx.x].first_val = acc; }
    if (is_last_run && !is_first_run) { bnd[blockIdx.x].last_key = key; bnd[blockIdx.x].last_val = acc; }
    if (!is_first_run && !is_last_run) {
      const int slot = (int)(key >> 32), bin = (int)(key & 0xffffffffLL);
      ring[(size_t)slot * rc.ring + (bin % rc.ring)] += acc;
    }
  }
}


__global__ void __launch_bounds__(kSegBlock)
route_seg_kernel(StaticView s, DynView d, RouteConst rc, const int32_t *__restrict__ seg_cell,
                 const int32_t *__restrict__ seg_slot, const double *__restrict__ seg_path,
                 const double *__restrict__ seg_area, const int32_t *__restrict__ slot_dev, int n_seg, double now,
                 double qstrm_outlet, const double *__restrict__ srf_of_cell, double *ring, SegBoundary *bnd) {
  const int e = blockIdx.x * kSegBlock + threadIdx.x;
  const double dtreff = rc.dt_reff;
  const int cur = (int)ceil(now / dtreff);
  const bool at_end = ((int)ceil(now / dtreff) == (int)floor(now / dtreff));
  const double same_box_div = at_end ? (rc.tstep * 3600.) : ((cur * dtreff - now + rc.tstep) * 3600.);
  long long key = kSegPadKey;
  double v = 0.0;
  if (e < n_seg) {
    const int slot = seg_slot[e];
    const double hv = slot_velocity(s, d, rc, slot_dev[slot], qstrm_outlet);
    const double tt = seg_path[e] / hv / 3600.;
    const int bin = (int)ceil((now + tt) / dtreff);
    key = ((long long)slot << 32) | (unsigned)bin;
    const double srf = srf_of_cell[seg_cell[e]];   // the live array, or a per-step snapshot of it (batches)
    if (srf > 0.0 && !(srf > 999999) && s.owned[seg_cell[e]]) {
      double vr = srf * seg_area[e] / 1000.;
      v = (bin == cur) ? vr / same_box_div : vr / (dtreff * 3600.);
    }
  }
  seg_block_reduce(key, v, rc, ring, bnd);
}

// second level: the boundary records of the first (two items per block: its first and its last run) are themselves a
// non-decreasing key sequence; reduce them the same way, so that the sequential stitch only sees a few dozen records
__global__ void __launch_bounds__(kSegBlock)
route_seg_level2_kernel(RouteConst rc, const SegBoundary *__restrict__ bnd1, int n_blocks1, double *ring, SegBoundary *bnd2) {
  const int e = blockIdx.x * kSegBlock + threadIdx.x;
  long long key = kSegPadKey;
  double v = 0.0;
  if (e < 2 * n_blocks1) {
    const SegBoundary x = bnd1[e >> 1];
    if (e & 1) {       // a block that is one single run has no "last" record: continue its key with nothing
      key = x.last_key >= 0 ? x.last_key : x.first_key;
      v = x.last_key >= 0 ? x.last_val : 0.0;
    } else {
      key = x.first_key;
      v = x.first_val;
    }
    if (key < 0) { key = kSegPadKey; v = 0.0; }
  }
  seg_block_reduce(key, v, rc, ring, bnd2);
}

// sequential stitch of the block-boundary runs (block order = entry order) of one partition; the
// latch of what the channel router reads (tKinemat::RetrieveQeff, tKinemat.cpp:1504-1551) follows in
// route_latch_kernel once every partition has added its runs
__global__ void __launch_bounds__(1024)
route_seg_fixup_kernel(RouteConst rc, const SegBoundary *__restrict__ bnd, int n_blocks, double *ring) {
  // The records are staged through shared memory 512 at a time (coalesced loads); one thread
  // merges them in order into a list of finished (key, sum) runs -- sequential by nature but free
  // of global memory -- and then all threads add the finished runs to the rings in parallel (the
  // keys of finished runs are distinct, so there are no conflicts).
  __shared__ SegBoundary sh_bnd[512];
  __shared__ long long sh_out_key[1025];
  __shared__ double sh_out_val[1025];
  __shared__ long long sh_ckey;
  __shared__ double sh_cval;
  __shared__ int sh_nout;
  if (threadIdx.x == 0) { sh_ckey = -1; sh_cval = 0.0; }
  for (int base = 0; base < n_blocks; base += 512) {
    const int cnt = min(512, n_blocks - base);
    __syncthreads();
    if ((int)threadIdx.x < cnt) sh_bnd[threadIdx.x] = bnd[base + threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
      long long ckey = sh_ckey;
      double cval = sh_cval;
      int nout = 0;
      auto flush = [&](long long key, double val) {
        if (key < 0) return;
        sh_out_key[nout] = key; sh_out_val[nout] = val; nout++;
      };
      for (int b = 0; b < cnt; b++) {
        const SegBoundary x = sh_bnd[b];
        if (x.first_key >= 0) {
          if (x.first_key == ckey) cval += x.first_val;
          else { flush(ckey, cval); ckey = x.first_key; cval = x.first_val; }
        }
        if (x.last_key >= 0) { flush(ckey, cval); ckey = x.last_key; cval = x.last_val; }
      }
      if (base + 512 >= n_blocks) { flush(ckey, cval); ckey = -1; cval = 0.0; }
      sh_ckey = ckey; sh_cval = cval; sh_nout = nout;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < sh_nout; k += blockDim.x) {
      const long long key = sh_out_key[k];
      const int slot = (int)(key >> 32), bin = (int)(key & 0xffffffffLL);
      ring[(size_t)slot * rc.ring + (bin % rc.ring)] += sh_out_val[k];
    }
  }
}
// what the channel router reads this step; the bin is consumed at the end of its dtReff interval
__global__ void __launch_bounds__(256)
route_latch_kernel(RouteConst rc, int n_slots, double now, double *ring, double *qeff_now) {
  const int cur = (int)ceil(now / rc.dt_reff);
  const bool at_end = ((int)ceil(now / rc.dt_reff) == (int)floor(now / rc.dt_reff));
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_slots) return;
  double *p = ring + (size_t)k * rc.ring + (cur % rc.ring);
  qeff_now[k] = *p;
  if (at_end) *p = 0.0;
}

// (3) per-cell tail of RunHydrologicRouting (FlowVelocity, travel time; tKinemat.cpp:958-970,
//     1084) fused with the tFlowNet::SurfaceFlow accumulations (tFlowNet.cpp:758-865) and tWaterBalance.
//     Every 32-cell tile leaves one record per step: 13 basin terms (sums by a fixed butterfly, max, min)
//     and its runoff volumes x 5 types in the `K` output bins from its first one on (the cells of a tile are
//     of one (segment, depth) bucket, so their travel times are close; K is planned at init). The records
//     are written by the last task of a tile in its step (pipelined batches: no per-cell snapshot of 15
//     fields any more) or by route_tiles_kernel (single steps), then reduced in a fixed order:
//     tiles -> blocks of a partition (tile_reduce_kernel) -> partition (part_reduce_kernel) -> partitions
//     in partition order (combine_results_kernel). The grouping depends on the partition layout only, so
//     one GPU and N GPUs give the same bits.
constexpr int kRouteWarps = 8;
constexpr int kMeans = 13;  // crr frac msm msmRt msmU sat mgw qunsat rmax rmin + basin rain/evap/runoff volumes
constexpr int kRecBins = 14;          // record layout: [0..12] the terms above, [13] first bin (-1: no runoff), then [5][K]
constexpr int kReduceBlocksMax = 64;
// tWaterBalance (tWaterBalance.cpp:116-345) rides on this sweep: it sits at the same place in the
// reference's step (Simulator::UpdateWaterBalance, after SurfaceFlow, before Reset) and reads the
// same end-of-step fields.
struct WaterBalanceArgs {
  double *uns_storage, *sat_storage;   // tCNode::UnSaturatedStorage / SaturatedStorage
  const double *evapotrans;            // tCNode::EvapoTranspiration (kernel 1), or nullptr
  double gwstep_h;
  int32_t gw_time;                     // fmod(currentTime, GWSTEP) == 0
};
struct PartVecs { const double *v[kMaxParts]; int n; };
struct StepBins { int32_t bin_base, cur_bin; };
struct RouteStep {
  double now, qstrm_outlet, hv_last;
  int32_t gw_time;       // Simulator::GW_label == 0 (tSimul.cpp:511)
  int32_t satsrf_zero;   // batches: a step without a saturated-zone phase has no groundwater runoff (no Reset() in between)
};

// one warp = one tile; `rec` (lane 0 writes) has 14 + 5K doubles
__device__ __forceinline__ void route_tile_record(const StaticView &s, const DynView &d, const RouteConst &rc,
                                                  const ModelConst &m, const WaterBalanceArgs &wb, const RouteStep &rs,
                                                  int i, bool act, int lane, int K, double *rec, int32_t *flags) {
  double vol[5] = {0, 0, 0, 0, 0};
  int init = 0, end = -1;
  double w_init = 0, w_mid = 0, w_end = 0;
  double means[kMeans];
#pragma unroll
  for (int k = 0; k < kMeans; k++) means[k] = 0.0;
  const double now = rs.now;
  if (act) {
    const int bnd = s.boundary[i];
    const double area = s.varea[i], areaf = area / m.bas_area_flow;
    const double srf = d.f[TRIBS_F_srf][i];
    const double dcalc = rc.tstep, dres = rc.output_interval, sec = rc.output_interval * 3600.0;
    double tt = d.f[TRIBS_F_TTime][i];
    if (bnd == 0) {
      const double hv = hill_velocity(s, d, rc, s.stream_node[i], rs.qstrm_outlet);
      d.f[TRIBS_F_FlowVelocity][i] = hv;
      if (srf > 0.0) { tt = s.hillpath[i] / hv / 3600.; d.f[TRIBS_F_TTime][i] = tt; }
    } else
      d.f[TRIBS_F_FlowVelocity][i] = rs.hv_last;
    if (srf > 0.0) {
      if (rc.flowexp > 0.0) {
        const double hv0 = rc.velcoef / rc.velratio, sv0 = rc.velcoef;
        tt = (s.hillpath[i] / hv0 + s.streampath[i] / sv0) / 3600.0;
        d.f[TRIBS_F_TTime][i] = tt;
      }
      vol[0] = srf * area / 1000.0;
      vol[1] = d.f[TRIBS_F_hsrf][i] * area / 1000.0;
      vol[2] = d.f[TRIBS_F_sbsrf][i] * area / 1000.0;
      vol[3] = d.f[TRIBS_F_psrf][i] * area / 1000.0;
      vol[4] = rs.satsrf_zero ? 0.0 : d.f[TRIBS_F_satsrf][i] * area / 1000.0;
      init = (int)floor((tt + .01 * dcalc + now) / dres);
      end = (int)floor((tt + .99 * dcalc + now) / dres);
      if (init == end) w_init = 1.0 / sec;
      else {
        const double dint = (init + 1) * dres - (now + tt);
        w_init = dint / dcalc / sec;
        if (end - init == 1) w_end = (dcalc - dint) / dcalc / sec;
        else {
          w_mid = dres / dcalc / sec;
          w_end = ((now + tt + dcalc) - end * dres) / dcalc / sec;
        }
      }
    }
    const double w = rc.tstep / rc.output_interval;
    const double rain = d.f[TRIBS_F_Rain][i];
    double cs = s.cos_gw[i];
    if (cs < 1E-9) cs = 1.E-9;
    means[0] = areaf * rain * rc.tstep / rc.output_interval;
    means[1] = (rain > 0.0) ? areaf * w : 0.0;
    means[2] = areaf * d.f[TRIBS_F_SoilMoistureSC][i] * w;
    means[3] = areaf * d.f[TRIBS_F_RootMoistureSC][i] * w;
    means[4] = areaf * d.f[TRIBS_F_SoilMoistureUNSC][i] * w;
    means[5] = (d.f[TRIBS_F_SoilMoistureSC][i] >= 1) ? areaf * w : 0.0;
    means[6] = areaf * (d.f[TRIBS_F_NwtNew][i] / cs) * w;
    means[7] = areaf * (d.f[TRIBS_F_Qpout][i] - d.f[TRIBS_F_Qpin][i]) * 1.E-6 / area * w;
    means[8] = rain;
    means[9] = rain;
    {   // UnSaturatedBalance / SaturatedBalance / BasinStorage for this cell
      const double uns_min = rc.tstep * 60.0, sat_min = wb.gwstep_h * 60.0, rech = d.f[TRIBS_F_Recharge][i];
      const double melt = d.f[TRIBS_F_LiqRouted][i] * 10.0;
      double inf = d.f[TRIBS_F_NetPrecipitation][i];
      inf = (d.f[TRIBS_F_LiqWE][i] + d.f[TRIBS_F_IceWE][i] > 1e-4) ? melt : inf + melt;
      const double et = d.f[TRIBS_F_EvapSoil][i] + d.f[TRIBS_F_EvapDryCanopy][i];
      const double del_u = inf + d.f[TRIBS_F_UnSatFlowIn][i] - d.f[TRIBS_F_UnSatFlowOut][i] - rech - et - srf / uns_min;
      wb.uns_storage[i] = wb.uns_storage[i] + del_u * (uns_min / 60.0) * area / 1000.0;
      if (rs.gw_time) {
        const double del_g = 0.0 - 0.0 + (rech - srf / sat_min) * area / 1000.0;
        wb.sat_storage[i] = wb.sat_storage[i] + del_g * (sat_min / 60.0);
      }
      const double a = area / 1000.0;
      means[10] = rain * a * uns_min / 60.0;
      means[11] = (wb.evapotrans ? wb.evapotrans[i] : 0.0) * a * uns_min / 60.0;
      means[12] = srf * a;
    }
  }
  // the sums: fixed-order butterfly; min/max: order free
#pragma unroll
  for (int k = 0; k < 8; k++) means[k] = warp_sum_fixed(means[k]);
  double mx = act ? means[8] : -1.0e300, mn = act ? means[9] : 1.0e300;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
#pragma unroll
  for (int k = 10; k < kMeans; k++) means[k] = warp_sum_fixed(means[k]);
  means[8] = mx; means[9] = mn;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < kMeans; k++) rec[k] = means[k];
  }
  // binned volumes of the runoff-producing lanes, from the tile's first bin on
  const bool has = act && end >= init;
  int b0 = has ? init : 0x7fffffff, b1 = has ? end : -0x7fffffff;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    b0 = min(b0, __shfl_xor_sync(0xffffffffu, b0, o));
    b1 = max(b1, __shfl_xor_sync(0xffffffffu, b1, o));
  }
  if (b0 == 0x7fffffff) {
    if (lane == 0) rec[13] = -1.0;
    return;
  }
  if (b1 - b0 >= K) {   // more bins than planned at init: the run must not go on silently
    if (lane == 0) flags[1] = 1;
    b1 = b0 + K - 1;
  }
  if (lane == 0) rec[13] = (double)b0;
  for (int k = 0; k < K; k++) {
    const int b = b0 + k;
    double wgt = 0.0;
    if (has && b >= init && b <= end) wgt = (b == init) ? w_init : ((b == end) ? w_end : w_mid);
    const bool any = b <= b1;       // uniform
#pragma unroll
    for (int q = 0; q < 5; q++) {
      // value * fraction / seconds in the reference; fraction / seconds folded into one weight here
      const double v = any ? warp_sum_fixed(vol[q] * wgt) : 0.0;
      if (lane == 0) rec[kRecBins + q * K + k] = v;
    }
  }
}

// single steps: one warp per tile of this handle's range
__global__ void __launch_bounds__(128)
route_tiles_kernel(StaticView s, DynView d, RouteConst rc, ModelConst m, WaterBalanceArgs wb, RouteStep rs, int cell_begin,
                   int n_tiles, int last_hill_dev, int K, double *recs, double *globals, int32_t *flags) {
  const int lane = threadIdx.x & 31;
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (t >= n_tiles) return;
  // velocity the stream cells inherit: that of the last hillslope cell in sweep order
  rs.hv_last = last_hill_dev >= 0 ? hill_velocity(s, d, rc, s.stream_node[last_hill_dev], rs.qstrm_outlet)
                                  : rc.velcoef / rc.velratio;
  if (t == 0 && lane == 0) globals[TRIBS_G_hillvel_last] = rs.hv_last;
  const int i = cell_begin + t * kTile + lane;
  const bool act = s.boundary[i] >= 0 && s.owned[i];
  route_tile_record(s, d, rc, m, wb, rs, i, act, lane, K, recs + (size_t)t * (kRecBins + 5 * K), flags);
}

// records of one partition's tiles -> G block partials per step. Block (g, step): warp w takes the tiles
// g*8+w, g*8+w+8G, ... of the partition in that order and adds their records into its shared slab (lanes
// over the record's entries); the slabs are added in warp order. part[(step*G + g)*stride + k].
__global__ void __launch_bounds__(kRouteWarps * 32)
tile_reduce_kernel(const double *__restrict__ recs, size_t step_stride, int tile0, int n_tiles, int K, int window,
                   const StepBins *__restrict__ sb, double *part, int32_t *flags) {
  extern __shared__ double sh[];  // [kRouteWarps][5*window + kMeans]
  const int stride = 5 * window + kMeans, R = kRecBins + 5 * K;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x, g = blockIdx.x, step = blockIdx.y;
  double *slab = sh + warp * stride;
  for (int k = lane; k < stride; k += 32) slab[k] = 0.0;
  if (lane == 0) { slab[5 * window + 8] = -1.0e300; slab[5 * window + 9] = 1.0e300; }
  __syncwarp();
  const int bin_base = sb[step].bin_base;
  const double *base = recs + (size_t)step * step_stride;
  for (int t = g * kRouteWarps + warp; t < n_tiles; t += G * kRouteWarps) {
    const double *rec = base + (size_t)(tile0 + t) * R;
    if (lane < kMeans) {
      const double v = rec[lane];
      double *dst = slab + 5 * window + lane;
      if (lane == 8) *dst = fmax(*dst, v);
      else if (lane == 9) *dst = fmin(*dst, v);
      else *dst += v;
    }
    const double b0d = rec[13];
    if (b0d >= 0.0) {
      const int off0 = (int)b0d - bin_base;
      for (int e = lane; e < 5 * K; e += 32) {
        const int q = e / K, k = e - q * K;
        const double v = rec[kRecBins + e];
        if (v != 0.0) {
          const int off = off0 + k;
          if (off < 0 || off >= window) flags[1] = 1;
          else slab[q * window + off] += v;
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();
  for (int k = threadIdx.x; k < stride; k += blockDim.x) {
    double acc = sh[k];
    if (k == 5 * window + 8) { for (int wv = 1; wv < kRouteWarps; wv++) acc = fmax(acc, sh[wv * stride + k]); }
    else if (k == 5 * window + 9) { for (int wv = 1; wv < kRouteWarps; wv++) acc = fmin(acc, sh[wv * stride + k]); }
    else { for (int wv = 1; wv < kRouteWarps; wv++) acc += sh[wv * stride + k]; }
    part[((size_t)step * G + g) * stride + k] = acc;
  }
}

// G block partials -> the partition's vector of a step: one warp per value, fixed order
__global__ void __launch_bounds__(256)
part_reduce_kernel(const double *__restrict__ part, int G, int window, int n_steps, double *vec) {
  const int stride = 5 * window + kMeans;
  const int lane = threadIdx.x & 31;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= (long long)n_steps * stride) return;
  const int step = (int)(w / stride), k = (int)(w - (long long)step * stride);
  const bool is_max = (k == 5 * window + 8), is_min = (k == 5 * window + 9);
  double acc = is_max ? -1.0e300 : (is_min ? 1.0e300 : 0.0);
  for (int b = lane; b < G; b += 32) {
    const double v = part[((size_t)step * G + b) * stride + k];
    acc = is_max ? fmax(acc, v) : (is_min ? fmin(acc, v) : acc + v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double v = __shfl_xor_sync(0xffffffffu, acc, o);
    acc = is_max ? fmax(acc, v) : (is_min ? fmin(acc, v) : acc + v);
  }
  if (lane == 0) vec[(size_t)step * stride + k] = acc;
}

// The partitions' vectors, added in partition order, go into the tFlowResults arrays step by step
// (tFlowResults::store_* 857-1161; what tParallel::sum/min/max do in write_inter_hyd 531-561, with a fixed
// order). vec[r] may live on another rank. One thread per (runoff type, output bin) and one per basin term.
__global__ void __launch_bounds__(256)
combine_results_kernel(PartVecs pv, const StepBins *__restrict__ sb, int n_steps, int window, int limit, int bin_lo,
                       int n_bins, double *results, double *globals) {
  const int stride = 5 * window + kMeans;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid < 5 * n_bins) {
    const int type = tid / n_bins, bin = bin_lo + tid - type * n_bins;
    if (bin < 0 || bin >= limit) return;
    double acc = results[(size_t)type * limit + bin];
    for (int st = 0; st < n_steps; st++) {
      const int off = bin - sb[st].bin_base;
      if (off < 0 || off >= window) continue;
      double tot = 0.0;
      for (int r = 0; r < pv.n; r++) tot += pv.v[r][(size_t)st * stride + type * window + off];
      if (tot != 0.0) acc += tot;
    }
    results[(size_t)type * limit + bin] = acc;
    return;
  }
  const int q = tid - 5 * n_bins;
  if (q >= kMeans) return;
  const int map[10] = {TRIBS_R_crr, TRIBS_R_frac, TRIBS_R_msm, TRIBS_R_msmRt, TRIBS_R_msmU,
                       TRIBS_R_sat, TRIBS_R_mgw, TRIBS_R_qunsat, TRIBS_R_rmax, TRIBS_R_rmin};
  for (int st = 0; st < n_steps; st++) {
    double tot = (q == 8) ? -1.0e300 : ((q == 9) ? 1.0e300 : 0.0);
    for (int r = 0; r < pv.n; r++) {
      const double v = pv.v[r][(size_t)st * stride + 5 * window + q];
      tot = (q == 8) ? fmax(tot, v) : ((q == 9) ? fmin(tot, v) : tot + v);
    }
    if (q >= 10) { globals[TRIBS_G_BasinRain + (q - 10)] += tot; continue; }
    const int cur_bin = sb[st].cur_bin;     // store_saturation etc. only write when init == end
    if (cur_bin < 0 || cur_bin >= limit) continue;
    double *dst = results + (size_t)map[q] * limit + cur_bin;
    if (q == 8) { if (tot > *dst) *dst = tot; }
    else if (q == 9) { if (tot <= *dst) *dst = tot; }
    else *dst += tot;
  }
}

// basin accumulators of tHydroModel: per-tile partials of one partition -> [phase][4] (block = phase),
// then the partitions in order, phase by phase in the reference's order
__global__ void __launch_bounds__(256)
sums_reduce_kernel(const double *__restrict__ tile_sums, size_t phase_stride, int tile0, int n_tiles, double *out) {
  __shared__ double sh[4][256];
  const double *ts = tile_sums + (size_t)blockIdx.x * phase_stride + (size_t)tile0 * 4;
  double acc[4] = {0, 0, 0, 0};
  for (int tix = threadIdx.x; tix < n_tiles; tix += 256)
    for (int k = 0; k < 4; k++) acc[k] += ts[4 * (size_t)tix + k];
  for (int k = 0; k < 4; k++) sh[k][threadIdx.x] = acc[k];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int k = 0; k < 4; k++) sh[k][threadIdx.x] += sh[k][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x < 4) out[(size_t)blockIdx.x * 4 + threadIdx.x] = sh[threadIdx.x][0];
}
__global__ void combine_sums_kernel(PartVecs pv, const int32_t *__restrict__ phase_is_sat, int n_phases, double *globals) {
  const int k = threadIdx.x;
  if (k >= 4) return;
  const int g[4] = {TRIBS_G_Stok, TRIBS_G_TotRain, TRIBS_G_TotGWchange, TRIBS_G_TotMoist};
  for (int p = 0; p < n_phases; p++) {
    if (k == 1 && phase_is_sat && phase_is_sat[p]) continue;
    double tot = 0.0;
    for (int r = 0; r < pv.n; r++) tot += pv.v[r][(size_t)p * 4 + k];
    globals[g[k]] += tot;
  }
}

// ------------------------------------------------------------------------------------ misc kernels
__global__ void fill_kernel(double *p, double v, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void gauge_fill_kernel(double *rain, const double *__restrict__ vals, const int32_t *__restrict__ gauge_of_cell, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && gauge_of_cell[i] >= 0) rain[i] = vals[gauge_of_cell[i]];
}
__global__ void gather_to_sweep_kernel(const double *__restrict__ src, const int32_t *__restrict__ dev_of_sweep,
                                       double *dst, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[dev_of_sweep[i]];
}
__global__ void scatter_from_sweep_kernel(const double *__restrict__ src, const int32_t *__restrict__ dev_of_sweep,
                                          double *dst, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[dev_of_sweep[i]] = src[i];
}

// ------------------------------------------------------------------------------------ timing helpers
static cudaEvent_t get_event(tribs_ctx *c) {
  if (!c->event_pool.empty()) { cudaEvent_t e = c->event_pool.back(); c->event_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
struct PhaseTimer {
  tribs_ctx *c; int phase; cudaEvent_t a = nullptr, b = nullptr;
  PhaseTimer(tribs_ctx *c_, int p) : c(c_), phase(p) {
    if (c->timing) { a = get_event(c); b = get_event(c); cudaEventRecord(a, c->stream); }
  }
  ~PhaseTimer() {
    if (c->timing) { cudaEventRecord(b, c->stream); c->pending.push_back({phase, {a, b}}); }
  }
};
static void drain_timers(tribs_ctx *c) {
  for (auto &p : c->pending) {
    float ms = 0;
    cudaEventSynchronize(p.second.second);
    cudaEventElapsedTime(&ms, p.second.first, p.second.second);
    c->phase_ms[p.first] += ms;
    c->event_pool.push_back(p.second.first);
    c->event_pool.push_back(p.second.second);
  }
  c->pending.clear();
}

// ------------------------------------------------------------------------------------ init
static inline void swap_old_new(tribs_ctx *c) {
  const int olds[7] = {TRIBS_F_NwtOld, TRIBS_F_MuOld, TRIBS_F_MiOld, TRIBS_F_NtOld, TRIBS_F_NfOld, TRIBS_F_RuOld, TRIBS_F_RiOld};
  const int news[7] = {TRIBS_F_NwtNew, TRIBS_F_MuNew, TRIBS_F_MiNew, TRIBS_F_NtNew, TRIBS_F_NfNew, TRIBS_F_RuNew, TRIBS_F_RiNew};
  for (int k = 0; k < 7; k++) std::swap(c->dv.f[olds[k]], c->dv.f[news[k]]);
}

extern "C" int tribs_gpu_abi_version(void) { return TRIBS_GPU_ABI_VERSION; }
extern "C" const char *tribs_gpu_last_error(void) { return g_err.c_str(); }

extern "C" int tribs_gpu_set_partition(tribs_handle_t c, const int32_t *cell_owner, int rank);
static PipeState *pipe_of(tribs_ctx *c);
static int ensure_pipe(tribs_ctx *c, PipeState *ps, int n_steps);
extern "C" int tribs_gpu_init(const tribs_mesh_t *mesh, const tribs_params_t *prm, const tribs_state_t *init,
                              const tribs_part_t *part, tribs_handle_t *out) {
  if (!mesh || !prm || !out) return fail("tribs_gpu_init: null argument");
  if (part && part->nranks > 1 && !part->cell_owner) return fail("tribs_gpu_init: a partitioned init needs cell_owner");
  // Kernels are loaded when the library is, not at their first launch: a load in the middle of a batch can wait for
  // running kernels, and with several ranks in one process the running kernel may be a barrier waiting for the very
  // rank whose host thread is stuck loading (no effect if the process initialised CUDA before; see also
  // preload_batch_kernels).
  setenv("CUDA_MODULE_LOADING", "EAGER", 0);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("tribs_gpu_init: no CUDA device (this library has no CPU path)");
  const int n = mesh->n_cells;
  if (n <= 0) return fail("tribs_gpu_init: empty mesh");
  const bool dbg = getenv("TRIBS_DEBUG") != nullptr;
  const auto t_init0 = std::chrono::steady_clock::now();
  auto tick = [&](const char *what) {
    if (dbg) fprintf(stderr, "[tribs] init %-28s %8.3f s\n", what,
                     std::chrono::duration<double>(std::chrono::steady_clock::now() - t_init0).count());
  };
  tribs_ctx *c = new tribs_ctx();
  CK(cudaGetDevice(&c->device));
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->n = n;
  c->n_edges = mesh->n_edges;
  // ---- partitions: layout by partition; with nranks > 1 this handle advances partition `rank`
  const int32_t *owner = (part && part->cell_owner) ? part->cell_owner : nullptr;
  c->n_ranks = part ? std::max(1, (int)part->nranks) : 1;
  c->rank = (part && c->n_ranks > 1) ? part->rank : 0;
  c->n_parts = owner ? std::max(1, (int)part->n_parts) : 1;
  if (c->n_ranks > 1 && c->n_parts != c->n_ranks) { delete c; return fail("tribs_gpu_init: n_parts must equal nranks when nranks > 1"); }
  if (c->n_parts > kMaxParts) { delete c; return fail("tribs_gpu_init: more than 16 partitions"); }
  if (c->n_ranks > 1 && (c->rank < 0 || c->rank >= c->n_ranks)) { delete c; return fail("tribs_gpu_init: rank out of range"); }
  if (c->n_ranks > 1 && part->max_batch_steps <= 0) { delete c; return fail("tribs_gpu_init: nranks > 1 needs max_batch_steps"); }
  if (owner) {
    c->sweep_owner.assign(owner, owner + n);
    for (int i = 0; i < n; i++)
      if (owner[i] < 0 || owner[i] >= c->n_parts) { delete c; return fail("tribs_gpu_init: cell_owner out of range"); }
  }
  auto own_of = [&](int i) { return owner ? owner[i] : 0; };

  // ---- sweep schedule: level of every cell in the within-step Qpin DAG
  std::vector<std::vector<int32_t>> early(n), late(n);
  for (int j = 0; j < n; j++) {
    int dst = mesh->flow_dest[j];
    if (dst < 0) continue;
    if (dst >= n) { delete c; return fail("tribs_gpu_init: flow_dest out of range"); }
    if (j < dst) early[dst].push_back(j);
    else if (j > dst) { late[dst].push_back(j); c->n_late++; }
  }
  std::vector<int32_t> level(n, 0);
  int nlev = 0;
  for (int i = 0; i < n; i++) {  // contributors j < i are already final
    int lv = 0;
    for (int j : early[i]) lv = std::max(lv, level[j] + 1);
    level[i] = lv;
    nlev = std::max(nlev, lv + 1);
  }
  c->n_levels = nlev;
  tick("levels");
  // ---- device order. Tiles (32 consecutive device cells) are the unit of scheduling, so a tile must
  //  (a) hold cells with no flow relation among them (a warp never waits on itself) and
  //  (b) hold cells whose inputs become ready at about the same time in EVERY phase; otherwise one
  //      tile couples the bottom of the river of phase p to the top of the river of phase p+1 and
  //      the phases of a pipelined batch serialise.
  // Buckets = (stream segment, hops to the stream): the stream cells are cut, in sweep order
  // (upstream first), into segments of G cells; a hillslope cell belongs to the segment of its
  // stream node and to the depth class "hops along its flow path to the stream". Segments are
  // closed under flow, depth strictly decreases along flow, and within a segment the tile
  // dependency graph is a chain over depth. Device order = (segment, hillslope ridge-first by
  // depth, then the segment's stream cells), so ascending tile index is a topological order.
  std::vector<int32_t> seg(n, 0), depth(n, 0), srank(n, -1);
  {
    int ns = 0, max_depth = 1;
    long n_hill = 0;
    for (int i = n - 1; i >= 0; i--) {
      if (mesh->boundary[i] == 3) depth[i] = 0;
      else {
        const int dst = mesh->flow_dest[i];
        depth[i] = (dst < 0 || dst <= i) ? 1 : depth[dst] + 1;
        max_depth = std::max(max_depth, (int)depth[i]);
        n_hill++;
      }
    }
    for (int i = 0; i < n; i++) if (mesh->boundary[i] == 3) srank[i] = ns++;
    long G = 1;
    // a bucket (segment, depth) then holds about 256 cells: fewer part-filled tiles (measured: 9 % padding
    // instead of 23 % with 64) at the price of coarser tile dependencies, which the fine buckets next to
    // the river (below) take back where it matters
    if (ns > 0 && n_hill > 0) G = std::lround(256.0 * max_depth * (double)ns / (double)n_hill);
    G = std::max(1L, std::min(G, (long)std::max(ns, 1)));
    if (const char *e = getenv("TRIBS_SEGMENT_CELLS")) G = std::max(1, atoi(e));
    const int n_groups = ns > 0 ? (int)((ns + G - 1) / G) : 1;
    for (int i = 0; i < n; i++) {
      const int sn = mesh->boundary[i] == 3 ? i : mesh->stream_node[i];
      seg[i] = (sn >= 0 && srank[sn] >= 0) ? (int)(srank[sn] / G) : n_groups - 1;
    }
    c->segment_cells = (int)G;
  }
  // Near the river the buckets are finer: a hillslope cell within `fine_depth` hops of the stream
  // is grouped with the cells that drain to the same `fine_group` consecutive stream cells. The
  // tiles next to the river then wait for 2-4 stream cells of the previous phase instead of a
  // whole segment, which shortens the critical chain of a pipelined batch from
  // levels + (G+2)*phases to about levels + 5*phases tasks.
  int fine_depth = 3, fine_group = 4;
  if (const char *e = getenv("TRIBS_FINE_DEPTH")) fine_depth = atoi(e);
  if (const char *e = getenv("TRIBS_FINE_GROUP")) fine_group = std::max(1, atoi(e));
  auto fine_key = [&](int i) -> long long {
    if (mesh->boundary[i] == 3 || depth[i] > fine_depth) return 0;
    const int sn = mesh->stream_node[i];
    return (sn >= 0 && srank[sn] >= 0) ? srank[sn] / fine_group : 0x7fffffff;
  };
  auto bucket_key = [&](int i) {
    return mesh->boundary[i] == 3 ? (long long)srank[i] : -(long long)depth[i] * 0x100000000LL + fine_key(i);
  };
  std::vector<int32_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    if (own_of(a) != own_of(b)) return own_of(a) < own_of(b);
    if (seg[a] != seg[b]) return seg[a] < seg[b];
    const int sa = mesh->boundary[a] == 3, sb = mesh->boundary[b] == 3;
    if (sa != sb) return sa < sb;
    return bucket_key(a) < bucket_key(b);
  });
  c->dev_of_sweep.assign(n, -1);
  c->part_begin.assign(c->n_parts, -1);
  c->part_end.assign(c->n_parts, -1);
  {
    int pos = 0;
    for (int k = 0; k < n; k++) {
      const int i = order[k];
      if (k > 0) {
        const int j = order[k - 1];
        if (own_of(i) != own_of(j) || seg[i] != seg[j] || (mesh->boundary[i] == 3) != (mesh->boundary[j] == 3) ||
            bucket_key(i) != bucket_key(j))
          pos = ((pos + kTile - 1) / kTile) * kTile;
      }
      if (c->part_begin[own_of(i)] < 0) c->part_begin[own_of(i)] = pos;
      c->dev_of_sweep[i] = pos++;
    }
    c->n_pad = ((pos + kTile - 1) / kTile) * kTile;
    // ranges are contiguous and ascending; a partition without cells is an empty range
    int next = c->n_pad;
    for (int r = c->n_parts - 1; r >= 0; r--) {
      if (c->part_begin[r] < 0) c->part_begin[r] = next;
      c->part_end[r] = next;
      next = c->part_begin[r];
    }
  }
  const int n_pad = c->n_pad;
  if (c->n_ranks > 1) {
    c->my_parts = {c->rank};
    c->own_begin = c->part_begin[c->rank]; c->own_end = c->part_end[c->rank];
  } else {
    for (int r = 0; r < c->n_parts; r++) c->my_parts.push_back(r);
    c->own_begin = 0; c->own_end = n_pad;
  }
  c->n_tiles = n_pad / kTile;
  c->sweep_of_dev.assign(n_pad, -1);
  for (int i = 0; i < n; i++) c->sweep_of_dev[c->dev_of_sweep[i]] = i;
  tick("device order");
  // tiles on the critical cycle of a pipelined batch (river cells and the cells next to the river)
  {
    std::vector<int32_t> hi(c->n_tiles, 0);
    int hi_depth = 3;
    if (const char *e = getenv("TRIBS_HI_DEPTH")) hi_depth = atoi(e);
    for (int i = 0; i < n; i++)
      if (mesh->boundary[i] == 3 || depth[i] <= hi_depth) hi[c->dev_of_sweep[i] / kTile] = 1;
    c->n_hi_tiles = 0;
    for (int v : hi) c->n_hi_tiles += v;
    if (dev_upload(c, &c->d_tile_hi, hi)) return 1;
  }
  // claim order of the single-phase sweep: tiles by their topological time (longest chain of upslope
  // tiles). Within a partition ascending tile index is a topological order by construction; across
  // partitions it is one when the partitions are numbered along the flow.
  {
    const int nt = c->n_tiles;
    std::vector<std::vector<int32_t>> dn(nt);
    std::vector<int32_t> indeg(nt, 0), tau(nt, 0), tiles;
    for (int i = 0; i < n; i++) {
      const int ti = c->dev_of_sweep[i] / kTile;
      for (int j : early[i]) {
        const int tj = c->dev_of_sweep[j] / kTile;
        if (tj == ti) { delete c; return fail("tribs_gpu_init: internal: a tile holds two cells with a flow relation"); }
        if (tj > ti) c->topo_order = false;
        dn[tj].push_back(ti);
      }
    }
    for (int t = 0; t < nt; t++) {
      std::sort(dn[t].begin(), dn[t].end());
      dn[t].erase(std::unique(dn[t].begin(), dn[t].end()), dn[t].end());
      for (int v : dn[t]) indeg[v]++;
    }
    tiles.reserve(nt);
    for (int t = 0; t < nt; t++) if (!indeg[t]) tiles.push_back(t);
    for (size_t q = 0; q < tiles.size(); q++) {
      const int t = tiles[q];
      for (int v : dn[t]) {
        tau[v] = std::max(tau[v], tau[t] + 1);
        if (--indeg[v] == 0) tiles.push_back(v);
      }
    }
    if ((int)tiles.size() != nt) { delete c; return fail("tribs_gpu_init: internal: the tiles' flow relation has a cycle"); }
    std::iota(tiles.begin(), tiles.end(), 0);
    std::stable_sort(tiles.begin(), tiles.end(), [&](int a, int b) { return tau[a] < tau[b]; });
    if (dev_upload(c, &c->d_tile_order, tiles)) return 1;
  }
  auto D = [&](int sweep) { return sweep < 0 ? -1 : c->dev_of_sweep[sweep]; };
  c->last_sweep_dev = c->dev_of_sweep[n - 1];

  tick("tile order");
  // ---- static arrays in device order
  auto permute_d = [&](const double *src, double padv) {
    std::vector<double> v(n_pad, padv);
    if (src) for (int i = 0; i < n; i++) v[c->dev_of_sweep[i]] = src[i];
    return v;
  };
  std::vector<int32_t> bnd(n_pad, -1), fdest(n_pad, -1), snode(n_pad, -1);
  std::vector<double> cosv(n_pad, 1.0), sinv(n_pad, 0.0), cosgw(n_pad, 1.0), vel(n_pad, 0.0);
  for (int i = 0; i < n; i++) {
    int dI = c->dev_of_sweep[i];
    bnd[dI] = mesh->boundary[i];
    fdest[dI] = D(mesh->flow_dest[i]);
    snode[dI] = D(mesh->stream_node[i]);
    // host libm (the same one the reference uses) evaluates the slope trigonometry once
    double alpha = atan(mesh->flow_slope[i]);
    cosv[dI] = alpha > 0.0 ? fabs(cos(alpha)) : 1.0;
    sinv[dI] = alpha > 0.0 ? fabs(sin(alpha)) : 0.0;
    cosgw[dI] = cos(alpha);
    vel[dI] = mesh->flow_vedglen[i] * 1000.;
  }
#define UPI(field, vec) { int32_t *p_; if (dev_upload(c, &p_, vec)) { return 1; } c->sv.field = p_; }
#define UPD(field, vec) { double *p_; if (dev_upload(c, &p_, vec)) { return 1; } c->sv.field = p_; }
  {
    std::vector<int32_t> ones(n_pad, 0);
    for (int i = 0; i < n; i++) ones[c->dev_of_sweep[i]] = (c->n_ranks == 1 || own_of(i) == c->rank) ? 1 : 0;
    if (dev_upload(c, &c->d_owned, ones)) return 1;
    c->sv.owned = c->d_owned;
  }
  UPI(boundary, bnd) UPI(flow_dest, fdest) UPI(stream_node, snode)
  UPD(cosv, cosv) UPD(sinv, sinv) UPD(cos_gw, cosgw) UPD(vel_mm, vel)
  UPD(varea, permute_d(mesh->varea, 1.0)) UPD(z, permute_d(mesh->z, 0.0)) UPD(bedrock, permute_d(prm->bedrock, 1.0))
  UPD(hillpath, permute_d(mesh->hillpath, 0.0)) UPD(streampath, permute_d(mesh->streampath, 0.0))
  UPD(contr_area, permute_d(mesh->contr_area, 1.0))
  UPD(ks, permute_d(prm->Ks, 1.0)) UPD(ths, permute_d(prm->ThetaS, 0.5)) UPD(thr, permute_d(prm->ThetaR, 0.1))
  UPD(lam, permute_d(prm->PoreSize, 0.5)) UPD(psib, permute_d(prm->AirEBubPres, -100.0)) UPD(f, permute_d(prm->DecayF, 0.01))
  UPD(ar, permute_d(prm->SatAnRatio, 1.0)) UPD(uar, permute_d(prm->UnsatAnRatio, 1.0))

  tick("static arrays");
  // ---- neighbour CSR in device order
  {
    std::vector<int32_t> rp(n_pad + 1, 0), col, lpos;
    std::vector<double> len, vl, len_in, vl_in;
    col.reserve(mesh->n_edges); lpos.reserve(mesh->n_edges);
    int maxdeg = 0;
    for (int dI = 0; dI < n_pad; dI++) {
      int i = c->sweep_of_dev[dI];
      if (i >= 0) {
        int k0 = mesh->nbr_rowptr[i], k1 = mesh->nbr_rowptr[i + 1];
        maxdeg = std::max(maxdeg, k1 - k0);
        for (int k = k0; k < k1; k++) {
          col.push_back(D(mesh->nbr_col[k]));
          lpos.push_back(mesh->nbr_listpos ? mesh->nbr_listpos[k] : k);
          len.push_back(mesh->nbr_len[k]);
          vl.push_back(mesh->nbr_vedglen[k]);
          len_in.push_back(mesh->nbr_len_in ? mesh->nbr_len_in[k] : mesh->nbr_len[k]);
          vl_in.push_back(mesh->nbr_vedglen_in ? mesh->nbr_vedglen_in[k] : mesh->nbr_vedglen[k]);
        }
      }
      rp[dI + 1] = (int32_t)col.size();
    }
    c->max_degree = maxdeg;
    if (maxdeg > 64) return fail("tribs_gpu_init: Voronoi degree > 64 is not supported");
    UPI(nbr_rowptr, rp) UPI(nbr_col, col) UPI(nbr_listpos, lpos)
    UPD(nbr_len, len) UPD(nbr_vedglen, vl) UPD(nbr_len_in, len_in) UPD(nbr_vedglen_in, vl_in)
  }
  tick("neighbour CSR");
  // ---- upslope contributor CSRs (rows in device order, entries by sweep position)
  {
    std::vector<int32_t> rp(n_pad + 1, 0), col, lrp(n_pad + 1, 0), lcol;
    for (int dI = 0; dI < n_pad; dI++) {
      int i = c->sweep_of_dev[dI];
      if (i >= 0) {
        for (int j : early[i]) col.push_back(c->dev_of_sweep[j]);   // ascending j by construction
        for (int j : late[i]) lcol.push_back(c->dev_of_sweep[j]);
      }
      rp[dI + 1] = (int32_t)col.size();
      lrp[dI + 1] = (int32_t)lcol.size();
    }
    if (dev_upload(c, &c->d_up_rowptr, rp) || dev_upload(c, &c->d_up_col, col)) return 1;
    if (dev_upload(c, &c->d_late_rowptr, lrp) || dev_upload(c, &c->d_late_col, lcol)) return 1;
  }
  if (dev_upload(c, &c->d_dev_of_sweep, c->dev_of_sweep) || dev_upload(c, &c->d_sweep_of_dev, c->sweep_of_dev)) return 1;

  tick("contributor CSRs");
  // ---- routing structures: contributors of every stream cell sorted by (hillpath, sweep pos); one set of
  //      entry arrays per partition (a stream cell and its hillslope always share a partition), so a
  //      partition's sums group the same way on one GPU and on many
  {
    std::vector<int32_t> scells;  // sweep indices of stream cells, sweep order
    for (int i = 0; i < n; i++) if (mesh->boundary[i] == 3) scells.push_back(i);
    c->stream_cells_sweep = scells;
    c->n_stream = (int)scells.size();
    const int slots = c->n_stream + 1;  // last slot = the outlet node
    std::vector<int32_t> slot_of(n, -1);
    for (int k = 0; k < c->n_stream; k++) slot_of[scells[k]] = k;
    std::vector<std::vector<int32_t>> members(slots);
    std::vector<int32_t> slot_part(slots, -1);
    for (int k = 0; k < c->n_stream; k++) slot_part[k] = own_of(scells[k]);
    int last_hill = -1;
    for (int i = 0; i < n; i++) {
      if (mesh->boundary[i] != 0) continue;
      last_hill = i;
      int sn = mesh->stream_node[i];
      int slot = (sn >= 0 && slot_of[sn] >= 0) ? slot_of[sn] : c->n_stream;
      members[slot].push_back(i);
      if (slot_part[slot] < 0) slot_part[slot] = own_of(i);
      if (slot_part[slot] != own_of(i))
        return fail("tribs_gpu_init: a hillslope cell and its stream node lie in different partitions (partitions must follow reaches)");
    }
    if (slot_part[c->n_stream] < 0) slot_part[c->n_stream] = c->n_parts - 1;
    c->last_hill_dev = last_hill >= 0 ? c->dev_of_sweep[last_hill] : -1;
    std::vector<int32_t> sdev(slots, -1);
    for (int k = 0; k < c->n_stream; k++) sdev[k] = c->dev_of_sweep[scells[k]];
    double max_path = 0.0;
    for (int i = 0; i < n; i++) if (mesh->boundary[i] == 0) max_path = std::max(max_path, mesh->hillpath[i]);
    c->seg_parts.assign(c->n_parts, SegPart{});
    if (const char *e = getenv("TRIBS_SEG_LEVEL2_MIN")) c->seg_level2_min = atoi(e);   // tests: 0 forces the two-level path on small basins
    for (int r : c->my_parts) {
      std::vector<int32_t> seg_cell, seg_slot;
      std::vector<double> seg_path, seg_area;
      for (int k = 0; k < slots; k++) {
        if (slot_part[k] != r) continue;
        auto &mm = members[k];
        std::stable_sort(mm.begin(), mm.end(), [&](int a, int b) { return mesh->hillpath[a] < mesh->hillpath[b]; });
        if (k < c->n_stream) {  // the stream cell itself: zero travel time, same bin arithmetic
          seg_cell.push_back(sdev[k]); seg_slot.push_back(k); seg_path.push_back(0.0);
          seg_area.push_back(mesh->varea[scells[k]]);
        }
        for (int i : mm) {
          seg_cell.push_back(c->dev_of_sweep[i]); seg_slot.push_back(k); seg_path.push_back(mesh->hillpath[i]);
          seg_area.push_back(mesh->varea[i]);
        }
      }
      SegPart &sp = c->seg_parts[r];
      sp.n_seg = (int)seg_cell.size();
      sp.blocks = std::max(1, (sp.n_seg + kSegBlock - 1) / kSegBlock);
      sp.blocks2 = std::max(1, (2 * sp.blocks + kSegBlock - 1) / kSegBlock);
      if (dev_upload(c, &sp.cell, seg_cell) || dev_upload(c, &sp.slot, seg_slot) || dev_upload(c, &sp.path, seg_path) ||
          dev_upload(c, &sp.area, seg_area) || dev_alloc(c, &sp.bnd, (size_t)sp.blocks) || dev_alloc(c, &sp.bnd2, (size_t)sp.blocks2))
        return 1;
    }
    if (dev_upload(c, &c->d_sn_dev, sdev)) return 1;
    RouteConst &rc = c->rc;
    rc.velcoef = prm->velcoef; rc.velratio = prm->velratio; rc.flowexp = prm->flowexp; rc.baseflow = prm->baseflow;
    rc.kincoef = prm->kincoef; rc.dt_reff = prm->dtReff; rc.outlet_contr_area = prm->outlet_contr_area;
    c->gwstep_h = prm->gwstep;
    rc.tstep = prm->tstep; rc.output_interval = prm->output_interval; rc.res_limit = std::max(prm->res_limit, 1);
    // the slowest hillslope velocity bounds the furthest dtReff bin ever written
    double vmin = prm->velcoef / prm->velratio;
    if (prm->flowexp > 0.0) {
      double amax = prm->outlet_contr_area;
      for (int i = 0; i < n; i++) amax = std::max(amax, mesh->contr_area[i]);
      double q = prm->baseflow > 0 ? prm->baseflow : 0.1;
      vmin = std::min(vmin, prm->kincoef * pow(q / amax, prm->flowexp));
    }
    long ring = (long)ceil(max_path / vmin / 3600. / prm->dtReff) + 4;
    if (ring > kMaxRing) return fail("tribs_gpu_init: hillslope travel time needs more than 4096 dtReff bins");
    rc.ring = (int32_t)ring;
    if (dev_alloc(c, &c->d_ring, (size_t)slots * ring) || dev_alloc(c, &c->d_qeff_now, (size_t)slots)) return 1;
    CK(cudaMemset(c->d_ring, 0, (size_t)slots * ring * sizeof(double)));
    CK(cudaMemset(c->d_qeff_now, 0, (size_t)slots * sizeof(double)));
    if (dev_alloc(c, &c->d_results, (size_t)TRIBS_N_RESULTS * rc.res_limit)) return 1;
    CK(cudaMemset(c->d_results, 0, (size_t)TRIBS_N_RESULTS * rc.res_limit * sizeof(double)));
    fill_kernel<<<(rc.res_limit + 255) / 256, 256>>>(c->d_results + (size_t)TRIBS_R_rmin * rc.res_limit, 9999.99, rc.res_limit);  // tFlowResults.cpp:224
    CK(cudaGetLastError());
    // window of output bins one step can touch: longest basin travel time / output interval
    double max_basin_tt = 0.0;
    for (int i = 0; i < n; i++) {
      double a = mesh->ttime ? mesh->ttime[i] : 0.0;
      double b2 = mesh->hillpath[i] / vmin / 3600.;
      double c3 = (mesh->hillpath[i] / (prm->velcoef / prm->velratio) + mesh->streampath[i] / prm->velcoef) / 3600.0;
      max_basin_tt = std::max(max_basin_tt, std::max(a, std::max(b2, c3)));
    }
    c->route_window = std::max(4, (int)ceil(max_basin_tt / prm->output_interval) + 3 +
                                      (int)ceil(prm->tstep / prm->output_interval));
    size_t stride = 5 * (size_t)c->route_window + kMeans;
    if (kRouteWarps * stride * sizeof(double) > 200 * 1024)
      return fail("tribs_gpu_init: basin travel-time window too wide for shared memory");
    CK(cudaFuncSetAttribute(tile_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)(kRouteWarps * stride * sizeof(double))));
    // Output bins one 32-cell tile can touch in one step: its cells are of one (segment, depth) bucket, so
    // their travel times are close. With flowexp == 0 the travel time used for binning is static: the
    // hillslope-only time for hillslope cells (rewritten before it is used, tKinemat.cpp:958-970), the
    // initial one for stream cells.
    int K = c->route_window;
    if (!(prm->flowexp > 0.0)) {
      const double hv0 = prm->velcoef / prm->velratio;
      std::vector<double> lo(c->n_tiles, 1e300), hi(c->n_tiles, -1e300);
      for (int i = 0; i < n; i++) {
        const double tt = mesh->boundary[i] == 0 ? mesh->hillpath[i] / hv0 / 3600. : (mesh->ttime ? mesh->ttime[i] : 0.0);
        const int t = c->dev_of_sweep[i] / kTile;
        lo[t] = std::min(lo[t], tt); hi[t] = std::max(hi[t], tt);
      }
      K = 1;
      for (int t = 0; t < c->n_tiles; t++)
        if (hi[t] >= lo[t]) K = std::max(K, (int)floor((hi[t] - lo[t] + 0.98 * prm->tstep) / prm->output_interval) + 2);
      K = std::min(K, c->route_window);
    }
    c->tile_bins = K;
    c->rec_stride = kRecBins + 5 * K;
  }
  tick("routing structures");
  c->mc.int_storm_max = prm->IntStormMAX; c->mc.surface_depth = prm->surfaceSoilDepth; c->mc.root_depth = prm->rootZoneDepth;
  c->mc.bas_area = prm->BasArea; c->mc.bas_area_flow = prm->BasArea_flow;
  c->mc.et_option = prm->EToption; c->mc.i_option = prm->Ioption; c->mc.sn_opt = prm->SnOpt; c->mc.rdstr_option = prm->RdstrOption;

  // ---- memory the peers address (nranks > 1): one allocation, carved identically on every rank
  if (c->n_ranks > 1) {
    // (sizes as in ensure_pipe: up to 3 phases per step - ET, sweep, saturated zone)
    const size_t S = (size_t)part->max_batch_steps, P3 = 3 * S, nt = (size_t)c->n_tiles;
    const size_t stride = 5 * (size_t)c->route_window + kMeans;
    size_t bytes = ((size_t)TRIBS_N_FIELDS + 4) * ((size_t)n_pad * 8 + 256)      // fields, qpub x2, wt_elev, transm
                   + (1 + kQueueClasses) * P3 * nt * 4 + 1024                      // task counters + the ready queues
                   + (size_t)c->n_parts * (S * stride + P3 * 4) * 8 + S * (size_t)(c->n_stream + 1) * 8 + (4u << 20);
    void *q = nullptr;
    CK(cudaMalloc(&q, bytes));
    CK(cudaMemset(q, 0, bytes));
    c->allocs.push_back(q);
    c->slab.base = (char *)q; c->slab.cap = bytes; c->slab.used = 0;
  }
  // ---- dynamic fields
  for (int k = 0; k < TRIBS_N_FIELDS; k++) {
    double *p;
    if (shared_alloc(c, &p, (size_t)n_pad)) return 1;
    std::vector<double> h(n_pad, 0.0);
    const double *src = (init && init->f[k]) ? init->f[k] : nullptr;
    if (k == TRIBS_F_TTime && !src) src = mesh->ttime;
    if (src) for (int i = 0; i < n; i++) h[c->dev_of_sweep[i]] = src[i];
    CK(cudaMemcpy(p, h.data(), (size_t)n_pad * sizeof(double), cudaMemcpyHostToDevice));
    c->dv.f[k] = p;
  }
  if (dev_alloc(c, &c->d_pub, (size_t)n_pad) || dev_alloc(c, &c->d_tile_counter, 4) || dev_alloc(c, &c->d_abort, 4) ||
      dev_alloc(c, &c->d_flags, 4) || dev_alloc(c, &c->d_tile_sums, (size_t)4 * c->n_tiles) ||
      dev_alloc(c, &c->d_globals, (size_t)TRIBS_N_GLOBALS) || shared_alloc(c, &c->d_wt_elev, (size_t)n_pad) ||
      shared_alloc(c, &c->d_transm, (size_t)n_pad) || dev_alloc(c, &c->d_stage, (size_t)n_pad))
    return 1;
  {
    const int32_t one = 1;
    if (dev_alloc(c, &c->d_stepbins, 1) || dev_alloc(c, &c->d_one, 1)) return 1;
    CK(cudaMemcpy(c->d_one, &one, sizeof(one), cudaMemcpyHostToDevice));
  }
  {   // buffers of the single-step routing path and of the per-partition sums
    const size_t own_tiles = (size_t)(c->own_end - c->own_begin) / kTile, stride = 5 * (size_t)c->route_window + kMeans;
    if (dev_alloc(c, &c->d_tile_rec, own_tiles * c->rec_stride) || dev_alloc(c, &c->d_block_part, (size_t)kReduceBlocksMax * stride * c->n_parts) ||
        dev_alloc(c, &c->d_part_vec, stride * kMaxParts) || dev_alloc(c, &c->d_sums_part, (size_t)4 * kMaxParts))
      return 1;
  }
  CK(cudaMemset(c->d_pub, 0, (size_t)n_pad * sizeof(PubWord)));
  CK(cudaMemset(c->d_tile_counter, 0, 16));
  CK(cudaMemset(c->d_abort, 0, 16));
  CK(cudaMemset(c->d_flags, 0, 16));
  CK(cudaMemset(c->d_globals, 0, TRIBS_N_GLOBALS * sizeof(double)));
  CK(cudaMemset(c->d_tile_sums, 0, (size_t)4 * c->n_tiles * sizeof(double)));

  tick("dynamic fields");
  // ---- persistent grid of the sweep: as many co-resident blocks as the device holds
  {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, c->device));
    if (!prop.cooperativeLaunch) return fail("tribs_gpu_init: device lacks cooperative launch");
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, unsat_sweep_kernel, c->unsat_threads, 0));
    if (per_sm < 1) return fail("tribs_gpu_init: unsat_sweep_kernel does not fit on an SM");
    int warps_per_block = c->unsat_threads / 32;
    int needed = (c->n_tiles + warps_per_block - 1) / warps_per_block;
    c->unsat_blocks = std::max(1, std::min(per_sm * prop.multiProcessorCount, needed));
  }
  tick("sweep grid");
  // nranks > 1: the batch buffers the peers address are carved now, in the same order on every rank
  if (c->n_ranks > 1 && ensure_pipe(c, pipe_of(c), part->max_batch_steps)) return 1;
  CK(cudaDeviceSynchronize());
  tick("done");
  *out = c;
  return 0;
}

// ------------------------------------------------------------------------------------ stepping
static void tribs_free_et(tribs_ctx *c);
static int check_sweep_abort(tribs_ctx *c) {
  int32_t flags[2] = {0, 0}, two[2] = {0, 0};
  CK(cudaMemcpyAsync(two, c->d_flags, 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(&flags[0], c->d_abort, 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(&flags[1], c->d_flags, 4, cudaMemcpyDeviceToHost, c->stream));   // bit 0 gather overflow, bit 1 bin overflow
  CK(cudaStreamSynchronize(c->stream));
  flags[1] = (two[0] ? 1 : 0) | (two[1] ? 2 : 0);
  if (flags[0] == 2) return fail("a peer rank did not reach the batch barrier within the time limit");
  if (flags[0]) return fail("sweep aborted: an input never arrived within the time limit (a schedule bug, or a peer rank that stopped)");
  if (flags[1] & 1) return fail("saturated zone: more non-zero edge fluxes at a cell than the gather buffer holds");
  if (flags[1] & 2) return fail("routing: a tile touched more output bins in one step than planned at init (travel times changed?)");
  return 0;
}

static WaterBalanceArgs water_balance_args(tribs_ctx *c, double now);
static int upload_sweep_field(tribs_ctx *c, const double *host, int field) {
  CK(cudaMemcpyAsync(c->d_stage, host, (size_t)c->n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  scatter_from_sweep_kernel<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->d_stage, c->d_dev_of_sweep, c->dv.f[field], c->n);
  c->launches++;
  return 0;
}

static inline int reduce_blocks(int n_tiles) { return std::max(1, std::min(kReduceBlocksMax, (n_tiles + 511) / 512)); }
static inline int part_tiles(const tribs_ctx *c, int r) { return (c->part_end[r] - c->part_begin[r]) / kTile; }

// basin accumulators: per-tile partials [n_phases][...] of this handle's partitions -> d_part[r][n_phases][4],
// then (single GPU) the partitions in partition order. buf_tile0 = global index of the buffer's first tile.
static int reduce_sums_local(tribs_ctx *c, const double *tile_sums, size_t phase_stride, int buf_tile0, int n_phases,
                             double *d_part) {
  for (int r : c->my_parts) {
    sums_reduce_kernel<<<n_phases, 256, 0, c->stream>>>(tile_sums, phase_stride, c->part_begin[r] / kTile - buf_tile0,
                                                         part_tiles(c, r), d_part + (size_t)r * n_phases * 4);
    c->launches++;
  }
  return 0;
}

// hillslope routing of one step from `srf_of_cell` (indexable by device index) into the rings, then the latch
static int route_segments(tribs_ctx *c, const DynView &dv, const double *srf_of_cell, double now, double qstrm_outlet,
                          double *qeff_out) {
  cudaStream_t st = c->stream;
  for (int r : c->my_parts) {
    const SegPart &sp = c->seg_parts[r];
    if (sp.n_seg == 0) continue;
    route_seg_kernel<<<sp.blocks, kSegBlock, 0, st>>>(c->sv, dv, c->rc, sp.cell, sp.slot, sp.path, sp.area, c->d_sn_dev, sp.n_seg,
                                                      now, qstrm_outlet, srf_of_cell, c->d_ring, sp.bnd);
    if (sp.blocks > c->seg_level2_min) {    // many blocks: their boundary runs are reduced once more before the sequential stitch
      route_seg_level2_kernel<<<sp.blocks2, kSegBlock, 0, st>>>(c->rc, sp.bnd, sp.blocks, c->d_ring, sp.bnd2);
      route_seg_fixup_kernel<<<1, 1024, 0, st>>>(c->rc, sp.bnd2, sp.blocks2, c->d_ring);
      c->launches++;
    } else
      route_seg_fixup_kernel<<<1, 1024, 0, st>>>(c->rc, sp.bnd, sp.blocks, c->d_ring);
    c->launches += 2;
  }
  const int slots = c->n_stream + 1;
  route_latch_kernel<<<(slots + 255) / 256, 256, 0, st>>>(c->rc, slots, now, c->d_ring, qeff_out);
  c->launches++;
  return 0;
}

static inline StepBins step_bins(const tribs_ctx *c, double now) {
  const int cur_bin0 = (int)floor((0.0 - .01 * c->rc.tstep + now) / c->rc.output_interval);
  const int cur_bin1 = (int)floor((0.0 - .99 * c->rc.tstep + now) / c->rc.output_interval);
  StepBins b;
  b.cur_bin = (cur_bin0 == cur_bin1) ? cur_bin0 : -1;   // store_saturation only writes when init==end
  b.bin_base = (int)floor((now + .01 * c->rc.tstep) / c->rc.output_interval);
  return b;
}

extern "C" int tribs_gpu_step(tribs_handle_t c, uint32_t mask, const tribs_step_t *s) {
  if (!c || !s) return fail("tribs_gpu_step: null argument");
  if (c->n_ranks > 1 && (mask & ~0u) != 0)
    return fail("tribs_gpu_step: a peer-memory partition (nranks > 1) advances through tribs_gpu_run only");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const int n_pad = c->n_pad;
  if (s->rain_mode == TRIBS_RAIN_UNIFORM) {
    fill_kernel<<<(n_pad + 255) / 256, 256, 0, st>>>(c->dv.f[TRIBS_F_Rain], s->rain_uniform, n_pad);
    c->launches++;
  } else if (s->rain_mode == TRIBS_RAIN_PER_CELL) {
    if (!s->rain) return fail("tribs_gpu_step: rain_mode PER_CELL without a rain array");
    if (upload_sweep_field(c, s->rain, TRIBS_F_Rain)) return 1;
  } else if (s->rain_mode == TRIBS_RAIN_GAUGES) {
    if (!s->rain || !c->d_gauge_of_cell) return fail("tribs_gpu_step: rain_mode GAUGES needs tribs_gpu_set_rain_gauges and the gauge values");
    CK(cudaMemcpyAsync(c->d_gauge_vals, s->rain, (size_t)c->n_gauges * sizeof(double), cudaMemcpyHostToDevice, st));
    gauge_fill_kernel<<<(n_pad + 255) / 256, 256, 0, st>>>(c->dv.f[TRIBS_F_Rain], c->d_gauge_vals, c->d_gauge_of_cell, n_pad);
    c->launches++;
  }
  if (s->qstrm && c->rc.flowexp > 0.0)
    if (upload_sweep_field(c, s->qstrm, TRIBS_F_Qstrm)) return 1;

  StepConst t;
  t.dt = s->dt; t.dt_gw = s->dt_gw; t.now = s->current_time; t.elapsed = (double)s->elapsed_steps;
  t.hourly_reset = s->hourly_reset; t.psrf_last = 0.0;

  const bool sweep_local = mask & (TRIBS_PHASE_UNSAT | TRIBS_PHASE_UNSAT_LOCAL);
  const bool sweep_remote = mask & (TRIBS_PHASE_UNSAT | TRIBS_PHASE_UNSAT_REMOTE);
  if (sweep_local || sweep_remote) {
    PhaseTimer pt(c, 0);
    if (sweep_local) {
      c->epoch++;
      if (c->partitioned) CK(cudaMemsetAsync(c->d_tile_sums, 0, (size_t)4 * c->n_tiles * sizeof(double), st));
    }
    long long epoch = c->epoch;
    int last = c->last_sweep_dev;
    auto sweep = [&](int32_t *list, int count) -> int {
      if (count <= 0) return 0;
      CK(cudaMemsetAsync(c->d_tile_counter, 0, 4, st));
      void *args[] = {&c->sv, &c->dv, &c->mc, &t, &c->d_up_rowptr, &c->d_up_col, &c->d_pub, &epoch, &count,
                      &list, &c->d_tile_counter, &c->d_abort, &c->d_tile_sums, &c->d_globals, &last};
      int blocks = std::max(1, std::min(c->unsat_blocks, (count + 3) / 4));
      CK(cudaLaunchCooperativeKernel((void *)unsat_sweep_kernel, dim3(blocks), dim3(c->unsat_threads), args, 0, st));
      c->launches++;
      return 0;
    };
    if (!c->partitioned) {
      if (sweep(c->d_tile_order, c->n_tiles)) return 1;
    } else {
      // pass "local": tiles whose inputs all live on this rank; pass "remote": tiles downstream of
      // a cell another rank owns -- the host unpacks that rank's outflow between the two passes
      if (sweep_local && sweep(c->d_tiles_local, c->n_tiles_local)) return 1;
      if (sweep_remote && sweep(c->d_tiles_remote, c->n_tiles_remote)) return 1;
    }
    if (sweep_remote) {
      if (c->n_late > 0) {
        unsat_late_qpin_kernel<<<(n_pad + 255) / 256, 256, 0, st>>>(c->dv, c->d_late_rowptr, c->d_late_col, n_pad);
        c->launches++;
      }
      if (reduce_sums_local(c, c->d_tile_sums, 0, 0, 1, c->d_sums_part)) return 1;
      PartVecs pv{};
      pv.n = c->n_parts;
      for (int r = 0; r < c->n_parts; r++) pv.v[r] = c->d_sums_part + (size_t)r * 4;
      combine_sums_kernel<<<1, 32, 0, st>>>(pv, nullptr, 1, c->d_globals);
      c->launches++;
    }
    c->new_valid = true;
    c->at_boundary = false;
  }
  if (mask & TRIBS_PHASE_SAT) {
    PhaseTimer pt(c, 1);
    swap_old_new(c);  // ResetGW: Old <- New without moving a byte; sat rewrites every New value
    sat_prepare_kernel<<<(n_pad + 255) / 256, 256, 0, st>>>(c->sv, c->dv, c->d_wt_elev, c->d_transm, n_pad);
    int blocks = (n_pad + 127) / 128;
    if (c->max_degree <= 16)
      sat_cells_kernel<16><<<blocks, 128, 0, st>>>(c->sv, c->dv, c->mc, t, c->d_wt_elev, c->d_transm, n_pad, c->d_tile_sums, c->d_flags, c->d_globals);
    else if (c->max_degree <= 32)
      sat_cells_kernel<32><<<blocks, 128, 0, st>>>(c->sv, c->dv, c->mc, t, c->d_wt_elev, c->d_transm, n_pad, c->d_tile_sums, c->d_flags, c->d_globals);
    else
      sat_cells_kernel<64><<<blocks, 128, 0, st>>>(c->sv, c->dv, c->mc, t, c->d_wt_elev, c->d_transm, n_pad, c->d_tile_sums, c->d_flags, c->d_globals);
    if (reduce_sums_local(c, c->d_tile_sums, 0, 0, 1, c->d_sums_part)) return 1;
    PartVecs pv{};
    pv.n = c->n_parts;
    for (int r = 0; r < c->n_parts; r++) pv.v[r] = c->d_sums_part + (size_t)r * 4;
    combine_sums_kernel<<<1, 32, 0, st>>>(pv, c->d_one, 1, c->d_globals);   // a SAT phase: no TotRain term
    c->launches += 3;
    c->gw_dirty = true;
    c->new_valid = true;
    c->at_boundary = false;
  }
  if (mask & TRIBS_PHASE_ROUTE) {
    PhaseTimer pt(c, 2);
    const double now = s->current_time;
    if (route_segments(c, c->dv, c->dv.f[TRIBS_F_srf], now, s->qstrm_outlet, c->d_qeff_now)) return 1;
    const int window = c->route_window, K = c->tile_bins;
    const size_t stride = 5 * (size_t)window + kMeans;
    const StepBins sb = step_bins(c, now);
    CK(cudaMemcpyAsync(c->d_stepbins, &sb, sizeof(sb), cudaMemcpyHostToDevice, st));
    WaterBalanceArgs wb = water_balance_args(c, now);
    RouteStep rs{now, s->qstrm_outlet, 0.0, wb.gw_time, 0};
    const int own_tiles = (c->own_end - c->own_begin) / kTile;
    route_tiles_kernel<<<(own_tiles * 32 + 127) / 128, 128, 0, st>>>(c->sv, c->dv, c->rc, c->mc, wb, rs, c->own_begin, own_tiles,
                                                                      c->last_hill_dev, K, c->d_tile_rec, c->d_globals, c->d_flags);
    c->launches++;
    PartVecs pv{};
    pv.n = c->n_parts;
    for (int r = 0; r < c->n_parts; r++) {
      const int nt = part_tiles(c, r), G = reduce_blocks(nt);
      double *bp = c->d_block_part + (size_t)r * kReduceBlocksMax * stride, *vec = c->d_part_vec + (size_t)r * stride;
      tile_reduce_kernel<<<dim3(G, 1), kRouteWarps * 32, kRouteWarps * stride * sizeof(double), st>>>(
          c->d_tile_rec, 0, (c->part_begin[r] - c->own_begin) / kTile, nt, K, window, c->d_stepbins, bp, c->d_flags);
      part_reduce_kernel<<<((int)stride * 32 + 255) / 256, 256, 0, st>>>(bp, G, window, 1, vec);
      pv.v[r] = vec;
      c->launches += 2;
    }
    combine_results_kernel<<<(5 * window + kMeans + 255) / 256, 256, 0, st>>>(pv, c->d_stepbins, 1, window, c->rc.res_limit,
                                                                              sb.bin_base, window, c->d_results, c->d_globals);
    c->launches++;
  }
  if (mask & TRIBS_PHASE_RESET) {
    PhaseTimer pt(c, 3);
    swap_old_new(c);  // Old <- New
    const int zero5[5] = {TRIBS_F_Qpin, TRIBS_F_srf, TRIBS_F_sbsrf, TRIBS_F_hsrf, TRIBS_F_psrf};
    for (int k = 0; k < 5; k++) CK(cudaMemsetAsync(c->dv.f[zero5[k]], 0, (size_t)n_pad * sizeof(double), st));
    if (c->gw_dirty) {
      CK(cudaMemsetAsync(c->dv.f[TRIBS_F_satsrf], 0, (size_t)n_pad * sizeof(double), st));
      CK(cudaMemsetAsync(c->dv.f[TRIBS_F_gwchange], 0, (size_t)n_pad * sizeof(double), st));
      c->gw_dirty = false;
    }
    c->new_valid = false;
    c->at_boundary = true;
  }
  CK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------ data movement
static double *field_ptr(tribs_ctx *c, int k) {
  // after Reset() the reference's New values equal the Old ones; here only Old holds them
  if (!c->new_valid && k >= TRIBS_F_NwtNew && k <= TRIBS_F_RiNew) return c->dv.f[k - TRIBS_F_NwtNew + TRIBS_F_NwtOld];
  return c->dv.f[k];
}

extern "C" int tribs_gpu_sync(tribs_handle_t c, uint64_t field_mask, tribs_state_t *host) {
  if (!c || !host) return fail("tribs_gpu_sync: null argument");
  CK(cudaSetDevice(c->device));
  if (check_sweep_abort(c)) return 1;
  for (int k = 0; k < TRIBS_N_FIELDS; k++) {
    if (!(field_mask & (1ull << k)) || !host->f[k]) continue;
    gather_to_sweep_kernel<<<(c->n + 255) / 256, 256, 0, c->stream>>>(field_ptr(c, k), c->d_dev_of_sweep, c->d_stage, c->n);
    c->launches++;
    CK(cudaMemcpyAsync(host->f[k], c->d_stage, (size_t)c->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

extern "C" int tribs_gpu_push(tribs_handle_t c, uint64_t field_mask, const tribs_state_t *host) {
  if (!c || !host) return fail("tribs_gpu_push: null argument");
  CK(cudaSetDevice(c->device));
  for (int k = 0; k < TRIBS_N_FIELDS; k++) {
    if (!(field_mask & (1ull << k)) || !host->f[k]) continue;
    CK(cudaMemcpyAsync(c->d_stage, host->f[k], (size_t)c->n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    scatter_from_sweep_kernel<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->d_stage, c->d_dev_of_sweep, field_ptr(c, k), c->n);
    c->launches++;
    CK(cudaStreamSynchronize(c->stream));
  }
  return 0;
}


// ====================================================================================
// Kernel (1): ET + interception (tribs_gpu_et_*). See et_update.cuh for the cell update.
// last-writer scan of the tSnowPack members along the sweep, see snow_probe_kernel
struct SnowCarry {
  int32_t *key_alb[2] = {nullptr, nullptr}, *key_can[2] = {nullptr, nullptr};   // [n + 1], ping-pong between probe rounds
  double *val_alb[2] = {nullptr, nullptr}, *val_can[2] = {nullptr, nullptr};    // [n]
  int32_t *last_alb = nullptr, *last_can = nullptr;                             // [n + 1] scanned keys
  uint8_t *sensitive = nullptr;                                                 // [n_pad] device order
  double *members = nullptr;                                                    // [2] albedo, snCanWE between calls
  int32_t *flags = nullptr;                                                     // [2] changed, n_sensitive
};
struct MaxKey {
  __host__ __device__ int32_t operator()(int32_t a, int32_t b) const { return a > b ? a : b; }
};

struct EtState {
  bool ready = false;          // tribs_gpu_et_init completed
  EtConst k{};
  EtStatic st{};
  EtView ev{};
  double *d_met[2] = {nullptr, nullptr};        // [9][cap] packed records of the two met slots
  int32_t *d_record[2] = {nullptr, nullptr};    // cell -> record, device order
  bool have_record[2] = {false, false};
  int met_cap = 0;
  // pinned staging ring for the records: a buffer is reused only after its copy has completed
  static constexpr int kRing = 4;
  double *h_pin[2][kRing] = {};
  cudaEvent_t pin_ev[2][kRing] = {};
  int pin_cap = 0, gen = 0;
  std::vector<int32_t> h_rec;
  // snow pack (tribs_gpu_snow_init)
  int snow_option = 0;
  bool snow_ready = false;
  SnowConst sc{};
  SnowView snv{};
  // carry-over of the tSnowPack members `albedo` and `snCanWE` along the sweep (canopy runs only)
  SnowCarry carry{};
  void *d_scan_tmp = nullptr;
  size_t scan_tmp_bytes = 0;
  int32_t *h_flags = nullptr;                   // pinned: [0] any answer changed, [1] cells that depend on the incoming albedo
  int last_probe_rounds = 0;
  double members[2] = {0.88, 0.0};              // tSnowPack constructor values (tSnowPack.cpp:103-150)
};

static void tribs_free_et(tribs_ctx *c) {
  if (!c->et) return;
  for (int sl = 0; sl < 2; sl++)
    for (int g = 0; g < EtState::kRing; g++) {
      if (c->et->h_pin[sl][g]) cudaFreeHost(c->et->h_pin[sl][g]);
      if (c->et->pin_ev[sl][g]) cudaEventDestroy(c->et->pin_ev[sl][g]);
    }
  if (c->et->h_flags) cudaFreeHost(c->et->h_flags);
  delete c->et;
  c->et = nullptr;
}

static WaterBalanceArgs water_balance_args(tribs_ctx *c, double now) {
  WaterBalanceArgs wb{};
  wb.uns_storage = c->dv.f[TRIBS_F_UnSaturatedStorage];
  wb.sat_storage = c->dv.f[TRIBS_F_SaturatedStorage];
  wb.evapotrans = c->et ? c->et->ev.f[TRIBS_ET_EvapoTrans] : nullptr;
  wb.gwstep_h = c->gwstep_h;
  wb.gw_time = fmod(now, c->gwstep_h) == 0.0;          // Simulator::GW_label, tSimul.cpp:511
  return wb;
}

__global__ void __launch_bounds__(128)
et_cells_kernel(EtConst k, EtStatic st, EtStepDev s, EtView ev, DynView dv, int n_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad || st.boundary[i] < 0) return;
  et_cell(k, st, s, ev, dv, i);
}

template <class T>
static std::vector<T> to_device_order(tribs_ctx *c, const T *sweep, T pad) {
  std::vector<T> v((size_t)c->n_pad, pad);
  for (int d = 0; d < c->n_pad; d++) {
    const int i = c->sweep_of_dev[d];
    if (i >= 0) v[d] = sweep[i];
  }
  return v;
}

static int et_init_impl(tribs_handle_t c, const tribs_et_params_t *p);
extern "C" int tribs_gpu_et_init(tribs_handle_t c, const tribs_et_params_t *p) {
  const int rc = et_init_impl(c, p);
  if (rc && c && c->et && !c->et->ready) tribs_free_et(c);   // a failed upload must not block a retry with "already initialised"
  return rc;
}
static int et_init_impl(tribs_handle_t c, const tribs_et_params_t *p) {
  if (!c || !p) return fail("tribs_gpu_et_init: null argument");
  if (c->et) return fail("tribs_gpu_et_init: already initialised");
  if (!p->flow_slope || !p->aspect || !p->vol_heat_cond || !p->soil_heat_cap || !p->soil_cutoff || !p->root_cutoff ||
      !p->land_class || !p->land_table || p->n_land_classes <= 0)
    return fail("tribs_gpu_et_init: missing per-cell arrays or land table");
  if (p->et_option < 0 || p->et_option > 3) return fail("tribs_gpu_et_init: OPTEVAPOTRANS must be 0..3 (pan evaporation is not covered)");
  if (p->i_option < 0 || p->i_option > 2) return fail("tribs_gpu_et_init: OPTINTERCEPT must be 0..2");
  if (p->et_option && p->gflux_option != 1 && p->gflux_option != 2) return fail("tribs_gpu_et_init: GFLUXOPTION must be 1 or 2");
  if (p->shelter_option >= 1 && p->shelter_option <= 3 &&
      (!p->hor_angle_0000 || !p->init || !p->init[TRIBS_ET_SheltFact]))
    return fail("tribs_gpu_et_init: OPTRADSHELT 1-3 needs tShelter's sky-view factor (init[SheltFact]) and hor_angle_0000");
  if (p->lu_option != 0) return fail("tribs_gpu_et_init: OPTLANDUSE 1 (dynamic land-use grids) is not covered");
  if (p->snow_option != 0 && p->et_option == 0) return fail("tribs_gpu_et_init: OPTSNOW 1 needs the evaporation scheme");
  if (p->et_option == 0 && p->i_option == 2) return fail("tribs_gpu_et_init: Rutter interception needs the evaporation scheme (tSimul.cpp:444-450)");
  CK(cudaSetDevice(c->device));
  const int n = c->n;
  for (int i = 0; i < n; i++) {
    if (p->land_class[i] < 0 || p->land_class[i] >= p->n_land_classes) return fail("tribs_gpu_et_init: land class out of range");
    if (p->et_option && p->gflux_option == 2) {   // ForceRestore's validity check (tEvapoTrans.cpp:2683-2691)
      const double w1 = 2. * (4.0 * atan(1.0)) / 86400.;
      const double nd = 0.1 / pow((2.0 * (p->vol_heat_cond[i] / p->soil_heat_cap[i]) / w1), 0.5);
      if (!(nd >= 0.0 && nd <= 5.0)) return fail("tribs_gpu_et_init: normalised depth of the force-restore equation out of 0..5");
    }
  }
  EtState *e = new EtState();
  c->et = e;
  e->k.et_option = p->et_option; e->k.i_option = p->i_option; e->k.gflux_option = p->gflux_option;
  e->k.shelter_option = p->shelter_option;
  e->snow_option = p->snow_option;
  e->k.metstep_min = p->metstep_min; e->k.etistep_h = p->etistep_h; e->k.tlinke = p->tlinke;
  e->k.temp_lapse = p->temp_lapse; e->k.max_interstorm = p->max_interstorm;
  et_derived_constants(e->k);
  std::vector<LandClass> cls((size_t)p->n_land_classes);
  for (int q = 0; q < p->n_land_classes; q++) cls[q] = make_land_class(p->land_table + (size_t)q * 15, p->et_option);
  LandClass *d_cls = nullptr;
  if (dev_upload(c, &d_cls, cls)) return 1;
  e->st.classes = d_cls;
  std::vector<double> cs(n), sn(n);
  for (int i = 0; i < n; i++) {
    const double slope = fabs(atan(p->flow_slope[i]));
    cs[i] = cos(slope); sn[i] = sin(slope);
  }
  auto up = [&](const double *sweep, const double **dst) {
    double *d = nullptr;
    if (dev_upload(c, &d, to_device_order<double>(c, sweep, 0.0))) return 1;
    *dst = d;
    return 0;
  };
  if (up(cs.data(), &e->st.cos_slope) || up(sn.data(), &e->st.sin_slope) || up(p->aspect, &e->st.aspect) ||
      up(p->vol_heat_cond, &e->st.heat_cond) || up(p->soil_heat_cap, &e->st.heat_cap) ||
      up(p->soil_cutoff, &e->st.soil_cutoff) || up(p->root_cutoff, &e->st.root_cutoff)) return 1;
  e->st.hor_angle_0000 = nullptr;
  if (p->shelter_option >= 1 && p->shelter_option <= 3 && up(p->hor_angle_0000, &e->st.hor_angle_0000)) return 1;
  int32_t *d_lc = nullptr;
  if (dev_upload(c, &d_lc, to_device_order<int32_t>(c, p->land_class, 0))) return 1;
  e->st.land_class = d_lc;
  e->st.z = c->sv.z; e->st.thr = c->sv.thr; e->st.bedrock = c->sv.bedrock; e->st.varea = c->sv.varea;
  e->st.boundary = c->sv.boundary; e->st.sweep_of_dev = c->d_sweep_of_dev;
  for (int f = 0; f < TRIBS_N_ET_FIELDS; f++) {
    const double *src = (p->init && p->init[f]) ? p->init[f] : nullptr;
    double *d = nullptr;
    if (src) { if (dev_upload(c, &d, to_device_order<double>(c, src, 0.0))) return 1; }
    else {
      if (dev_alloc(c, &d, (size_t)c->n_pad)) return 1;
      CK(cudaMemset(d, 0, (size_t)c->n_pad * sizeof(double)));
    }
    e->ev.f[f] = d;
  }
  for (int sl = 0; sl < 2; sl++) if (dev_alloc(c, &e->d_record[sl], (size_t)c->n_pad)) return 1;
  e->ready = true;
  return 0;
}

// copies one tribs_met_t into met slot `sl`; the records travel as one packed block
static int et_upload_met(tribs_ctx *c, int sl, const tribs_met_t *m, MetDev *out) {
  EtState *e = c->et;
  if (!m || m->n_records <= 0) return fail("tribs_gpu_et_step: met record missing");
  const double *src[9] = {m->air_temp, m->dew_temp, m->rel_humid, m->vap_press, m->atm_press,
                          m->wind_speed, m->sky_cover, m->rad_global, m->elev};
  for (int q = 0; q < 9; q++) if (!src[q]) return fail("tribs_gpu_et_step: met series missing");
  const int nr = m->n_records;
  if (nr > e->met_cap) {
    for (int q = 0; q < 2; q++) if (dev_alloc(c, &e->d_met[q], (size_t)9 * nr)) return 1;
    e->met_cap = nr;
  }
  if (nr > e->pin_cap) {
    CK(cudaStreamSynchronize(c->stream));
    for (int a = 0; a < 2; a++)
      for (int g = 0; g < EtState::kRing; g++) {
        if (e->h_pin[a][g]) CK(cudaFreeHost(e->h_pin[a][g]));
        CK(cudaHostAlloc((void **)&e->h_pin[a][g], (size_t)9 * nr * sizeof(double), cudaHostAllocDefault));
        if (!e->pin_ev[a][g]) CK(cudaEventCreateWithFlags(&e->pin_ev[a][g], cudaEventDisableTiming));
      }
    e->pin_cap = nr;
  }
  const int g = e->gen % EtState::kRing;
  CK(cudaEventSynchronize(e->pin_ev[sl][g]));
  double *hp = e->h_pin[sl][g];
  for (int q = 0; q < 9; q++) memcpy(hp + (size_t)q * nr, src[q], (size_t)nr * sizeof(double));
  CK(cudaMemcpyAsync(e->d_met[sl], hp, (size_t)9 * nr * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CK(cudaEventRecord(e->pin_ev[sl][g], c->stream));
  if (m->cell_record) {
    for (int i = 0; i < c->n; i++)
      if (m->cell_record[i] < 0 || m->cell_record[i] >= nr) return fail("tribs_gpu_et_step: cell_record out of range");
    e->h_rec = to_device_order<int32_t>(c, m->cell_record, 0);
    CK(cudaMemcpyAsync(e->d_record[sl], e->h_rec.data(), e->h_rec.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));        // h_rec is reused by the next call
    e->have_record[sl] = true;
  } else if (!e->have_record[sl]) return fail("tribs_gpu_et_step: cell_record not given yet");
  double *b = e->d_met[sl];
  out->cell_record = e->d_record[sl];
  out->air_temp = b; out->dew_temp = b + nr; out->rel_humid = b + 2 * (size_t)nr; out->vap_press = b + 3 * (size_t)nr;
  out->atm_press = b + 4 * (size_t)nr; out->wind_speed = b + 5 * (size_t)nr; out->sky_cover = b + 6 * (size_t)nr;
  out->rad_global = b + 7 * (size_t)nr; out->elev = b + 8 * (size_t)nr;
  return 0;
}

__global__ void __launch_bounds__(128)
snow_cells_kernel(EtConst k, EtStatic st, EtStepDev s, SnowConst sc, EtView ev, SnowView sv, DynView dv, int n_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad || st.boundary[i] < 0) return;
  SnowProbe pr;
  snow_cell_step(k, st, s, sc, ev, sv, dv, i, 0.88, 0.0, pr);   // no canopy: nothing reads the carried members
}

// ---- canopy runs: last-writer scan of the tSnowPack members along the sweep (snow_update.cuh) ----
// Per sweep position p: key = p + 1 if the cell at p writes the member, else 0; an exclusive
// max-scan of the keys gives every position the last writer before it.
__device__ __forceinline__ double carried_in(const int32_t *last, const double *val, const double *member, int pos) {
  const int32_t q = last[pos];
  return q > 0 ? val[q - 1] : *member;
}

// round 0: every cell answers with the member values of the previous call; later rounds: the cells
// whose answer can depend on the incoming albedo answer again with the scanned values of the
// previous round (buffer `rd`), everybody else repeats its answer
__global__ void __launch_bounds__(128)
snow_probe_kernel(EtConst k, EtStatic st, EtStepDev s, SnowConst sc, EtView ev, SnowView sv, DynView dv, SnowCarry cy,
                  int round, int rd, int wr, int n_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad || st.boundary[i] < 0) return;
  const int pos = st.sweep_of_dev[i];
  if (round > 0 && !cy.sensitive[i]) {
    cy.key_alb[wr][pos] = cy.key_alb[rd][pos]; cy.val_alb[wr][pos] = cy.val_alb[rd][pos];
    cy.key_can[wr][pos] = cy.key_can[rd][pos]; cy.val_can[wr][pos] = cy.val_can[rd][pos];
    return;
  }
  const double alb_in = (round == 0) ? cy.members[0] : carried_in(cy.last_alb, cy.val_alb[rd], cy.members, pos);
  SnowProbe pr;
  snow_cell_t<false>(k, st, s, sc, ev, sv, dv, i, alb_in, 0.0, pr);
  const int32_t ka = pr.alb_writer ? pos + 1 : 0, kc = pr.can_writer ? pos + 1 : 0;
  if (round == 0) {
    cy.sensitive[i] = (uint8_t)pr.sensitive;
    if (pr.sensitive) atomicAdd(&cy.flags[1], 1);
  } else if (ka != cy.key_alb[rd][pos] || kc != cy.key_can[rd][pos] ||
             __double_as_longlong(pr.alb_value) != __double_as_longlong(cy.val_alb[rd][pos]) ||
             __double_as_longlong(pr.can_value) != __double_as_longlong(cy.val_can[rd][pos])) {
    cy.flags[0] = 1;
  }
  cy.key_alb[wr][pos] = ka; cy.val_alb[wr][pos] = pr.alb_value;
  cy.key_can[wr][pos] = kc; cy.val_can[wr][pos] = pr.can_value;
}

__global__ void __launch_bounds__(128)
snow_commit_kernel(EtConst k, EtStatic st, EtStepDev s, SnowConst sc, EtView ev, SnowView sv, DynView dv, SnowCarry cy,
                   int buf, int n_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad || st.boundary[i] < 0) return;
  const int pos = st.sweep_of_dev[i];
  const double alb_in = carried_in(cy.last_alb, cy.val_alb[buf], cy.members, pos);
  const double can_in = carried_in(cy.last_can, cy.val_can[buf], cy.members + 1, pos);
  SnowProbe pr;
  snow_cell_step(k, st, s, sc, ev, sv, dv, i, alb_in, can_in, pr);
}

// what the last cell of the sweep leaves in the members for the next call
__global__ void snow_members_kernel(SnowCarry cy, int buf, int n) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const int32_t qa = cy.last_alb[n], qc = cy.last_can[n];
    if (qa > 0) cy.members[0] = cy.val_alb[buf][qa - 1];
    if (qc > 0) cy.members[1] = cy.val_can[buf][qc - 1];
  }
}

extern "C" int tribs_gpu_snow_init(tribs_handle_t c, const tribs_snow_params_t *p) {
  if (!c || !p) return fail("tribs_gpu_snow_init: null argument");
  if (!c->et || !c->et->snow_option) return fail("tribs_gpu_snow_init: call tribs_gpu_et_init with snow_option = 1 first");
  if (c->et->snow_ready) return fail("tribs_gpu_snow_init: already initialised");
  if (p->hill_albedo_option < 0 || p->hill_albedo_option > 2) return fail("tribs_gpu_snow_init: HILLALBOPT must be 0..2");
  CK(cudaSetDevice(c->device));
  EtState *e = c->et;
  e->sc.min_sn_temp = p->min_sn_temp; e->sc.snliqfrac = p->snliqfrac; e->sc.hill_albedo_option = p->hill_albedo_option;
  for (int f = 0; f < TRIBS_N_SNOW_FIELDS; f++) {
    const double *src = (p->init && p->init[f]) ? p->init[f] : nullptr;
    double *d = nullptr;
    if (src) { if (dev_upload(c, &d, to_device_order<double>(c, src, 0.0))) return 1; }
    else {
      if (dev_alloc(c, &d, (size_t)c->n_pad)) return 1;
      CK(cudaMemset(d, 0, (size_t)c->n_pad * sizeof(double)));
    }
    e->snv.f[f] = d;
  }
  if (e->k.i_option != 0) {     // canopy: buffers of the last-writer scan (snow_probe_kernel)
    SnowCarry &cy = e->carry;
    const size_t n1 = (size_t)c->n + 1;
    for (int b = 0; b < 2; b++) {
      if (dev_alloc(c, &cy.key_alb[b], n1) || dev_alloc(c, &cy.key_can[b], n1) ||
          dev_alloc(c, &cy.val_alb[b], (size_t)c->n) || dev_alloc(c, &cy.val_can[b], (size_t)c->n)) return 1;
      CK(cudaMemset(cy.key_alb[b], 0, n1 * sizeof(int32_t)));
      CK(cudaMemset(cy.key_can[b], 0, n1 * sizeof(int32_t)));
      CK(cudaMemset(cy.val_alb[b], 0, (size_t)c->n * sizeof(double)));
      CK(cudaMemset(cy.val_can[b], 0, (size_t)c->n * sizeof(double)));
    }
    if (dev_alloc(c, &cy.last_alb, n1) || dev_alloc(c, &cy.last_can, n1) || dev_alloc(c, &cy.sensitive, (size_t)c->n_pad) ||
        dev_alloc(c, &cy.members, (size_t)2) || dev_alloc(c, &cy.flags, (size_t)2)) return 1;
    CK(cudaMemset(cy.last_alb, 0, n1 * sizeof(int32_t)));
    CK(cudaMemset(cy.last_can, 0, n1 * sizeof(int32_t)));
    CK(cudaMemset(cy.sensitive, 0, (size_t)c->n_pad));
    CK(cudaMemcpy(cy.members, e->members, 2 * sizeof(double), cudaMemcpyHostToDevice));
    size_t bytes = 0;
    CK(cub::DeviceScan::ExclusiveScan(nullptr, bytes, cy.key_alb[0], cy.last_alb, MaxKey(), (int32_t)0, (int)n1, c->stream));
    uint8_t *tmp = nullptr;
    if (dev_alloc(c, &tmp, bytes ? bytes : 1)) return 1;
    e->d_scan_tmp = tmp; e->scan_tmp_bytes = bytes;
    CK(cudaHostAlloc((void **)&e->h_flags, 2 * sizeof(int32_t), cudaHostAllocDefault));
  }
  e->snow_ready = true;
  return 0;
}

// One callSnowPack over a basin with canopy: probe rounds until the carried members are settled, then commit.
static int snow_step_with_canopy(tribs_ctx *c, const EtStepDev &d, const SnowConst &sc) {
  EtState *e = c->et;
  SnowCarry &cy = e->carry;
  const int n1 = c->n + 1, grid = (c->n_pad + 127) / 128;
  auto scan = [&](const int32_t *key, int32_t *last) {
    size_t bytes = e->scan_tmp_bytes;
    return cub::DeviceScan::ExclusiveScan(e->d_scan_tmp, bytes, key, last, MaxKey(), (int32_t)0, n1, c->stream);
  };
  auto read_flags = [&]() {
    CK(cudaMemcpyAsync(e->h_flags, cy.flags, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
  };
  CK(cudaMemsetAsync(cy.flags, 0, 2 * sizeof(int32_t), c->stream));
  snow_probe_kernel<<<grid, 128, 0, c->stream>>>(e->k, e->st, d, sc, e->ev, e->snv, c->dv, cy, 0, 1, 0, c->n_pad);
  c->launches++;
  CK(scan(cy.key_alb[0], cy.last_alb));
  if (read_flags()) return 1;
  int buf = 0, rounds = 1;
  if (e->h_flags[1] > 0) {
    for (;;) {
      if (rounds > c->n + 1) return fail("tribs_gpu_snow_step: the albedo carry-over did not settle");
      const int rd = buf, wr = buf ^ 1;
      CK(cudaMemsetAsync(cy.flags, 0, sizeof(int32_t), c->stream));
      snow_probe_kernel<<<grid, 128, 0, c->stream>>>(e->k, e->st, d, sc, e->ev, e->snv, c->dv, cy, rounds, rd, wr, c->n_pad);
      c->launches++;
      rounds++;
      if (read_flags()) return 1;
      buf = wr;
      if (!e->h_flags[0]) break;            // same answers as the round before: last_alb already belongs to them
      CK(scan(cy.key_alb[buf], cy.last_alb));
    }
  }
  e->last_probe_rounds = rounds;
  CK(scan(cy.key_can[buf], cy.last_can));
  snow_commit_kernel<<<grid, 128, 0, c->stream>>>(e->k, e->st, d, sc, e->ev, e->snv, c->dv, cy, buf, c->n_pad);
  snow_members_kernel<<<1, 32, 0, c->stream>>>(cy, buf, c->n);
  c->launches += 2;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int tribs_gpu_snow_step(tribs_handle_t c, const tribs_et_step_t *s, int32_t hourly_time_step) {
  if (!c || !s) return fail("tribs_gpu_snow_step: null argument");
  if (!c->et || !c->et->snow_ready) return fail("tribs_gpu_snow_step: call tribs_gpu_snow_init first");
  if (s->hour < 0 || s->hour > 23) return fail("tribs_gpu_snow_step: hour out of range");
  CK(cudaSetDevice(c->device));
  EtState *e = c->et;
  static const double rsRatio[24] = {3.837, 3.589, 3.21, 2.43, 1.617, 1.196, 1.067, 1.014, 0.995, 0.976, 0.976, 1.0,
                                     1.053, 1.167, 1.354, 1.637, 2.043, 2.66, 3.215, 3.507, 3.689, 3.818, 3.923, 4.024};
  EtStepDev d{};
  d.do_potential = 1; d.do_components = 1; d.met_time = 1;   // callSnowPack runs at met_hour only
  d.intercept_flag = e->k.i_option != 0;                      // tSimul.cpp:474 / 489
  d.alphaD = s->alphaD; d.sinAlpha = s->sinAlpha; d.sunaz = s->sunaz; d.Io = s->Io;
  d.cos_alpha = cos(asin(s->sinAlpha));
  d.hour = s->hour; d.first_call = s->first_call; d.dew_hum_flag = s->dew_hum_flag; d.ts_option = s->ts_option;
  d.sky_flag_from = s->sky_flag_from; d.te_met = s->te_met; d.te_eti = s->te_eti;
  d.rs_ratio = rsRatio[s->hour];
  if (et_upload_met(c, 0, s->met_potential, &d.pot)) return 1;
  d.cmp = d.pot;        // ComputeETComponents runs inside the same cell loop, before the met clock advances
  SnowConst sc = e->sc;
  sc.hourly = hourly_time_step;
  if (e->k.i_option != 0) {
    if (snow_step_with_canopy(c, d, sc)) return 1;
  } else {
    snow_cells_kernel<<<(c->n_pad + 127) / 128, 128, 0, c->stream>>>(e->k, e->st, d, sc, e->ev, e->snv, c->dv, c->n_pad);
    c->launches++;
  }
  e->gen++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int tribs_gpu_snow_sync(tribs_handle_t c, uint64_t mask, double *const *host) {
  if (!c || !host) return fail("tribs_gpu_snow_sync: null argument");
  if (!c->et || !c->et->snow_ready) return fail("tribs_gpu_snow_sync: call tribs_gpu_snow_init first");
  CK(cudaSetDevice(c->device));
  for (int k = 0; k < TRIBS_N_SNOW_FIELDS; k++) {
    if (!(mask & (1ull << k)) || !host[k]) continue;
    gather_to_sweep_kernel<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->et->snv.f[k], c->d_dev_of_sweep, c->d_stage, c->n);
    c->launches++;
    CK(cudaMemcpyAsync(host[k], c->d_stage, (size_t)c->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

extern "C" int tribs_gpu_snow_push(tribs_handle_t c, uint64_t mask, const double *const *host) {
  if (!c || !host) return fail("tribs_gpu_snow_push: null argument");
  if (!c->et || !c->et->snow_ready) return fail("tribs_gpu_snow_push: call tribs_gpu_snow_init first");
  CK(cudaSetDevice(c->device));
  for (int k = 0; k < TRIBS_N_SNOW_FIELDS; k++) {
    if (!(mask & (1ull << k)) || !host[k]) continue;
    CK(cudaMemcpyAsync(c->d_stage, host[k], (size_t)c->n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    scatter_from_sweep_kernel<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->d_stage, c->d_dev_of_sweep, c->et->snv.f[k], c->n);
    c->launches++;
    CK(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

extern "C" int tribs_gpu_snow_set_members(tribs_handle_t c, const double *members) {
  if (!c || !members) return fail("tribs_gpu_snow_set_members: null argument");
  if (!c->et || !c->et->snow_ready) return fail("tribs_gpu_snow_set_members: call tribs_gpu_snow_init first");
  EtState *e = c->et;
  e->members[0] = members[0]; e->members[1] = members[1];
  if (e->carry.members) {
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(e->carry.members, e->members, 2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

extern "C" int tribs_gpu_snow_members(tribs_handle_t c, double *members, int32_t *probe_rounds) {
  if (!c) return fail("tribs_gpu_snow_members: null argument");
  if (!c->et || !c->et->snow_ready) return fail("tribs_gpu_snow_members: call tribs_gpu_snow_init first");
  EtState *e = c->et;
  if (members) {
    members[0] = e->members[0]; members[1] = e->members[1];
    if (e->carry.members) {
      CK(cudaSetDevice(c->device));
      CK(cudaMemcpyAsync(members, e->carry.members, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
    }
  }
  if (probe_rounds) *probe_rounds = e->last_probe_rounds;
  return 0;
}

extern "C" int tribs_gpu_et_step(tribs_handle_t c, const tribs_et_step_t *s) {
  if (!c || !s) return fail("tribs_gpu_et_step: null argument");
  if (!c->et) return fail("tribs_gpu_et_step: call tribs_gpu_et_init first");
  if (c->et->snow_option) return fail("tribs_gpu_et_step: with OPTSNOW 1 the simulator calls tribs_gpu_snow_step instead");
  if (!s->do_potential && !s->do_components) return 0;
  if (s->hour < 0 || s->hour > 23) return fail("tribs_gpu_et_step: hour out of range");
  CK(cudaSetDevice(c->device));
  EtState *e = c->et;
  static const double rsRatio[24] = {3.837, 3.589, 3.21, 2.43, 1.617, 1.196, 1.067, 1.014,   // stomResist, tEvapoTrans.cpp:1494-1496
                                     0.995, 0.976, 0.976, 1.0, 1.053, 1.167, 1.354, 1.637,
                                     2.043, 2.66, 3.215, 3.507, 3.689, 3.818, 3.923, 4.024};
  EtStepDev d{};
  d.do_potential = s->do_potential && e->k.et_option; d.do_components = s->do_components; d.intercept_flag = s->intercept_flag;
  d.met_time = s->do_potential;
  d.alphaD = s->alphaD; d.sinAlpha = s->sinAlpha; d.sunaz = s->sunaz; d.Io = s->Io;
  d.cos_alpha = cos(asin(s->sinAlpha));
  d.hour = s->hour; d.first_call = s->first_call; d.dew_hum_flag = s->dew_hum_flag; d.ts_option = s->ts_option;
  d.sky_flag_from = s->sky_flag_from; d.te_met = s->te_met; d.te_eti = s->te_eti;
  d.rs_ratio = rsRatio[s->hour];
  if (d.do_potential && et_upload_met(c, 0, s->met_potential, &d.pot)) return 1;
  if (d.do_components && e->k.et_option && et_upload_met(c, 1, s->met_components, &d.cmp)) return 1;
  if (!d.do_potential && !d.do_components && !d.met_time) return 0;
  et_cells_kernel<<<(c->n_pad + 127) / 128, 128, 0, c->stream>>>(e->k, e->st, d, e->ev, c->dv, c->n_pad);
  c->launches++;
  e->gen++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int tribs_gpu_et_sync(tribs_handle_t c, uint64_t mask, double *const *host) {
  if (!c || !host) return fail("tribs_gpu_et_sync: null argument");
  if (!c->et) return fail("tribs_gpu_et_sync: call tribs_gpu_et_init first");
  CK(cudaSetDevice(c->device));
  for (int k = 0; k < TRIBS_N_ET_FIELDS; k++) {
    if (!(mask & (1ull << k)) || !host[k]) continue;
    gather_to_sweep_kernel<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->et->ev.f[k], c->d_dev_of_sweep, c->d_stage, c->n);
    c->launches++;
    CK(cudaMemcpyAsync(host[k], c->d_stage, (size_t)c->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

extern "C" int tribs_gpu_et_push(tribs_handle_t c, uint64_t mask, const double *const *host) {
  if (!c || !host) return fail("tribs_gpu_et_push: null argument");
  if (!c->et) return fail("tribs_gpu_et_push: call tribs_gpu_et_init first");
  CK(cudaSetDevice(c->device));
  for (int k = 0; k < TRIBS_N_ET_FIELDS; k++) {
    if (!(mask & (1ull << k)) || !host[k]) continue;
    CK(cudaMemcpyAsync(c->d_stage, host[k], (size_t)c->n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    scatter_from_sweep_kernel<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->d_stage, c->d_dev_of_sweep, c->et->ev.f[k], c->n);
    c->launches++;
    CK(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

extern "C" int tribs_gpu_n_stream(tribs_handle_t c) { return c ? c->n_stream : -1; }
extern "C" int tribs_gpu_stream_cells(tribs_handle_t c, int32_t *out) {
  if (!c || !out) return fail("tribs_gpu_stream_cells: null argument");
  for (int k = 0; k < c->n_stream; k++) out[k] = c->stream_cells_sweep[k];
  return 0;
}
extern "C" int tribs_gpu_retrieve_qeff(tribs_handle_t c, double current_time, double *out) {
  (void)current_time;  // the ROUTE phase of that step already latched and consumed the bin
  if (!c || !out) return fail("tribs_gpu_retrieve_qeff: null argument");
  CK(cudaSetDevice(c->device));
  if (check_sweep_abort(c)) return 1;
  CK(cudaMemcpyAsync(out, c->d_qeff_now, (size_t)(c->n_stream + 1) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}
// The device's form of the (TimeInd, Qeff) stacks (tKinemat.cpp:1023-1067): out[slot * n_bins + k] is what
// the dtReff bin ceil(current_time / dtReff) + k of stream slot `slot` holds after the ROUTE phase of
// the step that ended at current_time (k = 0 is the bin RetrieveQeff reads; it is empty again once it
// has been consumed at the end of its interval). What a restart has to carry for the channel router.
extern "C" int tribs_gpu_pending_qeff(tribs_handle_t c, double current_time, int n_bins, double *out) {
  if (!c || !out || n_bins <= 0) return fail("tribs_gpu_pending_qeff: bad argument");
  CK(cudaSetDevice(c->device));
  if (check_sweep_abort(c)) return 1;
  const int slots = c->n_stream + 1, ring = c->rc.ring;
  std::vector<double> h((size_t)slots * ring);
  CK(cudaMemcpyAsync(h.data(), c->d_ring, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  const int cur = (int)ceil(current_time / c->rc.dt_reff);
  for (int sl = 0; sl < slots; sl++)
    for (int k = 0; k < n_bins; k++)
      out[(size_t)sl * n_bins + k] = k < ring ? h[(size_t)sl * ring + ((cur + k) % ring)] : 0.0;
  return 0;
}
extern "C" int tribs_gpu_get_results(tribs_handle_t c, int which, double *out, int nbins) {
  if (!c || !out || which < 0 || which >= TRIBS_N_RESULTS) return fail("tribs_gpu_get_results: bad argument");
  CK(cudaSetDevice(c->device));
  if (check_sweep_abort(c)) return 1;
  nbins = std::min(nbins, (int)c->rc.res_limit);
  CK(cudaMemcpyAsync(out, c->d_results + (size_t)which * c->rc.res_limit, (size_t)nbins * sizeof(double),
                     cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" int tribs_gpu_get_globals(tribs_handle_t c, double *out) {
  if (!c || !out) return fail("tribs_gpu_get_globals: null argument");
  CK(cudaSetDevice(c->device));
  if (check_sweep_abort(c)) return 1;
  CK(cudaMemcpyAsync(out, c->d_globals, TRIBS_N_GLOBALS * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}


// ------------------------------------------------------------------------------------ partitioned runs
// One reach partition per GPU (reference: tGraph::partition, src/tGraph/tGraph.cpp:401-608). Every
// rank holds the whole basin (topology and static data replicated) but advances only the cells it
// owns; values of cells owned elsewhere are refreshed by the host through the halo pack/unpack
// calls below and an NCCL exchange.
extern "C" int tribs_gpu_set_partition(tribs_handle_t c, const int32_t *cell_owner, int rank) {
  if (!c || !cell_owner) return fail("tribs_gpu_set_partition: null argument");
  CK(cudaSetDevice(c->device));
  const int n = c->n, n_pad = c->n_pad, nt = c->n_tiles;
  std::vector<int32_t> owned(n_pad, 0);
  for (int i = 0; i < n; i++) owned[c->dev_of_sweep[i]] = (cell_owner[i] == rank) ? 1 : 0;
  std::vector<int32_t> h_up_rp(n_pad + 1);
  CK(cudaMemcpy(h_up_rp.data(), c->d_up_rowptr, (n_pad + 1) * 4, cudaMemcpyDeviceToHost));
  std::vector<int32_t> h_up_col(h_up_rp[n_pad]), order(nt);
  if (!h_up_col.empty()) CK(cudaMemcpy(h_up_col.data(), c->d_up_col, h_up_col.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(order.data(), c->d_tile_order, nt * 4, cudaMemcpyDeviceToHost));
  // a cell depends on another rank if an upslope contributor is not owned here, or depends itself
  // (ascending device index is a topological order)
  std::vector<uint8_t> dep(n_pad, 0);
  for (int i = 0; i < n_pad; i++)
    for (int k = h_up_rp[i]; k < h_up_rp[i + 1]; k++) {
      const int j = h_up_col[k];
      if (!owned[j] || dep[j]) { dep[i] = 1; break; }
    }
  std::vector<int32_t> local, remote;
  for (int q = 0; q < nt; q++) {
    const int t = order[q];
    bool any = false, anydep = false;
    for (int l = 0; l < kTile; l++) {
      const int i = t * kTile + l;
      if (owned[i]) { any = true; anydep = anydep || dep[i]; }
    }
    if (!any) continue;
    (anydep ? remote : local).push_back(t);
  }
  CK(cudaMemcpy(c->d_owned, owned.data(), (size_t)n_pad * 4, cudaMemcpyHostToDevice));
  if (dev_upload(c, &c->d_tiles_local, local) || dev_upload(c, &c->d_tiles_remote, remote)) return 1;
  c->n_tiles_local = (int)local.size();
  c->n_tiles_remote = (int)remote.size();
  c->partitioned = true;
  c->rank = rank;
  return 0;
}

__global__ void halo_pack_kernel(DynView d, int kind, const int32_t *__restrict__ idx, int n, double *buf) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int i = idx[k];
  if (kind == TRIBS_HALO_QPOUT) buf[k] = d.f[TRIBS_F_Qpout][i];
  else if (kind == TRIBS_HALO_NWT) buf[k] = d.f[TRIBS_F_NwtNew][i];
  else
    for (int q = 0; q < 7; q++) buf[(size_t)q * n + k] = d.f[TRIBS_F_NwtNew + q][i];
}
__global__ void halo_unpack_kernel(DynView d, int kind, const int32_t *__restrict__ idx, int n, const double *buf,
                                   PubWord *pub, long long epoch) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int i = idx[k];
  if (kind == TRIBS_HALO_QPOUT) {
    d.f[TRIBS_F_Qpout][i] = buf[k];
    pub_store(pub + i, buf[k], epoch);   // the remote pass of the sweep reads it like a local outflow
  } else if (kind == TRIBS_HALO_NWT) d.f[TRIBS_F_NwtNew][i] = buf[k];
  else
    for (int q = 0; q < 7; q++) d.f[TRIBS_F_NwtNew + q][i] = buf[(size_t)q * n + k];
}
extern "C" int tribs_gpu_halo_pack(tribs_handle_t c, int kind, const int32_t *idx_dev, int n, double *buf_dev) {
  if (!c || (n > 0 && (!idx_dev || !buf_dev))) return fail("tribs_gpu_halo_pack: null argument");
  CK(cudaSetDevice(c->device));
  if (n > 0) { halo_pack_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(c->dv, kind, idx_dev, n, buf_dev); c->launches++; }
  CK(cudaGetLastError());
  return 0;
}
extern "C" int tribs_gpu_halo_unpack(tribs_handle_t c, int kind, const int32_t *idx_dev, int n, const double *buf_dev) {
  if (!c || (n > 0 && (!idx_dev || !buf_dev))) return fail("tribs_gpu_halo_unpack: null argument");
  CK(cudaSetDevice(c->device));
  if (n > 0) {
    halo_unpack_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(c->dv, kind, idx_dev, n, buf_dev, c->d_pub, c->epoch);
    c->launches++;
  }
  CK(cudaGetLastError());
  return 0;
}
extern "C" int tribs_gpu_set_global(tribs_handle_t c, int which, double value) {
  if (!c || which < 0 || which >= TRIBS_N_GLOBALS) return fail("tribs_gpu_set_global: bad argument");
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpyAsync(c->d_globals + which, &value, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" int tribs_gpu_padded_cells(tribs_handle_t c) { return c ? c->n_pad : -1; }

extern "C" int tribs_gpu_sweep_levels(tribs_handle_t c) { return c ? c->n_levels : -1; }
extern "C" int tribs_gpu_device_index(tribs_handle_t c, int32_t *out) {
  if (!c || !out) return fail("tribs_gpu_device_index: null argument");
  for (int i = 0; i < c->n; i++) out[i] = c->dev_of_sweep[i];
  return 0;
}
extern "C" int64_t tribs_gpu_launch_count(tribs_handle_t c) { return c ? c->launches : -1; }
extern "C" int tribs_gpu_set_timing(tribs_handle_t c, int enabled) {
  if (!c) return fail("tribs_gpu_set_timing: null handle");
  c->timing = enabled != 0;
  return 0;
}
extern "C" int tribs_gpu_phase_ms(tribs_handle_t c, double *out4, int reset) {
  if (!c || !out4) return fail("tribs_gpu_phase_ms: null argument");
  CK(cudaSetDevice(c->device));
  drain_timers(c);
  for (int k = 0; k < 4; k++) { out4[k] = c->phase_ms[k]; if (reset) c->phase_ms[k] = 0.0; }
  return 0;
}
extern "C" int tribs_gpu_wait(tribs_handle_t c) {
  if (!c) return fail("tribs_gpu_wait: null handle");
  CK(cudaSetDevice(c->device));
  return check_sweep_abort(c);   // synchronises; a batch that aborted (tribs_gpu_run reports late) fails here
}
extern "C" void *tribs_gpu_stream(tribs_handle_t c) { return c ? (void *)c->stream : nullptr; }
static void tribs_free_pipe(tribs_ctx *c);
static void tribs_free_snapshot(tribs_ctx *c);
static void tribs_free_channel(tribs_ctx *c);
extern "C" int tribs_gpu_destroy(tribs_handle_t c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  drain_timers(c);
  for (cudaEvent_t e : c->event_pool) cudaEventDestroy(e);
  for (void *p : c->allocs) cudaFree(p);
  cudaStreamDestroy(c->stream);
  tribs_free_pipe(c);
  tribs_free_et(c);
  tribs_free_snapshot(c);
  tribs_free_channel(c);
  delete c;
  return 0;
}

#include "pipeline.cuh"
#include "checkpoint.cuh"
#include "channel.cuh"
