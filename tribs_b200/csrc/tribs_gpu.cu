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
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <algorithm>
#include <numeric>
#include <string>
#include <vector>

#include "cell_update.cuh"

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
struct RouteConst {
  double velcoef, velratio, flowexp, baseflow, kincoef, dt_reff, outlet_contr_area;
  double tstep, output_interval;
  int32_t res_limit, ring;
};

struct tribs_ctx {
  int device = 0;
  int n = 0, n_pad = 0, n_edges = 0, n_levels = 0, n_tiles = 0, n_stream = 0, n_late = 0;
  int max_degree = 0;
  cudaStream_t stream = nullptr;
  // host maps
  std::vector<int32_t> dev_of_sweep, sweep_of_dev, stream_cells_sweep;
  // device static
  StaticView sv{};
  DynView dv{};
  std::vector<void *> allocs;
  int32_t *d_up_rowptr = nullptr, *d_up_col = nullptr, *d_ready = nullptr, *d_tile_counter = nullptr;
  int32_t *d_abort = nullptr, *d_sweep_of_dev = nullptr, *d_dev_of_sweep = nullptr;
  int32_t *d_late_rowptr = nullptr, *d_late_col = nullptr;
  double *d_tile_sums = nullptr, *d_globals = nullptr, *d_wt_elev = nullptr, *d_transm = nullptr;
  double *d_stage = nullptr;
  int32_t *d_flags = nullptr;  // [0] gather overflow
  // routing
  int32_t *d_sn_rowptr = nullptr, *d_sn_cells = nullptr, *d_sn_dev = nullptr;
  double *d_ring = nullptr, *d_qeff_now = nullptr, *d_results = nullptr, *d_route_part = nullptr;
  int route_blocks = 0, route_window = 0;
  RouteConst rc{};
  ModelConst mc{};
  int last_sweep_dev = -1;   // device index of the last cell in sweep order
  int last_hill_dev = -1;    // device index of the last hillslope cell in sweep order
  int epoch = 0;
  int unsat_blocks = 0, unsat_threads = 128;
  bool gw_dirty = false;
  bool new_valid = true;     // false between Reset() and the next zone update: New == Old there
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
template <class T>
static int dev_upload(tribs_ctx *c, T **p, const std::vector<T> &h) {
  if (dev_alloc(c, p, h.size())) return 1;
  if (!h.empty()) CK(cudaMemcpy(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

// ------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ int ld_acquire(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

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
                   const int32_t *__restrict__ up_col, int32_t *ready, int epoch, int n_tiles,
                   int32_t *tile_counter, int32_t *abort_flag, double *tile_sums, double *globals,
                   int last_sweep_dev) {
  const int lane = threadIdx.x & 31;
  for (;;) {
    int tile = 0;
    if (lane == 0) tile = atomicAdd(tile_counter, 1);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= n_tiles) break;
    const int i = tile * kTile + lane;
    UnsatSums sums{0.0, 0.0, 0.0, 0.0};
    if (s.boundary[i] >= 0) {
      UnsatCellInputs u;
      unsat_cell_load(u, s, d, i, m, t);
      // lateral inflow: new Qpout of the upslope cells swept earlier, summed in sweep order
      double qpin = 0.0;
      const int k1 = up_rowptr[i + 1];
      for (int k = up_rowptr[i]; k < k1; k++) {
        const int j = up_col[k];
        unsigned spins = 0;
        while (ld_acquire(ready + j) != epoch) {
          __nanosleep(40);
          if (((++spins) & 1023u) == 0u) {
            if (*(volatile int32_t *)abort_flag) break;
            if (spins > (1u << 24)) { atomicExch(abort_flag, 1); break; }
          }
        }
        qpin += __ldcg(d.f[TRIBS_F_Qpout] + j);
      }
      int state;
      const double qout = unsat_cell_advance(u, qpin, &state);
      // publish the outflow first: the downslope cell is waiting for it
      d.f[TRIBS_F_Qpout][i] = qout;
      __threadfence();
      st_release(ready + i, epoch);
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

// fixed-order reduction of per-tile partials: one block, each thread strides, then a shared tree
__global__ void __launch_bounds__(256)
reduce_sums_kernel(const double *__restrict__ tile_sums, int n_tiles, int width, double *globals,
                   int g0, int g1, int g2, int g3) {
  __shared__ double sh[4][256];
  double acc[4] = {0, 0, 0, 0};
  for (int tix = threadIdx.x; tix < n_tiles; tix += 256)
    for (int k = 0; k < width; k++) acc[k] += tile_sums[width * tix + k];
  for (int k = 0; k < 4; k++) sh[k][threadIdx.x] = acc[k];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int k = 0; k < 4; k++) sh[k][threadIdx.x] += sh[k][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int g[4] = {g0, g1, g2, g3};
    for (int k = 0; k < width; k++)
      if (g[k] >= 0) globals[g[k]] += sh[k][0];
  }
}

// ------------------------------------------------------------------------------------ SAT
__global__ void __launch_bounds__(256)
sat_prepare_kernel(StaticView s, DynView d, double *wt_elev, double *transm, int n_pad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad || s.boundary[i] < 0) return;
  sat_prepare_cell(s, d, i, wt_elev, transm);
}

template <int MAXC>
__global__ void __launch_bounds__(128)
sat_cells_kernel(StaticView s, DynView d, ModelConst m, StepConst t, const double *__restrict__ wt_elev,
                 const double *__restrict__ transm, int n_pad, double *tile_sums, int32_t *flags,
                 const double *__restrict__ globals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  SatSums sums{0.0, 0.0, 0.0};
  t.psrf_last = globals[TRIBS_G_psrf_last];
  if (i < n_pad && s.boundary[i] >= 0) {
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

__global__ void __launch_bounds__(256)
route_cells_kernel(StaticView s, DynView d, RouteConst rc, double qstrm_outlet, int n_pad, int last_hill_dev,
                   double *globals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  int bnd = s.boundary[i];
  if (bnd < 0) return;
  if (bnd == 0) {
    double hv = hill_velocity(s, d, rc, s.stream_node[i], qstrm_outlet);
    d.f[TRIBS_F_FlowVelocity][i] = hv;
    if (d.f[TRIBS_F_srf][i] > 0.0) d.f[TRIBS_F_TTime][i] = s.hillpath[i] / hv / 3600.;
    if (i == last_hill_dev) globals[TRIBS_G_hillvel_last] = hv;
  }
}
// stream cells keep the velocity of the last hillslope cell swept before them (stale member
// `hillvel`, tKinemat.cpp:1084); with the reference's node order that is the last hillslope cell.
__global__ void __launch_bounds__(256)
route_stream_velocity_kernel(StaticView s, DynView d, int n_pad, const double *globals, double fallback, int have_hill) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad || s.boundary[i] != 3) return;
  d.f[TRIBS_F_FlowVelocity][i] = have_hill ? globals[TRIBS_G_hillvel_last] : fallback;
}

// (2) hillslope runoff -> (stream node, dtReff bin). One warp per stream node; its contributors
//     are pre-sorted by hillslope path, so bins are non-decreasing along the segment and each
//     (node, bin) sum is a run: warp-shuffle segmented scan, last lane of a run accumulates into
//     the node's ring of bins. No atomics; the summation order is fixed by the static sort.
__global__ void __launch_bounds__(128)
route_stream_bins_kernel(StaticView s, DynView d, RouteConst rc, const int32_t *__restrict__ sn_rowptr,
                         const int32_t *__restrict__ sn_cells, const int32_t *__restrict__ sn_dev, int n_slots,
                         double now, double qstrm_outlet, double *ring, double *qeff_now) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_slots) return;
  const int node = sn_dev[warp];  // device index of the stream cell, -1 = the outlet node slot
  const double dtreff = rc.dt_reff;
  const int cur = (int)ceil(now / dtreff);
  const bool at_end = ((int)ceil(now / dtreff) == (int)floor(now / dtreff));
  double *my_ring = ring + (size_t)warp * rc.ring;
  const double same_box_div = at_end ? (rc.tstep * 3600.) : ((cur * dtreff - now + rc.tstep) * 3600.);

  // the stream cell's own runoff goes to the current bin with zero travel time (1001-1016)
  if (lane == 0 && node >= 0) {
    double srf = d.f[TRIBS_F_srf][node];
    if (srf > 0.0) {
      double vr = srf * s.varea[node] / 1000.;
      if (srf > 999999) vr = 0.;
      my_ring[cur % rc.ring] += vr / same_box_div;
    }
  }
  __syncwarp();
  const double hv = hill_velocity(s, d, rc, node, qstrm_outlet);
  const int k0 = sn_rowptr[warp], k1 = sn_rowptr[warp + 1];
  for (int base = k0; base < k1; base += 32) {
    int k = base + lane;
    int bin = 0x7fffffff;
    double v = 0.0;
    if (k < k1) {
      int cidx = sn_cells[k];
      double srf = d.f[TRIBS_F_srf][cidx];
      if (srf > 0.0) {
        double vr = srf * s.varea[cidx] / 1000.;
        if (srf > 999999) vr = 0.;
        double tt = s.hillpath[cidx] / hv / 3600.;
        bin = (int)ceil((now + tt) / dtreff);
        if (bin == cur) vr /= same_box_div;
        else vr /= (dtreff * 3600.);
        v = vr;
      } else
        bin = -1;  // no contribution
    }
    // runs of equal bins: inclusive segmented scan over lanes
    int prev_bin = __shfl_up_sync(0xffffffffu, bin, 1);
    bool head = (lane == 0) || (prev_bin != bin);
    unsigned heads = __ballot_sync(0xffffffffu, head);
    double acc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double up = __shfl_up_sync(0xffffffffu, acc, o);
      // add only if lane-o is still inside my run: no head strictly above lane-o up to lane
      unsigned mask_between = (lane >= o) ? (heads >> (lane - o + 1)) & ((1u << o) - 1u) : 1u;
      if (lane >= o && mask_between == 0u) acc += up;
    }
    int next_bin = __shfl_down_sync(0xffffffffu, bin, 1);
    bool tail = (lane == 31) || (next_bin != bin);
    if (tail && bin >= 0 && bin != 0x7fffffff) my_ring[bin % rc.ring] += acc;
    __syncwarp();
  }
  __syncwarp();
  // what the channel router reads this step, and consumption at the end of a dtReff interval
  if (lane == 0) {
    double val = my_ring[cur % rc.ring];
    qeff_now[warp] = val;
    if (at_end) my_ring[cur % rc.ring] = 0.0;
  }
}

// (3) tFlowNet::SurfaceFlow accumulations. Stage 1: every block reduces its cells into
//     [window] volume bins x 5 runoff types + 10 area-weighted means, deterministically
//     (per-warp slabs filled lane by lane, slabs added in warp order). Stage 2: one block sums
//     the block partials in block order into the tFlowResults arrays.
constexpr int kRouteWarps = 8;
constexpr int kMeans = 10;  // crr frac msm msmRt msmU sat mgw qunsat rmax rmin
__global__ void __launch_bounds__(kRouteWarps * 32)
route_basin_stage1_kernel(StaticView s, DynView d, RouteConst rc, ModelConst m, double now, int n_pad, int window,
                          int bin_base, double *part) {
  extern __shared__ double sh[];  // [kRouteWarps][5*window + kMeans]
  const int stride = 5 * window + kMeans;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *slab = sh + warp * stride;
  for (int k = lane; k < stride; k += 32) slab[k] = 0.0;
  if (lane == 0) { slab[5 * window + 8] = -1.0e300; slab[5 * window + 9] = 1.0e300; }
  __syncwarp();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = (i < n_pad) && (s.boundary[i] >= 0);
  double vol[5] = {0, 0, 0, 0, 0};
  int init = 0, end = -1;
  double w_init = 0, w_mid = 0, w_end = 0;
  double means[kMeans];
  for (int k = 0; k < kMeans; k++) means[k] = 0.0;
  means[8] = -1.0e300; means[9] = 1.0e300;
  if (act) {
    const double area = s.varea[i], areaf = area / m.bas_area_flow;
    const double srf = d.f[TRIBS_F_srf][i];
    const double dcalc = rc.tstep, dres = rc.output_interval, sec = rc.output_interval * 3600.0;
    if (srf > 0.0) {
      if (rc.flowexp > 0.0) {
        double hv = rc.velcoef / rc.velratio, sv = rc.velcoef;
        d.f[TRIBS_F_TTime][i] = (s.hillpath[i] / hv + s.streampath[i] / sv) / 3600.0;
      }
      const double tt = d.f[TRIBS_F_TTime][i];
      vol[0] = srf * area / 1000.0;
      vol[1] = d.f[TRIBS_F_hsrf][i] * area / 1000.0;
      vol[2] = d.f[TRIBS_F_sbsrf][i] * area / 1000.0;
      vol[3] = d.f[TRIBS_F_psrf][i] * area / 1000.0;
      vol[4] = d.f[TRIBS_F_satsrf][i] * area / 1000.0;
      init = (int)floor((tt + .01 * dcalc + now) / dres);
      end = (int)floor((tt + .99 * dcalc + now) / dres);
      if (init == end) { w_init = 1.0 / sec; }
      else {
        double dint = (init + 1) * dres - (now + tt);
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
  }
  // lane-serial fill keeps the order fixed (lane 0 first)
  for (int l = 0; l < 32; l++) {
    if (lane == l && act) {
      if (end >= init) {
        for (int b = init; b <= end; b++) {
          int off = b - bin_base;
          if (off < 0 || off >= window) continue;
          double wgt = (b == init) ? w_init : ((b == end) ? w_end : w_mid);
          // volume * weight written as (value * fraction) / seconds in the reference; the
          // fraction and 1/seconds are folded into wgt here (<= 1 ulp difference)
          for (int k = 0; k < 5; k++)
            if (vol[k] != 0.0) slab[k * window + off] += vol[k] * wgt;
        }
      }
      for (int k = 0; k < 8; k++) slab[5 * window + k] += means[k];
      if (means[8] > slab[5 * window + 8]) slab[5 * window + 8] = means[8];
      if (means[9] < slab[5 * window + 9]) slab[5 * window + 9] = means[9];
    }
    __syncwarp();
  }
  __syncthreads();
  for (int k = threadIdx.x; k < stride; k += blockDim.x) {
    double acc = sh[k];
    if (k == 5 * window + 8) { for (int wv = 1; wv < kRouteWarps; wv++) acc = fmax(acc, sh[wv * stride + k]); }
    else if (k == 5 * window + 9) { for (int wv = 1; wv < kRouteWarps; wv++) acc = fmin(acc, sh[wv * stride + k]); }
    else { for (int wv = 1; wv < kRouteWarps; wv++) acc += sh[wv * stride + k]; }
    part[(size_t)blockIdx.x * stride + k] = acc;
  }
}

__global__ void __launch_bounds__(256)
route_basin_stage2_kernel(const double *__restrict__ part, int n_blocks, int window, int bin_base, int cur_bin,
                          int limit, double *results) {
  const int stride = 5 * window + kMeans;
  for (int k = threadIdx.x; k < stride; k += blockDim.x) {
    double acc;
    if (k == 5 * window + 8) { acc = -1.0e300; for (int b = 0; b < n_blocks; b++) acc = fmax(acc, part[(size_t)b * stride + k]); }
    else if (k == 5 * window + 9) { acc = 1.0e300; for (int b = 0; b < n_blocks; b++) acc = fmin(acc, part[(size_t)b * stride + k]); }
    else { acc = 0.0; for (int b = 0; b < n_blocks; b++) acc += part[(size_t)b * stride + k]; }
    if (k < 5 * window) {
      int type = k / window, bin = bin_base + (k % window);
      if (bin >= 0 && bin < limit && acc != 0.0) results[(size_t)type * limit + bin] += acc;
    } else if (cur_bin >= 0 && cur_bin < limit) {
      int q = k - 5 * window;
      // order of TRIBS_R_*: crr rmax rmin frac msm msmRt msmU mgw sat qunsat
      const int map[kMeans] = {TRIBS_R_crr, TRIBS_R_frac, TRIBS_R_msm, TRIBS_R_msmRt, TRIBS_R_msmU,
                               TRIBS_R_sat, TRIBS_R_mgw, TRIBS_R_qunsat, TRIBS_R_rmax, TRIBS_R_rmin};
      double *dst = results + (size_t)map[q] * limit + cur_bin;
      if (q == 8) { if (acc > *dst) *dst = acc; }
      else if (q == 9) { if (acc <= *dst) *dst = acc; }
      else *dst += acc;
    }
  }
}

// ------------------------------------------------------------------------------------ misc kernels
__global__ void fill_kernel(double *p, double v, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
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

extern "C" int tribs_gpu_init(const tribs_mesh_t *mesh, const tribs_params_t *prm, const tribs_state_t *init,
                              const tribs_part_t *part, tribs_handle_t *out) {
  if (!mesh || !prm || !out) return fail("tribs_gpu_init: null argument");
  if (part && part->nranks > 1) return fail("tribs_gpu_init: partitioned runs are not built yet");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("tribs_gpu_init: no CUDA device (this library has no CPU path)");
  const int n = mesh->n_cells;
  if (n <= 0) return fail("tribs_gpu_init: empty mesh");
  tribs_ctx *c = new tribs_ctx();
  CK(cudaGetDevice(&c->device));
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->n = n;
  c->n_edges = mesh->n_edges;

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
  std::vector<int32_t> count(nlev, 0), start(nlev + 1, 0);
  for (int i = 0; i < n; i++) count[level[i]]++;
  for (int l = 0; l < nlev; l++) start[l + 1] = start[l] + ((count[l] + kTile - 1) / kTile) * kTile;
  const int n_pad = start[nlev];
  c->n_pad = n_pad;
  c->n_tiles = n_pad / kTile;
  c->dev_of_sweep.assign(n, -1);
  c->sweep_of_dev.assign(n_pad, -1);
  {
    std::vector<int32_t> fillp(start.begin(), start.end() - 1);
    for (int i = 0; i < n; i++) {  // stable: sweep order inside a level
      int dI = fillp[level[i]]++;
      c->dev_of_sweep[i] = dI;
      c->sweep_of_dev[dI] = i;
    }
  }
  auto D = [&](int sweep) { return sweep < 0 ? -1 : c->dev_of_sweep[sweep]; };
  c->last_sweep_dev = c->dev_of_sweep[n - 1];

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
  UPI(boundary, bnd) UPI(flow_dest, fdest) UPI(stream_node, snode)
  UPD(cosv, cosv) UPD(sinv, sinv) UPD(cos_gw, cosgw) UPD(vel_mm, vel)
  UPD(varea, permute_d(mesh->varea, 1.0)) UPD(z, permute_d(mesh->z, 0.0)) UPD(bedrock, permute_d(prm->bedrock, 1.0))
  UPD(hillpath, permute_d(mesh->hillpath, 0.0)) UPD(streampath, permute_d(mesh->streampath, 0.0))
  UPD(contr_area, permute_d(mesh->contr_area, 1.0))
  UPD(ks, permute_d(prm->Ks, 1.0)) UPD(ths, permute_d(prm->ThetaS, 0.5)) UPD(thr, permute_d(prm->ThetaR, 0.1))
  UPD(lam, permute_d(prm->PoreSize, 0.5)) UPD(psib, permute_d(prm->AirEBubPres, -100.0)) UPD(f, permute_d(prm->DecayF, 0.01))
  UPD(ar, permute_d(prm->SatAnRatio, 1.0)) UPD(uar, permute_d(prm->UnsatAnRatio, 1.0))

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

  // ---- routing structures: contributors of every stream cell sorted by (hillpath, sweep pos)
  {
    std::vector<int32_t> scells;  // sweep indices of stream cells, sweep order
    for (int i = 0; i < n; i++) if (mesh->boundary[i] == 3) scells.push_back(i);
    c->stream_cells_sweep = scells;
    c->n_stream = (int)scells.size();
    const int slots = c->n_stream + 1;  // last slot = the outlet node
    std::vector<int32_t> slot_of(n, -1);
    for (int k = 0; k < c->n_stream; k++) slot_of[scells[k]] = k;
    std::vector<std::vector<int32_t>> members(slots);
    int last_hill = -1;
    for (int i = 0; i < n; i++) {
      if (mesh->boundary[i] != 0) continue;
      last_hill = i;
      int sn = mesh->stream_node[i];
      int slot = (sn >= 0 && slot_of[sn] >= 0) ? slot_of[sn] : c->n_stream;
      members[slot].push_back(i);
    }
    c->last_hill_dev = last_hill >= 0 ? c->dev_of_sweep[last_hill] : -1;
    std::vector<int32_t> rp(slots + 1, 0), cells, sdev(slots, -1);
    double max_path = 0.0;
    for (int k = 0; k < slots; k++) {
      auto &mm = members[k];
      std::stable_sort(mm.begin(), mm.end(), [&](int a, int b) { return mesh->hillpath[a] < mesh->hillpath[b]; });
      for (int i : mm) { cells.push_back(c->dev_of_sweep[i]); max_path = std::max(max_path, mesh->hillpath[i]); }
      rp[k + 1] = (int32_t)cells.size();
      if (k < c->n_stream) sdev[k] = c->dev_of_sweep[scells[k]];
    }
    if (dev_upload(c, &c->d_sn_rowptr, rp) || dev_upload(c, &c->d_sn_cells, cells) || dev_upload(c, &c->d_sn_dev, sdev)) return 1;
    RouteConst &rc = c->rc;
    rc.velcoef = prm->velcoef; rc.velratio = prm->velratio; rc.flowexp = prm->flowexp; rc.baseflow = prm->baseflow;
    rc.kincoef = prm->kincoef; rc.dt_reff = prm->dtReff; rc.outlet_contr_area = prm->outlet_contr_area;
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
    std::vector<double> r0((size_t)TRIBS_N_RESULTS * rc.res_limit, 0.0);
    for (int k = 0; k < rc.res_limit; k++) r0[(size_t)TRIBS_R_rmin * rc.res_limit + k] = 9999.99;  // tFlowResults.cpp:224
    CK(cudaMemcpy(c->d_results, r0.data(), r0.size() * sizeof(double), cudaMemcpyHostToDevice));
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
    c->route_blocks = std::max(1, (n_pad + kRouteWarps * 32 - 1) / (kRouteWarps * 32));
    size_t stride = 5 * (size_t)c->route_window + kMeans;
    if (kRouteWarps * stride * sizeof(double) > 200 * 1024)
      return fail("tribs_gpu_init: basin travel-time window too wide for shared memory");
    if (dev_alloc(c, &c->d_route_part, (size_t)c->route_blocks * stride)) return 1;
    CK(cudaFuncSetAttribute(route_basin_stage1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)(kRouteWarps * stride * sizeof(double))));
  }
  c->mc.int_storm_max = prm->IntStormMAX; c->mc.surface_depth = prm->surfaceSoilDepth; c->mc.root_depth = prm->rootZoneDepth;
  c->mc.bas_area = prm->BasArea; c->mc.bas_area_flow = prm->BasArea_flow;
  c->mc.et_option = prm->EToption; c->mc.i_option = prm->Ioption; c->mc.sn_opt = prm->SnOpt; c->mc.rdstr_option = prm->RdstrOption;

  // ---- dynamic fields
  for (int k = 0; k < TRIBS_N_FIELDS; k++) {
    double *p;
    if (dev_alloc(c, &p, (size_t)n_pad)) return 1;
    std::vector<double> h(n_pad, 0.0);
    const double *src = (init && init->f[k]) ? init->f[k] : nullptr;
    if (k == TRIBS_F_TTime && !src) src = mesh->ttime;
    if (src) for (int i = 0; i < n; i++) h[c->dev_of_sweep[i]] = src[i];
    CK(cudaMemcpy(p, h.data(), (size_t)n_pad * sizeof(double), cudaMemcpyHostToDevice));
    c->dv.f[k] = p;
  }
  if (dev_alloc(c, &c->d_ready, (size_t)n_pad) || dev_alloc(c, &c->d_tile_counter, 4) || dev_alloc(c, &c->d_abort, 4) ||
      dev_alloc(c, &c->d_flags, 4) || dev_alloc(c, &c->d_tile_sums, (size_t)4 * c->n_tiles) ||
      dev_alloc(c, &c->d_globals, (size_t)TRIBS_N_GLOBALS) || dev_alloc(c, &c->d_wt_elev, (size_t)n_pad) ||
      dev_alloc(c, &c->d_transm, (size_t)n_pad) || dev_alloc(c, &c->d_stage, (size_t)n_pad))
    return 1;
  CK(cudaMemset(c->d_ready, 0, (size_t)n_pad * sizeof(int32_t)));
  CK(cudaMemset(c->d_tile_counter, 0, 16));
  CK(cudaMemset(c->d_abort, 0, 16));
  CK(cudaMemset(c->d_flags, 0, 16));
  CK(cudaMemset(c->d_globals, 0, TRIBS_N_GLOBALS * sizeof(double)));
  CK(cudaMemset(c->d_tile_sums, 0, (size_t)4 * c->n_tiles * sizeof(double)));

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
  CK(cudaDeviceSynchronize());
  *out = c;
  return 0;
}

// ------------------------------------------------------------------------------------ stepping
static int check_sweep_abort(tribs_ctx *c) {
  int32_t flags[2] = {0, 0};
  CK(cudaMemcpyAsync(&flags[0], c->d_abort, 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(&flags[1], c->d_flags, 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (flags[0]) return fail("unsat sweep aborted: an upslope cell never published its outflow (schedule bug)");
  if (flags[1]) return fail("saturated zone: more non-zero edge fluxes at a cell than the gather buffer holds");
  return 0;
}

static int upload_sweep_field(tribs_ctx *c, const double *host, int field) {
  CK(cudaMemcpyAsync(c->d_stage, host, (size_t)c->n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  scatter_from_sweep_kernel<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->d_stage, c->d_dev_of_sweep, c->dv.f[field], c->n);
  c->launches++;
  return 0;
}

extern "C" int tribs_gpu_step(tribs_handle_t c, uint32_t mask, const tribs_step_t *s) {
  if (!c || !s) return fail("tribs_gpu_step: null argument");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const int n_pad = c->n_pad;
  if (s->rain_mode == TRIBS_RAIN_UNIFORM) {
    fill_kernel<<<(n_pad + 255) / 256, 256, 0, st>>>(c->dv.f[TRIBS_F_Rain], s->rain_uniform, n_pad);
    c->launches++;
  } else if (s->rain_mode == TRIBS_RAIN_PER_CELL) {
    if (!s->rain) return fail("tribs_gpu_step: rain_mode PER_CELL without a rain array");
    if (upload_sweep_field(c, s->rain, TRIBS_F_Rain)) return 1;
  }
  if (s->qstrm && c->rc.flowexp > 0.0)
    if (upload_sweep_field(c, s->qstrm, TRIBS_F_Qstrm)) return 1;

  StepConst t{s->dt, s->dt_gw, s->current_time, (double)s->elapsed_steps, s->hourly_reset, 0.0};

  if (mask & TRIBS_PHASE_UNSAT) {
    PhaseTimer pt(c, 0);
    c->epoch++;
    CK(cudaMemsetAsync(c->d_tile_counter, 0, 4, st));
    int epoch = c->epoch, n_tiles = c->n_tiles, last = c->last_sweep_dev;
    void *args[] = {&c->sv, &c->dv, &c->mc, &t, &c->d_up_rowptr, &c->d_up_col, &c->d_ready, &epoch, &n_tiles,
                    &c->d_tile_counter, &c->d_abort, &c->d_tile_sums, &c->d_globals, &last};
    CK(cudaLaunchCooperativeKernel((void *)unsat_sweep_kernel, dim3(c->unsat_blocks), dim3(c->unsat_threads), args, 0, st));
    c->launches++;
    if (c->n_late > 0) {
      unsat_late_qpin_kernel<<<(n_pad + 255) / 256, 256, 0, st>>>(c->dv, c->d_late_rowptr, c->d_late_col, n_pad);
      c->launches++;
    }
    reduce_sums_kernel<<<1, 256, 0, st>>>(c->d_tile_sums, c->n_tiles, 4, c->d_globals, TRIBS_G_Stok, TRIBS_G_TotRain,
                                           TRIBS_G_TotGWchange, TRIBS_G_TotMoist);
    c->launches++;
    c->new_valid = true;
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
    reduce_sums_kernel<<<1, 256, 0, st>>>(c->d_tile_sums, c->n_tiles, 4, c->d_globals, TRIBS_G_Stok, -1,
                                           TRIBS_G_TotGWchange, TRIBS_G_TotMoist);
    c->launches += 3;
    c->gw_dirty = true;
    c->new_valid = true;
  }
  if (mask & TRIBS_PHASE_ROUTE) {
    PhaseTimer pt(c, 2);
    const double now = s->current_time;
    route_cells_kernel<<<(n_pad + 255) / 256, 256, 0, st>>>(c->sv, c->dv, c->rc, s->qstrm_outlet, n_pad, c->last_hill_dev, c->d_globals);
    route_stream_velocity_kernel<<<(n_pad + 255) / 256, 256, 0, st>>>(c->sv, c->dv, n_pad, c->d_globals,
                                                                      c->rc.velcoef / c->rc.velratio, c->last_hill_dev >= 0);
    const int slots = c->n_stream + 1;
    route_stream_bins_kernel<<<(slots * 32 + 127) / 128, 128, 0, st>>>(c->sv, c->dv, c->rc, c->d_sn_rowptr, c->d_sn_cells,
                                                                       c->d_sn_dev, slots, now, s->qstrm_outlet, c->d_ring, c->d_qeff_now);
    const int window = c->route_window;
    const int cur_bin0 = (int)floor((0.0 - .01 * c->rc.tstep + now) / c->rc.output_interval);
    const int cur_bin1 = (int)floor((0.0 - .99 * c->rc.tstep + now) / c->rc.output_interval);
    const int cur_bin = (cur_bin0 == cur_bin1) ? cur_bin0 : -1;   // store_saturation only writes when init==end
    const int bin_base = (int)floor((now + .01 * c->rc.tstep) / c->rc.output_interval);
    size_t shmem = (size_t)kRouteWarps * (5 * (size_t)window + kMeans) * sizeof(double);
    route_basin_stage1_kernel<<<c->route_blocks, kRouteWarps * 32, shmem, st>>>(c->sv, c->dv, c->rc, c->mc, now, n_pad, window,
                                                                                 bin_base, c->d_route_part);
    route_basin_stage2_kernel<<<1, 256, 0, st>>>(c->d_route_part, c->route_blocks, window, bin_base, cur_bin,
                                                 c->rc.res_limit, c->d_results);
    c->launches += 5;
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
  CK(cudaMemcpyAsync(out, c->d_qeff_now, (size_t)(c->n_stream + 1) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" int tribs_gpu_get_results(tribs_handle_t c, int which, double *out, int nbins) {
  if (!c || !out || which < 0 || which >= TRIBS_N_RESULTS) return fail("tribs_gpu_get_results: bad argument");
  CK(cudaSetDevice(c->device));
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
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" void *tribs_gpu_stream(tribs_handle_t c) { return c ? (void *)c->stream : nullptr; }
extern "C" int tribs_gpu_destroy(tribs_handle_t c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  drain_timers(c);
  for (cudaEvent_t e : c->event_pool) cudaEventDestroy(e);
  for (void *p : c->allocs) cudaFree(p);
  cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}
