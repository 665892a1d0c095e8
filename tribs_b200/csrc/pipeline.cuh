// pipeline.cuh — tribs_gpu_run: pipelined batches of time steps, on one GPU or across reach partitions.
// (Included by tribs_gpu.cu; not a stand-alone header.)
//
// A batch is a space-time dataflow over (phase, 32-cell tile) tasks with device-side ready queues.
// Phase with tag T reads the state buffer T%2 and writes buffer (T+1)%2; the new Qpout of a cell
// is handed to its receiver through qpub[T%2]. At tile granularity a task (phase p, tile t) needs
//   UNSAT: (p, u) for every tile u holding an upslope contributor of a cell of t         [Qpin]
//          and, from the previous phase p-1: t itself and every tile holding a flow_dest of a
//          cell of t (it reads the downslope Old values; this also keeps an upslope cell at most
//          one phase ahead of its receiver, so two qpub buffers suffice)  -- or, when phase p-1
//          was a saturated-zone phase, t and all Voronoi-neighbour tiles (their gathers read the
//          buffer this task overwrites);
//   SAT:   (p-1, t) and (p-1, v) for every Voronoi-neighbour tile v.
// Every task has a counter of unmet inputs; a finishing warp fences, decrements the counters of
// its dependents and pushes those that reach zero onto a ready queue, from which idle warps pop.
// No warp ever waits for a cell. Values never depend on the schedule (no atomics on data, partial
// sums are indexed by task). Cross-cell data is read from L2 (the library is compiled with -dlcm=cg).
//
// Across GPUs (tribs_part_t.nranks > 1; replaces tGraph::sendQpin / sendOverlap / sendNwt,
// src/tGraph/tGraph.cpp:2035-2155, 2431-2479, and tParallel's reductions): the same dataflow, one
// partition of the tiles per rank, all ranks inside one persistent launch each. A task whose results are
// read on another rank (the outflow of a reach outlet, the water table and front of a cell with a
// remote upslope cell or Voronoi neighbour, its water-table elevation and transmissivity before a
// saturated-zone phase) stores them straight into that rank's arrays over NVLink, fences at system
// scope and then releases the remote dependents: a system-scope atomic decrement of the peer's task
// counter and, if it reaches zero, a push onto the peer's ready queue. No host, no NCCL and no
// kernel boundary inside a batch: rank k works on step s while rank k+1 is still at step s-1.
// The last task of a tile in a step also leaves the tile's routing record (route_tile_record), so the
// only per-cell snapshot a batch keeps is the runoff the hillslope routing (route_seg_kernel) reads.
// ====================================================================================
struct PhaseInfo {
  int32_t kind;         // 0 UNSAT, 1 SAT, 2 ET (kernel (1): callEvapoPotential / callEvapoTrans before the step's sweep)
  int32_t step;         // step index inside the batch
  int32_t et_index;     // kind 2: which EtStepDev of the batch
  int32_t pad_;
  int32_t prev_sat;     // the phase before this one was a SAT phase
  int32_t next_sat;     // the next phase is a SAT phase: produce wt_elev / transm
  int32_t last_of_step; // last phase of its step: routing record + runoff snapshot
  int32_t has_next;     // another phase follows inside the batch
  long long tag;        // global phase number
  StepConst t;
  RouteStep rs;
};

// tile-level dependency graph (three CSRs over tiles, self excluded, entries distinct)
struct TileGraph {
  const int32_t *up_rp, *up;      // tiles holding upslope contributors
  const int32_t *down_rp, *down;  // tiles holding flow destinations
  const int32_t *nbr_rp, *nbr;    // tiles holding Voronoi neighbours
};

// what a rank knows about the others
struct PeerView {
  int32_t me, n_ranks;
  char *base[kMaxParts];        // exported memory of every rank as mapped here (base[me]: my own)
  const int32_t *tile_owner;    // [n_tiles]
  const int32_t *xp_rp, *xp_rank;   // per cell (device order): ranks that read this cell's values
  const uint8_t *tile_xp;       // [n_tiles] the tile has exports or remote dependents
};
template <class T>
__device__ __forceinline__ T *peer_ptr(const PeerView &pv, int r, T *mine) {
  return (T *)(pv.base[r] + ((char *)mine - pv.base[pv.me]));
}

__device__ __forceinline__ int ld_relaxed_i32(const int32_t *p) {
  int v;
  asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_i32(int32_t *p, int v) {
  asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_f64(double *p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void export_f64(const PeerView &pv, int i, double *mine, double v) {
  for (int k = pv.xp_rp[i]; k < pv.xp_rp[i + 1]; k++) st_relaxed_f64(peer_ptr(pv, pv.xp_rank[k], mine), v);
}

// Ready queues. Class 0 "hi": the tiles on the critical cycle of the batch (river cells and the cells next to the
// river, whose chain through all phases bounds the run time); classes 1-3 "lo": the bulk, one queue per kind of
// phase (sweep, saturated zone, ET) so that the warps of a block can start 16 tasks of the SAME kind together.
// ctl[2q] = head (consumer cursor), ctl[2q+1] = tail (producer cursor) of class q; class q's slots start at
// queue + q * n_tasks. Statistics live behind them (ctl + 16).
struct ReadyQueues { int32_t *ctl; int32_t *queue; const int32_t *tile_hi; int n_tasks; unsigned long long *trace; unsigned cap_hi_ns, cap_lo_ns; };

struct PipeArgs {
  const PhaseInfo *phases;
  TileGraph g;
  int32_t *cnt;
  ReadyQueues rq;
  const int32_t *up_rowptr, *up_col;
  double *qpub, *wt_elev, *transm;
  double *srf_steps;     // [steps][n_own], runoff of every owned cell at the end of each step
  double *tile_recs;     // [steps][n_own_tiles][rec_stride]
  double *tile_sums;     // [phases][n_own_tiles][4]
  double *psrf_slot;     // psrf of the last swept cell (written by its owner)
  double *uns_storage, *sat_storage;
  const double *evapotrans;
  double gwstep_h;
  int32_t n_pad, n_tiles, tile0, n_own_tiles, own_begin, n_own, K, rec_stride, last_sweep_dev;
  int32_t n_hi_tasks, n_lo_kind[3], n_hi_warps;
  int32_t mid_barrier;
  uint32_t patience_ns;  // how long thread 0 of a lo block waits for a partial chunk of ready tasks to fill up
  int32_t waves;         // a lo block claims up to 16 * waves ready tasks per round (the later waves start without a barrier)
  int32_t *abort_flag, *flags;
  const struct EtPack *et;         // kernel (1)'s constants, static arrays and fields (batches with ET phases)
  const EtStepDev *et_steps;       // one per ET phase of the batch
  PeerView pv;
};
struct EtPack { EtConst k; EtStatic st; EtView ev; SnowConst sc; SnowView sv; int32_t snow; int32_t pad_; };
// kernel (1) for one cell inside a batch: out of line, so that its 168 registers' worth of temporaries
// live in this call's frame instead of raising the register count of every sweep task. With OPTSNOW 1 (no canopy:
// tSnowPack::callSnowPack then needs nothing but the cell itself) the snow pack takes the place of both ET sweeps.
__device__ __noinline__ void et_phase_cell(const EtPack *et, const EtStepDev *s, const DynView *d, int i) {
  if (et->snow) {
    SnowConst sc = et->sc;
    sc.hourly = s->snow_hourly;
    SnowProbe pr;
    snow_cell_step(et->k, et->st, *s, sc, et->ev, et->sv, *d, i, 0.88, 0.0, pr);   // no canopy: nothing reads the carried members
  } else
    et_cell(et->k, et->st, *s, et->ev, *d, i);
}

__device__ __forceinline__ void fence_tasks(bool sys) {
  if (sys) __threadfence_system(); else __threadfence();
}
__device__ __forceinline__ void release_task(const PipeArgs &a, int p, int tile) {
  const int task = p * a.n_tiles + tile;
  const int q = a.rq.tile_hi[tile] ? 0 : 1 + a.phases[p].kind;
  const int r = a.pv.n_ranks > 1 ? a.pv.tile_owner[tile] : a.pv.me;
  if (r == a.pv.me) {
    if (atomicSub(a.cnt + task, 1) == 1) {
      fence_tasks(a.pv.n_ranks > 1);   // acquire the other producers' data before handing the task on
      const int slot = atomicAdd(a.rq.ctl + 2 * q + 1, 1);
      st_relaxed_i32(a.rq.queue + (size_t)q * a.rq.n_tasks + slot, task);
    }
  } else {   // the dependent lives on another rank: its counter and its ready queue, over NVLink
    if (atomicSub_system(peer_ptr(a.pv, r, a.cnt) + task, 1) == 1) {
      __threadfence_system();
      const int slot = atomicAdd_system(peer_ptr(a.pv, r, a.rq.ctl) + 2 * q + 1, 1);
      st_relaxed_i32(peer_ptr(a.pv, r, a.rq.queue) + (size_t)q * a.rq.n_tasks + slot, task);
    }
  }
}

// unmet-input counters of this rank's tasks and the initially ready ones
__global__ void __launch_bounds__(256)
pipeline_init_kernel(TileGraph g, const PhaseInfo *__restrict__ phases, int n_phases, int n_tiles, int tile0, int n_own_tiles,
                     int32_t *cnt, ReadyQueues rq) {
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (long long)n_phases * n_own_tiles) return;
  const int p = (int)(id / n_own_tiles), t = tile0 + (int)(id - (long long)p * n_own_tiles);
  const PhaseInfo &ph = phases[p];
  int c = 0;
  if (ph.kind == 0 || ph.kind == 2) {   // an ET phase waits like a sweep, without the same-phase inflow
    if (ph.kind == 0) c += g.up_rp[t + 1] - g.up_rp[t];
    if (p > 0) c += 1 + (ph.prev_sat ? g.nbr_rp[t + 1] - g.nbr_rp[t] : g.down_rp[t + 1] - g.down_rp[t]);
  } else if (p > 0)
    c += 1 + g.nbr_rp[t + 1] - g.nbr_rp[t];
  const int task = p * n_tiles + t;
  cnt[task] = c;
  if (c == 0) {
    const int q = rq.tile_hi[t] ? 0 : 1 + ph.kind;
    rq.queue[(size_t)q * rq.n_tasks + atomicAdd(rq.ctl + 2 * q + 1, 1)] = task;
  }
}

// All ranks meet here: nothing of the previous batch is still being read or written anywhere, and every
// rank's counters and queues of this batch are initialised before any peer can touch them.
__global__ void peer_barrier_kernel(PeerView pv, long long *flags, long long epoch, int32_t *abort_flag) {
  const int r = threadIdx.x;
  __threadfence_system();
  if (r < pv.n_ranks) {
    long long *dst = peer_ptr(pv, r, flags) + pv.me;
    asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(dst), "l"(epoch) : "memory");
    const unsigned long long t0 = global_ns();
    unsigned spins = 0;
    for (;;) {
      long long v;
      asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(flags + r) : "memory");
      if (v >= epoch) break;
      __nanosleep(256);
      if (((++spins) & 1023u) == 0u) {
        if (*(volatile int32_t *)abort_flag) break;
        if (global_ns() - t0 > kSpinTimeoutNs) { atomicExch(abort_flag, 2); break; }
      }
    }
  }
  __threadfence_system();
}

#ifndef TRIBS_PIPE_MIN_BLOCKS
#define TRIBS_PIPE_MIN_BLOCKS 4
#endif
#ifndef TRIBS_PIPE_THREADS
#define TRIBS_PIPE_THREADS 128
#endif
constexpr int kPipeThreads = TRIBS_PIPE_THREADS, kPipeWarps = kPipeThreads / 32;
template <int MAXC>
__global__ void __launch_bounds__(kPipeThreads, TRIBS_PIPE_MIN_BLOCKS)
pipeline_kernel(StaticView s, DynView d_even, DynView d_odd, ModelConst m, RouteConst rc, PipeArgs a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // Whole blocks serve either the "hi" queue or the three "lo" queues.
  //  hi: every warp on its own takes a ticket (atomicAdd on the head) and waits for exactly that slot to be
  //      filled, so a newly released river task is picked up at once by the warp that holds its ticket; the
  //      chain along the river through all phases is the critical path and never sits behind bulk work.
  //  lo: the 16 warps of a block start their tasks TOGETHER, and all of one kind of phase: thread 0 claims up to
  //      16 READY tasks of one class (compare-and-swap on the head, never past the tail, so a block never commits
  //      to work that does not exist yet), preferring a class that has 16 and the class of its last round. Tasks of
  //      one bucket and phase are released together and are neighbours in their queue, so the warps walk the same
  //      part of the (large, branchy) column code at about the same time and share its lines in the SM's 32 KB
  //      instruction cache, instead of each streaming its own part of ~80 KB of hot code (profiles/r02_block_tickets.md).
  const bool hi_warp = (int)blockIdx.x * kPipeWarps < a.n_hi_warps;
  __shared__ int sh_q, sh_base, sh_cnt;
  int last_q = 1;
  int my_q = 1, my_base = 0, my_next = 0, my_end = 0;      // this warp's remaining slots of the block's current round
  const unsigned ns_cap = hi_warp ? a.rq.cap_hi_ns : a.rq.cap_lo_ns;   // back-off ceiling of the queue poll
  const bool multi = a.pv.n_ranks > 1;
  long long t_wait = 0, t_busy = 0, n_done = 0, t_bar = 0;
  long long t_kind[3] = {0, 0, 0}, n_kind[3] = {0, 0, 0};
  long long t_mark = clock64();
  for (;;) {
    int task = -1;
    const int32_t *slot = nullptr;
    if (hi_warp) {
      if (lane == 0) {
        const int idx = atomicAdd(a.rq.ctl + 0, 1);
        if (idx < a.n_hi_tasks) slot = a.rq.queue + idx; else task = -2;
      }
    } else if (my_next < my_end) {   // a later wave of this block's round: no barrier, straight to the next slot
      if (lane == 0) slot = a.rq.queue + (size_t)my_q * a.rq.n_tasks + my_base + my_next;
      my_next += kPipeWarps;
    } else {
      const long long tb0 = clock64();
      __syncthreads();     // everybody is done with the previous round (and with sh_*)
      if (threadIdx.x == 0) {
        unsigned spins = 0, ns = 64;
        unsigned long long t_first = 0, t_seen = 0;
        int cnt = -1;
        for (;;) {
          bool all_done = true;
          int best = -1, best_n = 0;
          for (int k = 0; k < 3; k++) {
            const int q = 1 + ((last_q - 1 + k) % 3);      // the class of the last round first
            const int h = ld_relaxed_i32(a.rq.ctl + 2 * q), t = ld_relaxed_i32(a.rq.ctl + 2 * q + 1);
            if (h < a.n_lo_kind[q - 1]) all_done = false;
            const int n = min(t - h, kPipeWarps * a.waves);
            if (n > best_n) { best = q; best_n = n; }      // strict: ties go to the class tried first
          }
          if (best >= 0 && best_n < kPipeWarps && a.patience_ns) {
            // fewer ready tasks than warps: give the rest of their bucket a moment to arrive, so that they start
            // together on ONE block (and share its instruction cache) instead of one by one on sixteen
            const unsigned long long t_now = global_ns();
            if (!t_seen) t_seen = t_now;
            if (t_now - t_seen < a.patience_ns) { __nanosleep(64); continue; }
          }
          if (best >= 0) {
            const int h = ld_relaxed_i32(a.rq.ctl + 2 * best);
            const int n = min(ld_relaxed_i32(a.rq.ctl + 2 * best + 1) - h, kPipeWarps * a.waves);
            if (n > 0 && atomicCAS(a.rq.ctl + 2 * best, h, h + n) == h) { sh_q = best; sh_base = h; cnt = n; break; }
            continue;                                       // another block was faster: look again
          }
          if (all_done) break;
          __nanosleep(ns);
          if (ns < ns_cap) ns += ns;
          if (((++spins) & 63u) == 0u) {
            if (*(volatile int32_t *)a.abort_flag) break;
            const unsigned long long t_now = global_ns();
            if (!t_first) t_first = t_now;
            else if (t_now - t_first > kSpinTimeoutNs) { atomicExch(a.abort_flag, 1); break; }
          }
        }
        sh_cnt = cnt;
      }
      __syncthreads();
      t_bar += clock64() - tb0;
      const int cnt = sh_cnt;
      if (cnt < 0) break;                                   // uniform over the block: nothing is left (or abort)
      last_q = sh_q;
      my_q = sh_q; my_base = sh_base; my_end = cnt; my_next = warp + kPipeWarps;
      if (warp < cnt) { if (lane == 0) slot = a.rq.queue + (size_t)my_q * a.rq.n_tasks + my_base + warp; }
      else task = -3;                                       // fewer ready tasks than warps: this warp sits the round out
    }
    if (lane == 0 && slot) {
      unsigned spins = 0, ns = 32;
      unsigned long long t_first = 0;
      while ((task = ld_relaxed_i32(slot)) < 0) {   // hi: not released yet; lo: the tail is ahead of the slot's store by a moment
        __nanosleep(ns);
        if (ns < ns_cap) ns += ns;
        if (((++spins) & 255u) == 0u) {
          if (*(volatile int32_t *)a.abort_flag) { task = -2; break; }
          const unsigned long long t_now = global_ns();
          if (!t_first) t_first = t_now;
          else if (t_now - t_first > kSpinTimeoutNs) { atomicExch(a.abort_flag, 1); task = -2; break; }
        }
      }
    }
    task = __shfl_sync(0xffffffffu, task, 0);
    { const long long now_c = clock64(); t_wait += now_c - t_mark; t_mark = now_c; }
    if (task < 0) { if (hi_warp) break; else continue; }   // a lo warp without a task goes to the block's next round
    const int p = task / a.n_tiles, tile = task - p * a.n_tiles;
    if (a.rq.trace && lane == 0) a.rq.trace[2 * (size_t)task] = global_ns();   // TRIBS_TRACE: task timeline
    fence_tasks(multi);   // acquire side of the producers' fence -> counter -> queue hand-off
    const PhaseInfo &ph = a.phases[p];
    const int i = tile * kTile + lane;
    const long long tag = ph.tag;
    const int kind = ph.kind, next_sat = ph.next_sat;
    const DynView &d = (tag & 1) ? d_odd : d_even;   // Old = buffer tag%2, New = the other
    const bool act = s.boundary[i] >= 0 && s.owned[i];
    const bool xp = multi && a.pv.tile_xp[tile];     // some value of this tile is read on another rank
    double sums[4] = {0.0, 0.0, 0.0, 0.0};
    UnsatCellInputs u;
    double qpin = 0.0;
    if (kind == 0) {
      if (act) {
        unsat_cell_load(u, s, d, i, m, ph.t);
        const double *qsrc = a.qpub + (size_t)(tag & 1) * a.n_pad;
        const int k1 = a.up_rowptr[i + 1];
        for (int k = a.up_rowptr[i]; k < k1; k++) qpin += qsrc[a.up_col[k]];   // sweep order
        int state;
        const double qout = unsat_cell_advance(u, qpin, &state);
        double *qdst = a.qpub + (size_t)(tag & 1) * a.n_pad + i;
        *qdst = qout;
        if (xp) export_f64(a.pv, i, qdst, qout);
      }
      // the downslope tiles of this phase only need the new Qpout: release them before the
      // write-back and the diagnostics, which are off the critical chain along the river
      fence_tasks(xp);
      __syncwarp();
      const int r0 = a.g.down_rp[tile], r1 = a.g.down_rp[tile + 1];
      for (int k = r0 + lane; k < r1; k += 32) release_task(a, p, a.g.down[k]);
      // experiment (TRIBS_MID_BARRIER): re-align the round's warps before the write-back / diagnostics half
      if (a.mid_barrier && !hi_warp && my_end > 1 && my_end <= kPipeWarps)
        asm volatile("bar.sync 1, %0;" ::"r"(my_end * 32) : "memory");
    }
    if (kind == 2) {
      if (act) {
        // tRainfall::NewRain comes before SurfaceHydroProcesses in the reference's step: the ET sweep reads this step's rain
        if (ph.t.rain_mode != 0)
          d.f[TRIBS_F_Rain][i] = (ph.t.rain_mode == 1) ? ph.t.rain_uniform
                                 : ((ph.t.rain_mode == 2) ? ph.t.rain_cells[i] : ph.t.rain_cells[ph.t.rain_gauge[i]]);
        et_phase_cell(a.et, a.et_steps + ph.et_index, &d, i);
      }
    } else if (act) {
      if (kind == 0) {
        UnsatSums us;
        unsat_cell_store(u, d, i, qpin, m, ph.t, us);
        sums[0] = us.stok; sums[1] = us.tot_rain; sums[2] = us.tot_gw; sums[3] = us.tot_moist;
        if (i == a.last_sweep_dev) *a.psrf_slot = u.c.psrf;
        if (next_sat) {
          // ResetGW + the edge pre-pass for this cell: the post-sweep state is what SAT calls Old
          sat_prepare_cell(s, d.f[TRIBS_F_NwtNew], i, a.wt_elev, a.transm);
        }
      } else {
        int overflow = 0;
        const double gw = sat_gather<MAXC>(s, d, i, a.wt_elev, a.transm, &overflow);
        if (overflow) a.flags[0] = 1;
        SatSums ss;
        sat_cell(s, d, i, gw, m, ph.t, ss);
        sums[0] = ss.stok; sums[2] = ss.tot_gw; sums[3] = ss.tot_moist;
      }
      if (xp && a.pv.xp_rp[i + 1] > a.pv.xp_rp[i]) {
        // what other ranks read of this cell: as downslope cell of the next sweep (water table and front,
        // tGraph::sendOverlap) and as Voronoi neighbour of the next saturated-zone phase (sendNwt)
        export_f64(a.pv, i, d.f[TRIBS_F_NwtNew] + i, d.f[TRIBS_F_NwtNew][i]);
        export_f64(a.pv, i, d.f[TRIBS_F_NfNew] + i, d.f[TRIBS_F_NfNew][i]);
        if (kind == 0 && next_sat) {
          export_f64(a.pv, i, a.wt_elev + i, a.wt_elev[i]);
          export_f64(a.pv, i, a.transm + i, a.transm[i]);
        }
      }
    }
    const int tl = tile - a.tile0;
    if (ph.last_of_step) {
      // routing inputs of this step: the runoff snapshot for the hillslope routing and the tile's record
      if (act) a.srf_steps[(size_t)ph.step * a.n_own + (i - a.own_begin)] = d.f[TRIBS_F_srf][i];
      WaterBalanceArgs wb{a.uns_storage, a.sat_storage, a.evapotrans, a.gwstep_h, ph.rs.gw_time};
      route_tile_record(s, d, rc, m, wb, ph.rs, i, act, lane, a.K,
                        a.tile_recs + ((size_t)ph.step * a.n_own_tiles + tl) * a.rec_stride, a.flags);
    }
    double sa = warp_sum_fixed(sums[0]), sb = warp_sum_fixed(sums[1]);
    double sc = warp_sum_fixed(sums[2]), se = warp_sum_fixed(sums[3]);
    if (lane == 0) {
      double *ts = a.tile_sums + ((size_t)p * a.n_own_tiles + tl) * 4;
      ts[0] = sa; ts[1] = sb; ts[2] = sc; ts[3] = se;
    }
    // hand the results on: fence, then lanes release the dependents in parallel
    fence_tasks(xp);
    __syncwarp();
    if (ph.has_next) {
      const bool wide = (kind == 1) || next_sat;   // a SAT phase on either side: neighbours
      const int32_t *rp = wide ? a.g.nbr_rp : a.g.up_rp;
      const int32_t *lst = wide ? a.g.nbr : a.g.up;
      const int r0 = rp[tile], r1 = rp[tile + 1];
      for (int k = r0 + lane; k < r1; k += 32) release_task(a, p + 1, lst[k]);
      if (lane == 0) release_task(a, p + 1, tile);
    }
    if (a.rq.trace && lane == 0) a.rq.trace[2 * (size_t)task + 1] = global_ns();
    { const long long now_c = clock64(); t_busy += now_c - t_mark; t_kind[kind] += now_c - t_mark; n_kind[kind]++; t_mark = now_c; n_done++; }
  }
  if (lane == 0) {   // scheduler statistics: [0..2] lo warps, [3..5] hi warps
    unsigned long long *st = (unsigned long long *)(a.rq.ctl + 16) + (hi_warp ? 3 : 0);
    atomicAdd(st + 0, (unsigned long long)t_wait);
    atomicAdd(st + 1, (unsigned long long)t_busy);
    atomicAdd(st + 2, (unsigned long long)n_done);
    if (!hi_warp) {
      unsigned long long *sk = (unsigned long long *)(a.rq.ctl + 16) + 6;
      atomicAdd(sk, (unsigned long long)t_bar);
      for (int q = 0; q < 3; q++) { atomicAdd(sk + 1 + 2 * q, (unsigned long long)t_kind[q]); atomicAdd(sk + 2 + 2 * q, (unsigned long long)n_kind[q]); }
    }
  }
}

__global__ void __launch_bounds__(256)
esrf_fixup_kernel(StaticView s, DynView d, int begin, int end, double dtgw, const double *psrf_last) {
  int i = begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= end || s.boundary[i] < 0) return;
  double esrf = *psrf_last + d.f[TRIBS_F_esrf][i];
  esrf = esrf / s.cosv[i];
  d.f[TRIBS_F_esrf][i] = esrf * dtgw;
}
__global__ void set_global_kernel(double *globals, int which, const double *src) { globals[which] = *src; }
// per-step lateral inflows of all ranks (every stream slot is non-zero on its owner only)
__global__ void qeff_combine_kernel(PartVecs pv, size_t count, double *out) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  double tot = 0.0;
  for (int r = 0; r < pv.n; r++) tot += pv.v[r][k];
  out[k] = tot;
}
__global__ void scatter_rain_kernel(const double *__restrict__ src, const int32_t *__restrict__ dev_of_sweep, double *dst,
                                    int n, int own_begin, int own_end) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int dI = dev_of_sweep[i];
  if (dI >= own_begin && dI < own_end) dst[dI - own_begin] = src[i];
}

struct PipePlan {
  std::vector<uint8_t> pattern;   // per step: 1 = has a SAT phase
  std::vector<PhaseInfo> phases;  // tags/time constants are filled per run
  std::vector<StepBins> bins;
  std::vector<int32_t> is_sat;
};

struct PipeState {
  double *d_qpub = nullptr;        // [2][n_pad] new Qpout by phase parity
  double *d_srf = nullptr;         // [max_steps][n_own]
  double *d_recs = nullptr;        // [max_steps][n_own_tiles][rec_stride]
  double *d_tile_sums = nullptr;   // [2*max_steps][n_own_tiles][4]
  double *d_qeff_steps = nullptr;  // [max_steps][slots] this rank's slots (exported)
  double *d_qeff_out = nullptr;    // [max_steps][slots] all ranks
  double *d_rain = nullptr;        // [max_steps][n_own] per-cell forcing, device order
  double *d_gauge = nullptr;       // [max_steps][n_gauges]
  double *d_block_part = nullptr;  // [my parts][max_steps][kReduceBlocksMax][stride]
  double *d_part_vec = nullptr;    // [n_parts][max_steps][stride] (exported: peers read my rows)
  double *d_sums_part = nullptr;   // [n_parts][2*max_steps][4] (exported)
  double *d_psrf = nullptr;        // psrf of the last swept cell (exported)
  long long *d_bar = nullptr;      // [kMaxParts] barrier flags (exported)
  PhaseInfo *d_phases = nullptr;
  EtPack *d_etpack = nullptr;      // kernel (1) inside batches
  EtPack h_etpack{};
  EtStepDev *d_et_steps = nullptr; // [max_steps]
  double *d_met_batch = nullptr;   // [max_steps][2][9 * met_cap] station records of the batch's ET phases
  int met_batch_cap = 0;
  std::vector<EtStepDev> h_et_steps;
  StepBins *d_bins = nullptr;
  int32_t *d_is_sat = nullptr;
  int32_t *d_cnt = nullptr, *d_queue = nullptr, *d_headtail = nullptr;
  int32_t *d_tile_owner = nullptr, *d_xp_rp = nullptr, *d_xp_rank = nullptr;
  uint8_t *d_tile_xp = nullptr;
  TileGraph graph{};
  int max_steps = 0, last_steps = 0, gauge_cap = 0;
  long long g = 0;                 // phases completed so far
  long long bar_epoch = 0;
  int blocks = 0, n_sms = 0, my_hi_tiles = 0, share = 1;
  char *peer_base[kMaxParts] = {};
  void *ipc_opened[kMaxParts] = {};
  bool attached = false;
  std::vector<PipePlan> plans;
};
static PipeState *pipe_of(tribs_ctx *c) {
  if (!c->pipe) c->pipe = new PipeState();
  return c->pipe;
}

static int build_plan(PipeState *ps, const std::vector<uint8_t> &pattern, PipePlan **out) {
  for (auto &pl : ps->plans) if (pl.pattern == pattern) { *out = &pl; return 0; }
  if (ps->plans.size() >= 8) ps->plans.erase(ps->plans.begin());   // host memory only; oldest first
  PipePlan pl;
  pl.pattern = pattern;
  int n_et = 0;
  for (size_t sidx = 0; sidx < pattern.size(); sidx++) {
    if (pattern[sidx] & 2) {
      PhaseInfo e{};
      e.kind = 2; e.step = (int)sidx; e.et_index = n_et++;
      pl.phases.push_back(e);
    }
    PhaseInfo u{};
    u.kind = 0; u.step = (int)sidx; u.last_of_step = (pattern[sidx] & 1) ? 0 : 1; u.next_sat = pattern[sidx] & 1;
    pl.phases.push_back(u);
    if (pattern[sidx] & 1) {
      PhaseInfo v{};
      v.kind = 1; v.step = (int)sidx; v.last_of_step = 1;
      pl.phases.push_back(v);
    }
  }
  const int P = (int)pl.phases.size();
  for (int p = 0; p < P; p++) {
    pl.phases[p].prev_sat = p > 0 && pl.phases[p - 1].kind == 1;
    pl.phases[p].has_next = p + 1 < P;
    pl.is_sat.push_back(pl.phases[p].kind == 1 ? 1 : 0);
  }
  pl.bins.resize(pattern.size());
  ps->plans.push_back(pl);
  *out = &ps->plans.back();
  return 0;
}

// tile-level CSRs from the cell-level graph (host copies of the device arrays), tile owners and the
// per-cell export lists of a rank
static int build_tile_graph(tribs_ctx *c, PipeState *ps) {
  const int nt = c->n_tiles, n_pad = c->n_pad;
  std::vector<int32_t> h_up_rp(n_pad + 1), h_nbr_rp(n_pad + 1), h_fd(n_pad);
  CK(cudaMemcpy(h_up_rp.data(), c->d_up_rowptr, (n_pad + 1) * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h_nbr_rp.data(), c->sv.nbr_rowptr, (n_pad + 1) * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h_fd.data(), c->sv.flow_dest, n_pad * 4, cudaMemcpyDeviceToHost));
  std::vector<int32_t> h_up_col(h_up_rp[n_pad]), h_nbr_col(h_nbr_rp[n_pad]);
  if (!h_up_col.empty()) CK(cudaMemcpy(h_up_col.data(), c->d_up_col, h_up_col.size() * 4, cudaMemcpyDeviceToHost));
  if (!h_nbr_col.empty()) CK(cudaMemcpy(h_nbr_col.data(), c->sv.nbr_col, h_nbr_col.size() * 4, cudaMemcpyDeviceToHost));
  std::vector<int32_t> up_rp(nt + 1, 0), down_rp(nt + 1, 0), nbr_rp(nt + 1, 0), up, down, nbr, tmp;
  auto flush = [&](std::vector<int32_t> &dst, std::vector<int32_t> &rp, int t) {
    std::sort(tmp.begin(), tmp.end());
    tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
    for (int v : tmp) if (v != t) dst.push_back(v);
    rp[t + 1] = (int32_t)dst.size();
    tmp.clear();
  };
  for (int t = 0; t < nt; t++) {
    for (int l = 0; l < kTile; l++) { const int i = t * kTile + l; for (int k = h_up_rp[i]; k < h_up_rp[i + 1]; k++) tmp.push_back(h_up_col[k] / kTile); }
    flush(up, up_rp, t);
    for (int l = 0; l < kTile; l++) { const int i = t * kTile + l; if (h_fd[i] >= 0) tmp.push_back(h_fd[i] / kTile); }
    flush(down, down_rp, t);
    for (int l = 0; l < kTile; l++) {
      const int i = t * kTile + l;
      for (int k = h_nbr_rp[i]; k < h_nbr_rp[i + 1]; k++) if (h_nbr_col[k] >= 0) tmp.push_back(h_nbr_col[k] / kTile);
      if (h_fd[i] >= 0) tmp.push_back(h_fd[i] / kTile);
      for (int k = h_up_rp[i]; k < h_up_rp[i + 1]; k++) tmp.push_back(h_up_col[k] / kTile);
    }
    flush(nbr, nbr_rp, t);
  }
  // the neighbour relation must be symmetric for the counters to balance: symmetrise explicitly
  {
    std::vector<std::vector<int32_t>> adj(nt);
    for (int t = 0; t < nt; t++) for (int k = nbr_rp[t]; k < nbr_rp[t + 1]; k++) { adj[t].push_back(nbr[k]); adj[nbr[k]].push_back(t); }
    nbr.clear();
    for (int t = 0; t < nt; t++) {
      auto &a2 = adj[t];
      std::sort(a2.begin(), a2.end());
      a2.erase(std::unique(a2.begin(), a2.end()), a2.end());
      for (int v : a2) nbr.push_back(v);
      nbr_rp[t + 1] = (int32_t)nbr.size();
    }
  }
  int32_t *p1, *p2, *p3, *p4, *p5, *p6;
  if (dev_upload(c, &p1, up_rp) || dev_upload(c, &p2, up) || dev_upload(c, &p3, down_rp) || dev_upload(c, &p4, down) ||
      dev_upload(c, &p5, nbr_rp) || dev_upload(c, &p6, nbr))
    return 1;
  ps->graph = TileGraph{p1, p2, p3, p4, p5, p6};
  // hi tiles of this handle's range
  {
    std::vector<int32_t> hi(nt);
    CK(cudaMemcpy(hi.data(), c->d_tile_hi, (size_t)nt * 4, cudaMemcpyDeviceToHost));
    ps->my_hi_tiles = 0;
    for (int t = c->own_begin / kTile; t < c->own_end / kTile; t++) ps->my_hi_tiles += hi[t];
  }
  if (c->n_ranks > 1) {
    // owner of every tile, and for every owned cell the ranks that read it: the rank of its downslope
    // cell (Qpin), of its upslope cells (they read its Old water table and front) and of its Voronoi
    // neighbours (saturated-zone gather)
    std::vector<int32_t> towner(nt, 0);
    for (int r = 0; r < c->n_parts; r++)
      for (int t = c->part_begin[r] / kTile; t < c->part_end[r] / kTile; t++) towner[t] = r;
    auto owner_of = [&](int dI) { return towner[dI / kTile]; };
    std::vector<int32_t> xp_rp(n_pad + 1, 0), xp_rank;
    std::vector<uint8_t> tile_xp(nt, 0);
    for (int i = 0; i < n_pad; i++) {
      if (owner_of(i) == c->rank && c->sweep_of_dev[i] >= 0) {
        tmp.clear();
        if (h_fd[i] >= 0) tmp.push_back(owner_of(h_fd[i]));
        for (int k = h_up_rp[i]; k < h_up_rp[i + 1]; k++) tmp.push_back(owner_of(h_up_col[k]));
        for (int k = h_nbr_rp[i]; k < h_nbr_rp[i + 1]; k++) if (h_nbr_col[k] >= 0) tmp.push_back(owner_of(h_nbr_col[k]));
        std::sort(tmp.begin(), tmp.end());
        tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
        for (int r : tmp) if (r != c->rank) { xp_rank.push_back(r); tile_xp[i / kTile] = 1; }
      }
      xp_rp[i + 1] = (int32_t)xp_rank.size();
    }
    tmp.clear();
    for (int t = c->own_begin / kTile; t < c->own_end / kTile; t++) {   // tiles with a dependent on another rank
      for (int k = up_rp[t]; k < up_rp[t + 1]; k++) if (towner[up[k]] != c->rank) tile_xp[t] = 1;
      for (int k = down_rp[t]; k < down_rp[t + 1]; k++) if (towner[down[k]] != c->rank) tile_xp[t] = 1;
      for (int k = nbr_rp[t]; k < nbr_rp[t + 1]; k++) if (towner[nbr[k]] != c->rank) tile_xp[t] = 1;
    }
    if (dev_upload(c, &ps->d_tile_owner, towner) || dev_upload(c, &ps->d_xp_rp, xp_rp) || dev_upload(c, &ps->d_xp_rank, xp_rank) ||
        dev_upload(c, &ps->d_tile_xp, tile_xp))
      return 1;
  }
  return 0;
}

template <int MAXC>
static int pipe_occupancy(int *per_sm) {
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, pipeline_kernel<MAXC>, kPipeThreads, 0) != cudaSuccess;
}

static int ensure_pipe(tribs_ctx *c, PipeState *ps, int n_steps) {
  const int n_own = c->own_end - c->own_begin, own_tiles = n_own / kTile;
  const size_t stride = 5 * (size_t)c->route_window + kMeans;
  if (!ps->d_qpub) {
    if (shared_alloc(c, &ps->d_qpub, (size_t)2 * c->n_pad) || shared_alloc(c, &ps->d_headtail, 4096) ||
        shared_alloc(c, &ps->d_bar, (size_t)kMaxParts) || shared_alloc(c, &ps->d_psrf, (size_t)1))
      return 1;
    CK(cudaMemset(ps->d_qpub, 0, (size_t)2 * c->n_pad * sizeof(double)));
    CK(cudaMemset(ps->d_bar, 0, kMaxParts * sizeof(long long)));
    CK(cudaMemset(ps->d_psrf, 0, sizeof(double)));
    if (build_tile_graph(c, ps)) return 1;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, c->device));
    int per_sm = 0;
    if (c->max_degree <= 16 ? pipe_occupancy<16>(&per_sm) : (c->max_degree <= 32 ? pipe_occupancy<32>(&per_sm) : pipe_occupancy<64>(&per_sm)))
      return fail("tribs_gpu_run: occupancy query failed");
    if (per_sm < 1) return fail("tribs_gpu_run: pipeline_kernel does not fit on an SM");
    ps->blocks = per_sm * prop.multiProcessorCount;
    ps->n_sms = prop.multiProcessorCount;
    ps->peer_base[c->rank] = c->slab.base;
  }
  if (n_steps > ps->max_steps) {
    if (c->n_ranks > 1 && ps->max_steps > 0) return fail("tribs_gpu_run: more steps than tribs_part_t.max_batch_steps");
    // (re)allocate the per-step staging; the superseded buffers are released first
    CK(cudaStreamSynchronize(c->stream));
    auto drop = [&](void *q) {
      if (!q) return;
      auto it = std::find(c->allocs.begin(), c->allocs.end(), q);
      if (it != c->allocs.end()) { cudaFree(q); c->allocs.erase(it); }
    };
    drop(ps->d_srf); drop(ps->d_recs); drop(ps->d_tile_sums); drop(ps->d_qeff_out); drop(ps->d_rain); drop(ps->d_block_part);
    drop(ps->d_phases); drop(ps->d_bins); drop(ps->d_is_sat); drop(ps->d_et_steps); drop(ps->d_met_batch);
    ps->d_met_batch = nullptr; ps->met_batch_cap = 0;
    if (!c->slab.base) { drop(ps->d_qeff_steps); drop(ps->d_part_vec); drop(ps->d_sums_part); drop(ps->d_cnt); drop(ps->d_queue); }
    const size_t S = (size_t)n_steps, slots = (size_t)c->n_stream + 1;
    if (shared_alloc(c, &ps->d_cnt, 3 * S * c->n_tiles) || shared_alloc(c, &ps->d_queue, (size_t)kQueueClasses * 3 * S * c->n_tiles) ||
        shared_alloc(c, &ps->d_part_vec, (size_t)c->n_parts * S * stride) || shared_alloc(c, &ps->d_sums_part, (size_t)c->n_parts * 3 * S * 4) ||
        shared_alloc(c, &ps->d_qeff_steps, S * slots))
      return 1;
    if (dev_alloc(c, &ps->d_srf, S * n_own) || dev_alloc(c, &ps->d_recs, S * own_tiles * c->rec_stride) ||
        dev_alloc(c, &ps->d_tile_sums, 3 * S * own_tiles * 4) || dev_alloc(c, &ps->d_qeff_out, S * slots) ||
        dev_alloc(c, &ps->d_block_part, c->my_parts.size() * S * kReduceBlocksMax * stride) ||
        dev_alloc(c, &ps->d_phases, 3 * S) || dev_alloc(c, &ps->d_bins, S) || dev_alloc(c, &ps->d_is_sat, 3 * S) ||
        dev_alloc(c, &ps->d_et_steps, S) || (!ps->d_etpack && dev_alloc(c, &ps->d_etpack, 1)))
      return 1;
    ps->d_rain = nullptr;      // allocated when a step brings per-cell rain ...
    if (c->n_ranks > 1 && dev_alloc(c, &ps->d_rain, S * n_own)) return 1;   // ... but never inside a batch other ranks wait for
    ps->max_steps = n_steps;
  }
  return 0;
}

static int launch_barrier(tribs_ctx *c, PipeState *ps, const PeerView &pv) {
  ps->bar_epoch++;
  peer_barrier_kernel<<<1, 32, 0, c->stream>>>(pv, ps->d_bar, ps->bar_epoch, c->d_abort);
  c->launches++;
  return 0;
}

template <class T>
static const T *peer_host_ptr(tribs_ctx *c, PipeState *ps, int r, const T *mine) {
  if (c->n_ranks == 1 || r == c->rank) return mine;
  return (const T *)(ps->peer_base[r] + ((const char *)mine - c->slab.base));
}

// host-side part of tribs_gpu_et_step for one ET phase of a batch: the device form of the step and its station
// records, staged into this phase's slice of the batch's record buffer
static int stage_et_phase(tribs_ctx *c, PipeState *ps, int index, const tribs_et_step_t *s, int snow_hourly, EtStepDev *out) {
  EtState *e = c->et;
  const bool snow = e->snow_option != 0;
  if (s->hour < 0 || s->hour > 23) return fail("tribs_gpu_run_et: hour out of range");
  static const double rsRatio[24] = {3.837, 3.589, 3.21, 2.43, 1.617, 1.196, 1.067, 1.014,   // stomResist, tEvapoTrans.cpp:1494-1496
                                     0.995, 0.976, 0.976, 1.0, 1.053, 1.167, 1.354, 1.637,
                                     2.043, 2.66, 3.215, 3.507, 3.689, 3.818, 3.923, 4.024};
  EtStepDev d{};
  d.do_potential = s->do_potential && e->k.et_option; d.do_components = s->do_components; d.intercept_flag = s->intercept_flag;
  d.met_time = s->do_potential;
  if (snow) {   // as tribs_gpu_snow_step: callSnowPack runs at met_hour only and holds both sweeps
    d.do_potential = 1; d.do_components = 1; d.met_time = 1; d.intercept_flag = 0; d.snow_hourly = snow_hourly;
  }
  d.alphaD = s->alphaD; d.sinAlpha = s->sinAlpha; d.sunaz = s->sunaz; d.Io = s->Io;
  d.cos_alpha = cos(asin(s->sinAlpha));
  d.hour = s->hour; d.first_call = s->first_call; d.dew_hum_flag = s->dew_hum_flag; d.ts_option = s->ts_option;
  d.sky_flag_from = s->sky_flag_from; d.te_met = s->te_met; d.te_eti = s->te_eti;
  d.rs_ratio = rsRatio[s->hour];
  const tribs_met_t *mets[2] = {d.do_potential ? s->met_potential : nullptr,
                                (!snow && d.do_components && e->k.et_option) ? s->met_components : nullptr};
  for (int sl = 0; sl < 2; sl++) {
    const tribs_met_t *m = mets[sl];
    if (!m) continue;
    if (m->n_records <= 0) return fail("tribs_gpu_run_et: met record missing");
    const double *src[9] = {m->air_temp, m->dew_temp, m->rel_humid, m->vap_press, m->atm_press,
                            m->wind_speed, m->sky_cover, m->rad_global, m->elev};
    for (int q = 0; q < 9; q++) if (!src[q]) return fail("tribs_gpu_run_et: met series missing");
    const int nr = m->n_records;
    if (nr > ps->met_batch_cap) {
      if (ps->met_batch_cap > 0) return fail("tribs_gpu_run_et: the number of met records grew inside a run");
      if (dev_alloc(c, &ps->d_met_batch, (size_t)ps->max_steps * 2 * 9 * nr)) return 1;
      ps->met_batch_cap = nr;
    }
    if (m->cell_record) {
      for (int i = 0; i < c->n; i++)
        if (m->cell_record[i] < 0 || m->cell_record[i] >= nr) return fail("tribs_gpu_run_et: cell_record out of range");
      e->h_rec = to_device_order<int32_t>(c, m->cell_record, 0);
      CK(cudaMemcpyAsync(e->d_record[sl], e->h_rec.data(), e->h_rec.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
      CK(cudaStreamSynchronize(c->stream));        // h_rec is reused by the next call
      e->have_record[sl] = true;
    } else if (!e->have_record[sl]) return fail("tribs_gpu_run_et: cell_record not given yet");
    std::vector<double> pack((size_t)9 * nr);
    for (int q = 0; q < 9; q++) memcpy(pack.data() + (size_t)q * nr, src[q], (size_t)nr * sizeof(double));
    double *b = ps->d_met_batch + ((size_t)index * 2 + sl) * 9 * ps->met_batch_cap;
    CK(cudaMemcpyAsync(b, pack.data(), pack.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));   // pageable: staged before the call returns
    MetDev &o = sl == 0 ? d.pot : d.cmp;
    o.cell_record = e->d_record[sl];
    o.air_temp = b; o.dew_temp = b + nr; o.rel_humid = b + 2 * (size_t)nr; o.vap_press = b + 3 * (size_t)nr;
    o.atm_press = b + 4 * (size_t)nr; o.wind_speed = b + 5 * (size_t)nr; o.sky_cover = b + 6 * (size_t)nr;
    o.rad_global = b + 7 * (size_t)nr; o.elev = b + 8 * (size_t)nr;
  }
  if (snow) d.cmp = d.pot;        // ComputeETComponents runs inside the same cell loop, before the met clock advances
  *out = d;
  return 0;
}

static int run_batch(tribs_handle_t c, int n_steps, const tribs_step_t *steps, const uint32_t *masks,
                     const tribs_et_step_t *const *et_steps, const int32_t *snow_hourly);
extern "C" int tribs_gpu_run(tribs_handle_t c, int n_steps, const tribs_step_t *steps, const uint32_t *masks) {
  return run_batch(c, n_steps, steps, masks, nullptr, nullptr);
}
// The same with kernel (1) inside the batch: et_steps[k] (or NULL) is what tribs_gpu_et_step would be called with
// before the sweep of step k (Simulator::SurfaceHydroProcesses, tSimul.cpp:419-461). The ET / interception update
// of a cell only needs that cell, so it is one more phase of the tile's dataflow: no launch boundary, no
// basin-wide wait at every METSTEP.
extern "C" int tribs_gpu_run_et(tribs_handle_t c, int n_steps, const tribs_step_t *steps, const uint32_t *masks,
                                const tribs_et_step_t *const *et_steps) {
  return run_batch(c, n_steps, steps, masks, et_steps, nullptr);
}
// The same with OPTSNOW 1 and no canopy (OPTINTERCEPT 0): snow_steps[k] (or NULL) and hourly_time_step[k] are what
// tribs_gpu_snow_step would be called with before the sweep of step k (tSimul.cpp:464-492). With a canopy the snow pack
// carries two members from cell to cell in sweep order and stays a call between batches.
extern "C" int tribs_gpu_run_snow(tribs_handle_t c, int n_steps, const tribs_step_t *steps, const uint32_t *masks,
                                  const tribs_et_step_t *const *snow_steps, const int32_t *hourly_time_step) {
  if (!hourly_time_step) return fail("tribs_gpu_run_snow: hourly_time_step missing");
  return run_batch(c, n_steps, steps, masks, snow_steps, hourly_time_step);
}

static int run_batch(tribs_handle_t c, int n_steps, const tribs_step_t *steps, const uint32_t *masks,
                     const tribs_et_step_t *const *et_steps, const int32_t *snow_hourly) {
  if (!c || !steps || !masks || n_steps <= 0) return fail("tribs_gpu_run: bad argument");
  CK(cudaSetDevice(c->device));
  const bool dbg_t = getenv("TRIBS_DEBUG") != nullptr;
  const auto t_run0 = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (dbg_t) fprintf(stderr, "[tribs] run host %-22s %8.3f ms\n", what,
                       1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t_run0).count());
  };
  if (c->partitioned) return fail("tribs_gpu_run: this handle was partitioned with tribs_gpu_set_partition (host-exchanged single steps); "
                                  "pipelined batches over partitions need tribs_part_t at tribs_gpu_init");
  if (c->n_late > 0) return fail("tribs_gpu_run: the sweep order is not topological (late contributors); use tribs_gpu_step");
  if (c->rc.flowexp > 0.0) return fail("tribs_gpu_run: flowexp > 0 needs the channel router between steps; use tribs_gpu_step");
  const uint32_t need = TRIBS_PHASE_UNSAT | TRIBS_PHASE_ROUTE | TRIBS_PHASE_RESET;
  std::vector<uint8_t> pattern(n_steps);
  for (int k = 0; k < n_steps; k++) {
    if ((masks[k] & need) != need) return fail("tribs_gpu_run: every step must contain UNSAT, ROUTE and RESET");
    pattern[k] = (masks[k] & TRIBS_PHASE_SAT) ? 1 : 0;
    if (et_steps && et_steps[k] && (et_steps[k]->do_potential || et_steps[k]->do_components)) {
      if (!c->et) return fail("tribs_gpu_run_et: call tribs_gpu_et_init first");
      if (c->et->snow_option && (!c->et->snow_ready || c->et->k.i_option != 0 || !snow_hourly))
        return fail("tribs_gpu_run_et: with OPTSNOW 1 use tribs_gpu_run_snow (no canopy), or tribs_gpu_snow_step between batches (canopy)");
      if (!c->et->snow_option && snow_hourly) return fail("tribs_gpu_run_snow: the handle has no snow pack");
      pattern[k] |= 2;
    }
  }
  PipeState *ps = pipe_of(c);
  if (c->n_ranks > 1 && !ps->attached) return fail("tribs_gpu_run: nranks > 1 needs tribs_gpu_attach_peers / tribs_gpu_attach_local first");
  if (ensure_pipe(c, ps, n_steps)) return 1;
  PipePlan *pl = nullptr;
  if (build_plan(ps, pattern, &pl)) return 1;
  cudaStream_t st = c->stream;
  const int n_pad = c->n_pad, P = (int)pl->phases.size();
  const int n_own = c->own_end - c->own_begin, own_tiles = n_own / kTile, tile0 = c->own_begin / kTile;
  const size_t stride = 5 * (size_t)c->route_window + kMeans;
  const int slots = c->n_stream + 1, window = c->route_window;
  if (!c->at_boundary) return fail("tribs_gpu_run: call it on a step boundary (after RESET)");
  // the pipeline works on fixed even/odd views; phase tag T reads buffer T%2
  DynView d_even = c->dv, d_odd = c->dv;
  {
    const int olds[7] = {TRIBS_F_NwtOld, TRIBS_F_MuOld, TRIBS_F_MiOld, TRIBS_F_NtOld, TRIBS_F_NfOld, TRIBS_F_RuOld, TRIBS_F_RiOld};
    const int news[7] = {TRIBS_F_NwtNew, TRIBS_F_MuNew, TRIBS_F_MiNew, TRIBS_F_NtNew, TRIBS_F_NfNew, TRIBS_F_RuNew, TRIBS_F_RiNew};
    DynView &second = ((ps->g + 1) & 1) ? d_even : d_odd;   // the first phase (tag g+1) reads the handle's Old
    for (int k = 0; k < 7; k++) std::swap(second.f[olds[k]], second.f[news[k]]);
  }
  const double hv0 = c->rc.velcoef / c->rc.velratio;
  int bin_lo = 0x7fffffff, bin_hi = -0x7fffffff;
  std::vector<StepConst> step_const(n_steps);
  for (int k = 0; k < n_steps; k++) {
    const tribs_step_t &sp = steps[k];
    StepConst t;
    t.dt = sp.dt; t.dt_gw = sp.dt_gw; t.now = sp.current_time; t.elapsed = (double)sp.elapsed_steps;
    t.hourly_reset = sp.hourly_reset; t.psrf_last = 0.0; t.defer_esrf = 1;
    t.rain_mode = 0;
    if (sp.rain_mode == TRIBS_RAIN_UNIFORM) { t.rain_mode = 1; t.rain_uniform = sp.rain_uniform; }
    else if (sp.rain_mode == TRIBS_RAIN_PER_CELL) {
      if (!sp.rain) return fail("tribs_gpu_run: rain_mode PER_CELL without a rain array");
      if (!ps->d_rain && dev_alloc(c, &ps->d_rain, (size_t)ps->max_steps * n_own)) return 1;
      double *dst = ps->d_rain + (size_t)k * n_own;
      CK(cudaMemcpyAsync(c->d_stage, sp.rain, (size_t)c->n * sizeof(double), cudaMemcpyHostToDevice, st));
      scatter_rain_kernel<<<(c->n + 255) / 256, 256, 0, st>>>(c->d_stage, c->d_dev_of_sweep, dst, c->n, c->own_begin, c->own_end);
      c->launches++;
      t.rain_mode = 2; t.rain_cells = dst - c->own_begin;
    } else if (sp.rain_mode == TRIBS_RAIN_GAUGES) {
      if (!sp.rain || !c->d_gauge_of_cell) return fail("tribs_gpu_run: rain_mode GAUGES needs tribs_gpu_set_rain_gauges and the gauge values");
      if (ps->gauge_cap < ps->max_steps * c->n_gauges) {
        if (dev_alloc(c, &ps->d_gauge, (size_t)ps->max_steps * c->n_gauges)) return 1;
        ps->gauge_cap = ps->max_steps * c->n_gauges;
      }
      double *dst = ps->d_gauge + (size_t)k * c->n_gauges;
      CK(cudaMemcpyAsync(dst, sp.rain, (size_t)c->n_gauges * sizeof(double), cudaMemcpyHostToDevice, st));
      t.rain_mode = 3; t.rain_cells = dst; t.rain_gauge = c->d_gauge_of_cell;
    }
    step_const[k] = t;
  }
  int n_state = 0, n_et = 0;
  ps->h_et_steps.clear();
  for (int p = 0; p < P; p++) {
    PhaseInfo &ph = pl->phases[p];
    const tribs_step_t &sp = steps[ph.step];
    ph.tag = ps->g + 1 + n_state;      // an ET phase carries the tag of the sweep that follows it
    if (ph.kind != 2) n_state++;
    else {
      EtStepDev d{};
      if (stage_et_phase(c, ps, n_et, et_steps[ph.step], snow_hourly ? snow_hourly[ph.step] : 0, &d)) return 1;
      ps->h_et_steps.push_back(d);
      n_et++;
    }
    ph.t = step_const[ph.step];
    ph.rs = RouteStep{sp.current_time, sp.qstrm_outlet, hv0, (int32_t)(fmod(sp.current_time, c->gwstep_h) == 0.0),
                      (int32_t)((pattern[ph.step] & 1) ? 0 : 1)};
    if (ph.last_of_step) {
      pl->bins[ph.step] = step_bins(c, sp.current_time);
      bin_lo = std::min(bin_lo, pl->bins[ph.step].bin_base);
      bin_hi = std::max(bin_hi, pl->bins[ph.step].bin_base + window);
    }
  }
  if (n_et > 0) {
    ps->h_etpack.k = c->et->k; ps->h_etpack.st = c->et->st; ps->h_etpack.ev = c->et->ev;
    ps->h_etpack.snow = c->et->snow_option != 0;
    if (ps->h_etpack.snow) { ps->h_etpack.sc = c->et->sc; ps->h_etpack.sv = c->et->snv; }
    CK(cudaMemcpyAsync(ps->d_etpack, &ps->h_etpack, sizeof(EtPack), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ps->d_et_steps, ps->h_et_steps.data(), (size_t)n_et * sizeof(EtStepDev), cudaMemcpyHostToDevice, st));
    c->et->gen += n_et;
  }
  lap("forcing staged");
  CK(cudaMemcpyAsync(ps->d_phases, pl->phases.data(), (size_t)P * sizeof(PhaseInfo), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ps->d_bins, pl->bins.data(), (size_t)n_steps * sizeof(StepBins), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ps->d_is_sat, pl->is_sat.data(), (size_t)P * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  const int n_tasks = P * c->n_tiles;
  CK(cudaMemsetAsync(ps->d_headtail, 0, 64 * sizeof(int32_t), st));
  CK(cudaMemsetAsync(ps->d_queue, 0xFF, (size_t)kQueueClasses * n_tasks * sizeof(int32_t), st));
  PeerView pv{};
  pv.me = c->rank; pv.n_ranks = c->n_ranks;
  for (int r = 0; r < c->n_ranks; r++) pv.base[r] = ps->peer_base[r];
  pv.tile_owner = ps->d_tile_owner; pv.xp_rp = ps->d_xp_rp; pv.xp_rank = ps->d_xp_rank; pv.tile_xp = ps->d_tile_xp;
  {
    PhaseTimer pt(c, 0);
    ReadyQueues rq{ps->d_headtail, ps->d_queue, c->d_tile_hi, n_tasks, nullptr, 128u, 2048u};   // poll ceilings: a hi warp on its slot, thread 0 of an idle lo block
    if (const char *e = getenv("TRIBS_POLL_CAP_NS")) rq.cap_lo_ns = (unsigned)std::max(32, atoi(e));
    if (const char *e = getenv("TRIBS_POLL_CAP_HI_NS")) rq.cap_hi_ns = (unsigned)std::max(32, atoi(e));
    const char *trace_path = c->n_ranks == 1 ? getenv("TRIBS_TRACE") : nullptr;   // debugging aid: start/end time of every task of this launch
    unsigned long long *d_trace = nullptr;
    if (trace_path) {
      CK(cudaMalloc((void **)&d_trace, (size_t)2 * n_tasks * sizeof(unsigned long long)));
      CK(cudaMemsetAsync(d_trace, 0, (size_t)2 * n_tasks * sizeof(unsigned long long), st));
      rq.trace = d_trace;
    }
    // blocks of the persistent grid: all an SM holds, or this handle's share of them when several handles
    // of one process run on the same device (tribs_gpu_attach_local)
    const int blocks = std::max(2, ps->blocks / std::max(1, ps->share));    // at least one hi and one lo block
    // warps serving the "hi" queue: one per SM for a single river; on a dendritic network the stream
    // tiles are a sizeable share of all tasks and get that share of the warps
    const int total_warps = blocks * kPipeWarps;
    // whole blocks: the hi tasks are about as long as the others but lie on the chain every hillslope waits for;
    // measured on the 1M-cell valley (5 % hi tiles, 2368 warps): 160 warps 697 M, 256: 768 M, 400: 747 M, 600: 692 M
    int n_hi_warps = std::max(std::min(total_warps / 2, ps->n_sms),
                              (int)std::lround(2.2 * total_warps * (double)ps->my_hi_tiles / std::max(1, own_tiles)));
    n_hi_warps = ((n_hi_warps + kPipeWarps - 1) / kPipeWarps) * kPipeWarps;
    n_hi_warps = std::max(kPipeWarps, std::min(n_hi_warps, total_warps - kPipeWarps));
    if (const char *e = getenv("TRIBS_HI_WARPS")) n_hi_warps = std::max(1, std::min(total_warps - 1, atoi(e)));
    const long long n_init = (long long)P * own_tiles;
    pipeline_init_kernel<<<(unsigned)((n_init + 255) / 256), 256, 0, st>>>(ps->graph, ps->d_phases, P, c->n_tiles, tile0, own_tiles,
                                                                           ps->d_cnt, rq);
    c->launches++;
    if (c->n_ranks > 1 && launch_barrier(c, ps, pv)) return 1;
    PipeArgs a{};
    a.phases = ps->d_phases; a.g = ps->graph; a.cnt = ps->d_cnt; a.rq = rq;
    a.up_rowptr = c->d_up_rowptr; a.up_col = c->d_up_col;
    a.qpub = ps->d_qpub; a.wt_elev = c->d_wt_elev; a.transm = c->d_transm;
    a.srf_steps = ps->d_srf; a.tile_recs = ps->d_recs; a.tile_sums = ps->d_tile_sums; a.psrf_slot = ps->d_psrf;
    a.uns_storage = c->dv.f[TRIBS_F_UnSaturatedStorage]; a.sat_storage = c->dv.f[TRIBS_F_SaturatedStorage];
    a.evapotrans = c->et ? c->et->ev.f[TRIBS_ET_EvapoTrans] : nullptr;
    a.gwstep_h = c->gwstep_h;
    a.n_pad = n_pad; a.n_tiles = c->n_tiles; a.tile0 = tile0; a.n_own_tiles = own_tiles; a.own_begin = c->own_begin; a.n_own = n_own;
    a.K = c->tile_bins; a.rec_stride = c->rec_stride; a.last_sweep_dev = c->last_sweep_dev;
    {
      int n_kind[3] = {0, 0, 0};
      for (int q = 0; q < P; q++) n_kind[pl->phases[q].kind]++;
      a.n_hi_tasks = P * ps->my_hi_tiles;
      for (int q = 0; q < 3; q++) a.n_lo_kind[q] = n_kind[q] * (own_tiles - ps->my_hi_tiles);
      a.n_hi_warps = n_hi_warps;
      a.waves = 1;
      a.mid_barrier = getenv("TRIBS_MID_BARRIER") != nullptr;
      a.patience_ns = 0;
      if (const char *e = getenv("TRIBS_PATIENCE_NS")) a.patience_ns = (uint32_t)std::max(0, atoi(e));
      if (const char *e = getenv("TRIBS_ROUND_WAVES")) a.waves = std::max(1, std::min(8, atoi(e)));
    }
    a.abort_flag = c->d_abort; a.flags = c->d_flags; a.pv = pv;
    a.et = ps->d_etpack; a.et_steps = ps->d_et_steps;
    void *args[] = {&c->sv, &d_even, &d_odd, &c->mc, &c->rc, &a};
    const void *fn = c->max_degree <= 16 ? (const void *)pipeline_kernel<16>
                     : (c->max_degree <= 32 ? (const void *)pipeline_kernel<32> : (const void *)pipeline_kernel<64>);
    CK(cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(kPipeThreads), args, 0, st));
    c->launches++;
    lap("batch kernel launched");
    if (d_trace) {   // file: int32 n_phases, n_tiles; per tile int32 hi flag, int32 first cell (sweep index); 2 x u64 per task
      std::vector<unsigned long long> h((size_t)2 * n_tasks);
      CK(cudaMemcpyAsync(h.data(), d_trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      CK(cudaFree(d_trace));
      std::vector<int32_t> hi(c->n_tiles), first(c->n_tiles), kind(P);
      CK(cudaMemcpy(hi.data(), c->d_tile_hi, hi.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
      for (int t = 0; t < c->n_tiles; t++) first[t] = c->sweep_of_dev[(size_t)t * kTile];
      for (int q = 0; q < P; q++) kind[q] = pl->phases[q].kind;
      if (FILE *f = fopen(trace_path, "wb")) {
        int32_t hdr[2] = {P, c->n_tiles};
        fwrite(hdr, sizeof(int32_t), 2, f);
        fwrite(kind.data(), sizeof(int32_t), kind.size(), f);
        fwrite(hi.data(), sizeof(int32_t), hi.size(), f);
        fwrite(first.data(), sizeof(int32_t), first.size(), f);
        fwrite(h.data(), sizeof(unsigned long long), h.size(), f);
        fclose(f);
      }
    }
    // basin accumulators of this handle's partitions, per phase
    if (reduce_sums_local(c, ps->d_tile_sums, (size_t)own_tiles * 4, tile0, P, ps->d_sums_part) ) return 1;
  }
  if (getenv("TRIBS_DEBUG") && c->n_ranks == 1) {
    unsigned long long h[13];
    CK(cudaMemcpyAsync(h, ps->d_headtail + 16, sizeof(h), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    fprintf(stderr, "[tribs] rank %d batch: phases=%d tiles=%d hi_tiles=%d | lo warps: wait %.1f%% busy/task %.1f us tasks %llu (block barrier %.1f%%) | hi warps: wait %.1f%% busy/task %.1f us tasks %llu\n",
            c->rank, P, own_tiles, ps->my_hi_tiles, 100.0 * h[0] / std::max(1.0, (double)(h[0] + h[1])), h[1] / std::max(1.0, (double)h[2]) / 1.9e3, h[2], 100.0 * h[6] / std::max(1.0, (double)(h[0] + h[1])),
            100.0 * h[3] / std::max(1.0, (double)(h[3] + h[4])), h[4] / std::max(1.0, (double)h[5]) / 1.9e3, h[5]);
    fprintf(stderr, "[tribs]   lo tasks by kind: sweep %.1f us x %llu | saturated zone %.1f us x %llu | ET %.1f us x %llu\n",
            h[7] / std::max(1.0, (double)h[8]) / 1.9e3, h[8], h[9] / std::max(1.0, (double)h[10]) / 1.9e3, h[10],
            h[11] / std::max(1.0, (double)h[12]) / 1.9e3, h[12]);
  }
  ps->g += n_state;
  // the handle's views after the batch: New = what the last phase wrote
  c->dv = ((ps->g) & 1) ? d_odd : d_even;
  c->new_valid = true;
  for (int k = 0; k < n_steps; k++) if (pattern[k] & 1) c->gw_dirty = true;
  {
    PhaseTimer pt(c, 2);
    // hillslope routing of every step from its runoff snapshot, in step order
    for (int k = 0; k < n_steps; k++)
      if (route_segments(c, c->dv, ps->d_srf + (size_t)k * n_own - c->own_begin, steps[k].current_time, steps[k].qstrm_outlet,
                         ps->d_qeff_steps + (size_t)k * slots))
        return 1;
    // tile records -> this handle's partition vectors
    for (size_t q = 0; q < c->my_parts.size(); q++) {
      const int r = c->my_parts[q], nt = part_tiles(c, r), G = reduce_blocks(nt);
      double *bp = ps->d_block_part + q * (size_t)ps->max_steps * kReduceBlocksMax * stride;
      tile_reduce_kernel<<<dim3(G, n_steps), kRouteWarps * 32, kRouteWarps * stride * sizeof(double), st>>>(
          ps->d_recs, (size_t)own_tiles * c->rec_stride, (c->part_begin[r] - c->own_begin) / kTile, nt, c->tile_bins, window, ps->d_bins, bp,
          c->d_flags);
      const long long warps = (long long)n_steps * stride;
      part_reduce_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(bp, G, window, n_steps,
                                                                              ps->d_part_vec + (size_t)r * ps->max_steps * stride);
      c->launches += 2;
    }
    // every rank's partials are complete: add the partitions in partition order
    if (c->n_ranks > 1 && launch_barrier(c, ps, pv)) return 1;
    PartVecs pvs{}, pvv{}, pvq{};
    pvs.n = pvv.n = c->n_parts;
    for (int r = 0; r < c->n_parts; r++) {
      pvs.v[r] = peer_host_ptr(c, ps, r, ps->d_sums_part + (size_t)r * P * 4);
      pvv.v[r] = peer_host_ptr(c, ps, r, ps->d_part_vec + (size_t)r * ps->max_steps * stride);
    }
    combine_sums_kernel<<<1, 32, 0, st>>>(pvs, ps->d_is_sat, P, c->d_globals);
    const int n_bins = bin_hi - bin_lo;
    combine_results_kernel<<<(5 * n_bins + kMeans + 255) / 256, 256, 0, st>>>(pvv, ps->d_bins, n_steps, window, c->rc.res_limit, bin_lo,
                                                                              n_bins, c->d_results, c->d_globals);
    c->launches += 2;
    const double *qeff_all = ps->d_qeff_steps;
    if (c->n_ranks > 1) {
      pvq.n = c->n_ranks;
      for (int r = 0; r < c->n_ranks; r++) pvq.v[r] = peer_host_ptr(c, ps, r, ps->d_qeff_steps);
      const size_t count = (size_t)n_steps * slots;
      qeff_combine_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(pvq, count, ps->d_qeff_out);
      c->launches++;
      qeff_all = ps->d_qeff_out;
    } else
      CK(cudaMemcpyAsync(ps->d_qeff_out, ps->d_qeff_steps, (size_t)n_steps * slots * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(c->d_qeff_now, qeff_all + (size_t)(n_steps - 1) * slots, (size_t)slots * sizeof(double),
                       cudaMemcpyDeviceToDevice, st));
    // psrf of the last swept cell (its owner wrote it) and the constant hillslope velocity
    int last_owner = c->rank;
    if (c->n_ranks > 1) last_owner = c->sweep_owner[c->n - 1];
    const double *psrf_src = peer_host_ptr(c, ps, last_owner, ps->d_psrf);
    set_global_kernel<<<1, 1, 0, st>>>(c->d_globals, TRIBS_G_psrf_last, psrf_src);
    c->launches++;
    CK(cudaMemcpyAsync(c->d_globals + TRIBS_G_hillvel_last, &hv0, sizeof(double), cudaMemcpyHostToDevice, st));
    if (pattern[n_steps - 1] & 1) {
      esrf_fixup_kernel<<<(n_own + 255) / 256, 256, 0, st>>>(c->sv, c->dv, c->own_begin, c->own_end, steps[n_steps - 1].dt_gw,
                                                             c->d_globals + TRIBS_G_psrf_last);
      c->launches++;
    }
  }
  ps->last_steps = n_steps;
  {  // RESET of the last step
    PhaseTimer pt(c, 3);
    swap_old_new(c);
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
  lap("all launched");
  return 0;
}

extern "C" int tribs_gpu_retrieve_qeff_steps(tribs_handle_t c, int n_steps, double *out) {
  if (!c || !out) return fail("tribs_gpu_retrieve_qeff_steps: null argument");
  PipeState *ps = pipe_of(c);
  if (n_steps > ps->last_steps) return fail("tribs_gpu_retrieve_qeff_steps: more steps than the last tribs_gpu_run held");
  CK(cudaSetDevice(c->device));
  if (check_sweep_abort(c)) return 1;
  CK(cudaMemcpyAsync(out, ps->d_qeff_out, (size_t)n_steps * (c->n_stream + 1) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

// ------------------------------------------------------------------------------------ peers
// CUDA loads a kernel's code at its first launch, and that load can wait for running kernels. A rank
// whose barrier kernel spins on the device while its host thread is stuck loading the next kernel cannot
// return to the caller, which - with several ranks in one process - is the one who would launch the peer the
// barrier waits for. Every kernel a batch launches is therefore loaded before the first batch.
static int preload_batch_kernels() {
  cudaFuncAttributes a;
#define TRIBS_PRELOAD(k) CK(cudaFuncGetAttributes(&a, k))
  TRIBS_PRELOAD(pipeline_init_kernel); TRIBS_PRELOAD(peer_barrier_kernel);
  TRIBS_PRELOAD(pipeline_kernel<16>); TRIBS_PRELOAD(pipeline_kernel<32>); TRIBS_PRELOAD(pipeline_kernel<64>);
  TRIBS_PRELOAD(sums_reduce_kernel); TRIBS_PRELOAD(combine_sums_kernel); TRIBS_PRELOAD(route_seg_kernel);
  TRIBS_PRELOAD(route_seg_fixup_kernel); TRIBS_PRELOAD(route_seg_level2_kernel); TRIBS_PRELOAD(route_latch_kernel); TRIBS_PRELOAD(tile_reduce_kernel);
  TRIBS_PRELOAD(part_reduce_kernel); TRIBS_PRELOAD(combine_results_kernel); TRIBS_PRELOAD(qeff_combine_kernel);
  TRIBS_PRELOAD(set_global_kernel); TRIBS_PRELOAD(esrf_fixup_kernel); TRIBS_PRELOAD(scatter_rain_kernel);
  TRIBS_PRELOAD(fill_kernel); TRIBS_PRELOAD(gauge_fill_kernel); TRIBS_PRELOAD(gather_to_sweep_kernel);
  TRIBS_PRELOAD(scatter_from_sweep_kernel);
#undef TRIBS_PRELOAD
  return 0;
}
// the driver's own memset / copy kernels are loaded by using them once, with the sizes and patterns a batch uses
// (the handle is on a step boundary: the arrays zeroed here are zero there anyway)
static int warm_driver_kernels(tribs_ctx *c, PipeState *ps) {
  cudaStream_t st = c->stream;
  const size_t slots = (size_t)c->n_stream + 1;
  CK(cudaMemsetAsync(ps->d_headtail, 0, 64 * sizeof(int32_t), st));
  CK(cudaMemsetAsync(ps->d_queue, 0xFF, (size_t)kQueueClasses * 3 * ps->max_steps * c->n_tiles * sizeof(int32_t), st));
  CK(cudaMemsetAsync(ps->d_queue, 0xFF, 64, st));
  if (c->at_boundary) {
    const int zero7[7] = {TRIBS_F_Qpin, TRIBS_F_srf, TRIBS_F_sbsrf, TRIBS_F_hsrf, TRIBS_F_psrf, TRIBS_F_satsrf, TRIBS_F_gwchange};
    for (int k = 0; k < 7; k++) CK(cudaMemsetAsync(c->dv.f[zero7[k]], 0, (size_t)c->n_pad * sizeof(double), st));
  }
  CK(cudaMemcpyAsync(ps->d_qeff_out, ps->d_qeff_steps, (size_t)ps->max_steps * slots * sizeof(double), cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(c->d_qeff_now, ps->d_qeff_out, slots * sizeof(double), cudaMemcpyDeviceToDevice, st));
  std::vector<double> host((size_t)std::max(c->n, 4096), 0.0);      // pageable uploads, small and large
  CK(cudaMemcpyAsync(c->d_stage, host.data(), (size_t)c->n * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(c->d_stage, host.data(), 4096, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(c->d_stage, host.data(), sizeof(double), cudaMemcpyHostToDevice, st));
  if (ps->d_rain) scatter_rain_kernel<<<1, 32, 0, st>>>(c->d_stage, c->d_dev_of_sweep, ps->d_rain, 0, 0, 0);
  // The first launch of a kernel with a larger stack frame than any before makes the driver grow the context's
  // local-memory arena, which waits for every running kernel: the batch kernel is therefore launched once here with
  // nothing to do (every ticket counter is already at its total), not for the first time behind a barrier kernel
  // that waits for a peer.
  {
    PipeArgs a{};
    a.rq = ReadyQueues{ps->d_headtail, ps->d_queue, c->d_tile_hi, 1, nullptr, 128u, 2048u};
    a.n_hi_warps = kPipeWarps; a.waves = 1;
    a.abort_flag = c->d_abort; a.flags = c->d_flags;
    a.pv.me = c->rank; a.pv.n_ranks = 1;
    DynView dv = c->dv;
    void *args[] = {&c->sv, &dv, &dv, &c->mc, &c->rc, &a};
    const void *fn = c->max_degree <= 16 ? (const void *)pipeline_kernel<16>
                     : (c->max_degree <= 32 ? (const void *)pipeline_kernel<32> : (const void *)pipeline_kernel<64>);
    CK(cudaLaunchCooperativeKernel(fn, dim3(2), dim3(kPipeThreads), args, 0, st));
    peer_barrier_kernel<<<1, 32, 0, st>>>(PeerView{c->rank, 0, {}, nullptr, nullptr, nullptr, nullptr}, ps->d_bar, 0, c->d_abort);
  }
  CK(cudaStreamSynchronize(st));
  CK(cudaMemset(c->d_qeff_now, 0, slots * sizeof(double)));
  CK(cudaMemset(ps->d_headtail, 0, 64 * sizeof(int32_t)));
  return 0;
}

extern "C" int tribs_gpu_owned_range(tribs_handle_t c, int32_t *begin, int32_t *end) {
  if (!c || !begin || !end) return fail("tribs_gpu_owned_range: null argument");
  *begin = c->own_begin; *end = c->own_end;
  return 0;
}
extern "C" int tribs_gpu_ipc_export(tribs_handle_t c, void *handle_out) {
  if (!c || !handle_out) return fail("tribs_gpu_ipc_export: null argument");
  if (c->n_ranks < 2 || !c->slab.base) return fail("tribs_gpu_ipc_export: the handle was not initialised with nranks > 1");
  static_assert(sizeof(cudaIpcMemHandle_t) == TRIBS_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
  CK(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, c->slab.base));
  memcpy(handle_out, &h, sizeof(h));
  return 0;
}
extern "C" int tribs_gpu_attach_peers(tribs_handle_t c, const void *handles) {
  if (!c || !handles) return fail("tribs_gpu_attach_peers: null argument");
  if (c->n_ranks < 2) return fail("tribs_gpu_attach_peers: the handle was not initialised with nranks > 1");
  CK(cudaSetDevice(c->device));
  PipeState *ps = pipe_of(c);
  if (ps->attached) return fail("tribs_gpu_attach_peers: already attached");
  for (int r = 0; r < c->n_ranks; r++) {
    if (r == c->rank) { ps->peer_base[r] = c->slab.base; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)handles + (size_t)r * TRIBS_IPC_HANDLE_BYTES, sizeof(h));
    void *q = nullptr;
    CK(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
    ps->ipc_opened[r] = q;
    ps->peer_base[r] = (char *)q;
  }
  if (preload_batch_kernels() || warm_driver_kernels(c, ps)) return 1;
  ps->attached = true;
  return 0;
}
extern "C" int tribs_gpu_attach_local(const tribs_handle_t *handles, int nranks) {
  if (!handles || nranks < 2) return fail("tribs_gpu_attach_local: bad argument");
  for (int r = 0; r < nranks; r++) {
    tribs_ctx *c = handles[r];
    if (!c || c->n_ranks != nranks || c->rank != r || !c->slab.base)
      return fail("tribs_gpu_attach_local: handles[r] must be the handle of rank r of an nranks-rank partition");
  }
  for (int r = 0; r < nranks; r++) {
    tribs_ctx *c = handles[r];
    PipeState *ps = pipe_of(c);
    int share = 0;
    for (int q = 0; q < nranks; q++) {
      ps->peer_base[q] = handles[q]->slab.base;
      if (handles[q]->device == c->device) share++;
      else {
        CK(cudaSetDevice(c->device));
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, c->device, handles[q]->device));
        if (!can) return fail("tribs_gpu_attach_local: the devices cannot address each other's memory");
        cudaError_t e = cudaDeviceEnablePeerAccess(handles[q]->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
        cudaGetLastError();
      }
    }
    ps->share = share;
    CK(cudaSetDevice(c->device));
    if (preload_batch_kernels() || warm_driver_kernels(c, ps)) return 1;
    ps->attached = true;
  }
  return 0;
}

extern "C" int tribs_gpu_set_rain_gauges(tribs_handle_t c, int32_t n_gauges, const int32_t *cell_gauge) {
  if (!c || !cell_gauge || n_gauges <= 0) return fail("tribs_gpu_set_rain_gauges: bad argument");
  CK(cudaSetDevice(c->device));
  std::vector<int32_t> g((size_t)c->n_pad, -1);
  for (int i = 0; i < c->n; i++) {
    if (cell_gauge[i] < 0 || cell_gauge[i] >= n_gauges) return fail("tribs_gpu_set_rain_gauges: gauge index out of range");
    g[c->dev_of_sweep[i]] = cell_gauge[i];
  }
  CK(cudaStreamSynchronize(c->stream));
  if (dev_upload(c, &c->d_gauge_of_cell, g) || dev_alloc(c, &c->d_gauge_vals, (size_t)n_gauges)) return 1;
  c->n_gauges = n_gauges;
  if (c->pipe) c->pipe->gauge_cap = 0;
  return 0;
}

static void tribs_free_pipe(tribs_ctx *c) {
  if (c->pipe)
    for (int r = 0; r < kMaxParts; r++)
      if (c->pipe->ipc_opened[r]) cudaIpcCloseMemHandle(c->pipe->ipc_opened[r]);
  delete c->pipe;   // its device buffers live in the handle's allocation list
  c->pipe = nullptr;
}
