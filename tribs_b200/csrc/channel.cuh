// Kinematic-wave channel routing on the device: what tKinemat::RunRoutingModel does after the
// hillslope routing of a step (src/tFlowNet/tKinemat.cpp:803-897), for OPTPERCOLATION 0 / OPTRESERVOIR 0:
//   InitializeStreamReach 1136-1252, ComputeCoefficientArrays 1267-1285 (constant per run: done once on the host),
//   AssignLateralInflux 1296-1494, AssignQin 1561-1573, SolveForTwoNodeReach 1736-1765, KinematWave 1788-1931
//   (ComputeFunction 1950-1998, ComputeJacobian 2052-2072, SolveLinearSystem 2085-2145 = bandec / banbks 2154-2230,
//   lnsrch 2321-2412), ComputeQout 1582-1586, UpdateStreamVars 1617-1672.
//
// The lateral inflows of every step of a batch are already on the device (PipeState::d_qeff_out, [step][stream slot]),
// and with FLOWEXP 0 nothing flows back from the channels to the hillslopes, so the router is a consumer of the batch:
// ONE launch routes all steps of the batch over all reaches, and only the outlet discharge per step travels to the host
// instead of a lateral inflow per stream node and step.
//
// Layout: one 1024-thread block. The work is a chain (steps x reaches in list order x Newton iterations x a pivoted
// banded LU), a few thousand unknowns per reach: everything element-wise (function values with their x^(5/3), the
// Jacobian, the trial levels of the line search, maxima) is spread over the block; the pivoted LU of the Newton step is cut
// into chunks eliminated side by side on the lanes of one warp plus a small interface system (channel_solve.cuh:
// the reference's single chain of N rows costs ~420 cycles per row on this GPU); the sums of the line search (|F|^2, p.p,
// grad.p) are added up by one warp each in a fixed order (chan_warp_sum). FP64 throughout; differs from the reference through
// the library's pow() and through the order of operations of the linear solve and of those sums (1e-15 relative).
#pragma once

#include "channel_solve.cuh"

constexpr int kChanThreads = 1024;
constexpr int kChanSharedVecs = 6;         // band rows (3), right-hand sides y / v / w of the partition solve (3)
constexpr int kChanInterface = 2 * kChanMaxParts * (2 * kChanMaxParts + 1) + 2 * kChanMaxParts;   // interface matrix + unknowns
constexpr int kChanSharedRows = (226 * 1024 / 8 - kChanInterface - 256) / kChanSharedVecs;       // rows (+2) per vector that fit

struct ChanNet {
  int n_reach, max_n, slots;               // slots = stream slots + the basin outlet
  const int32_t *off, *slot;               // per reach: entry range; per entry: stream slot
  const double *bis, *siis, *C, *Y1, *Y2, *Y3;   // per entry (Y*: indexed from the reach's first entry)
  double roughness;
  double *hlevel, *qstrm, *velocity;       // per slot
  double *outlet_hlev;                     // per reach (tKinemat::OutletHlev)
  double *ws;                              // workspace: kChanWsVecs vectors of (max_n + 2)
  int32_t *status;                         // [0] first error, [1] Newton iterations, [2] backtracks, [3] MAXITS exceeded
  long long *cycles;                       // thread 0's clock64() spent in [0] band solve, [1] line-search sums, [2] function sums, [3] whole kernel
};
constexpr int kChanWsVecs = 16;

__device__ __forceinline__ double chan_block_max(double v, double *red) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); w++) r = fmax(r, red[w]);
  return r;
}

// A sum over 0..m-1 by one warp in a fixed order: every lane adds up its contiguous chunk, the 32 partial sums are then
// added in lane order. Not the reference's left-to-right order (the sums only feed the step-length decisions of the line
// search), but the same on every run. Call with all 32 lanes of a warp; every lane receives the total.
template <typename Term>
__device__ __forceinline__ double chan_warp_sum(int m, Term term) {
  const int lane = threadIdx.x & 31, per = (m + 31) / 32, lo = lane * per, hi = min(m, lo + per);
  double s = 0.0;
  for (int i = lo; i < hi; i++) s += term(i);
  double tot = 0.0;
  for (int l = 0; l < 32; l++) tot += __shfl_sync(0xffffffffu, s, l);
  return tot;
}

// tKinemat.cpp:1950-1998 for element i; X = levels of nodes 1..m at (t+1) and XP = X^(5/3) (the reference's XFK2 array),
// H = levels of nodes 0..m at t
__device__ __forceinline__ double chan_function(int i, int m, const double *X, const double *XP, const double *H, const double *c,
                                                const double *y1, const double *y2, const double *y3, const double *reff,
                                                double qit, double hupp) {
  const int m1 = m - 1;
  if (i == 0)
    return 0.5 * c[2] * XP[1] - qit + y3[0] * (hupp - H[0]) + y2[0] * (X[0] - H[1]) + y1[0] * (X[1] - H[2]) - reff[0];
  if (i == m1)
    return 0.5 * c[m] * XP[m1] - 0.5 * c[m1] * XP[m1 - 1] + y3[m1] * (X[m1 - 1] - H[m1]) + y2[m1] * (X[m1] - H[m]) - reff[m1];
  return 0.5 * c[i + 2] * XP[i + 1] - 0.5 * c[i] * XP[i - 1] + y3[i] * (X[i - 1] - H[i]) + y2[i] * (X[i] - H[i + 1])
         + y1[i] * (X[i + 1] - H[i + 2]) - reff[i];
}

// mathutil.cpp:1682-1759 (rtsafe on polynomialH), thread 0 only
__device__ double chan_two_node_root(double c1, double c2, double c3, double x1, double x2, double xacc, double xguess) {
  const double fk2 = 5. / 3., fk3 = 2. / 3.;
  auto poly = [&](double x, double &fn, double &df) { fn = c1 * pow(x, fk2) + c2 * x + c3; df = fk2 * c1 * pow(x, fk3) + c2; };
  double df, dx, dxold, f, fh, fl, temp, xh, xl, rts;
  poly(x1, fl, df);
  poly(x2, fh, df);
  if (fl == 0.0) return x1;
  if (fh == 0.0) return x2;
  if (fl < 0.0) { xl = x1; xh = x2; } else { xh = x1; xl = x2; }
  rts = xguess;
  dxold = fabs(x2 - x1);
  dx = dxold;
  poly(rts, f, df);
  for (int j = 1; j <= 200; j++) {
    if ((((rts - xh) * df - f) * ((rts - xl) * df - f) > 0.0) || (fabs(2.0 * f) > fabs(dxold * df))) {
      dxold = dx; dx = 0.5 * (xh - xl); rts = xl + dx;
      if (xl == rts) return rts;
    } else {
      dxold = dx; dx = f / df; temp = rts; rts -= dx;
      if (temp == rts) return rts;
    }
    if (fabs(dx) < xacc) return rts;
    poly(rts, f, df);
    if (f < 0.0) xl = rts; else xh = rts;
  }
  return 0.0;
}

__global__ void __launch_bounds__(kChanThreads, 1)
channel_route_kernel(ChanNet cn, const double *qeff_steps, int n_steps, double *qout, int shared_rows) {
  extern __shared__ double chan_sm[];
  __shared__ double red[kChanThreads / 32];
  __shared__ double sc[8];          // 0 f, 1 max|F|, 2 stpmax, 3 slope, 6 level of a two-node reach, 7 discharge at the reach outlet
  const int tid = threadIdx.x, T = blockDim.x;
  const size_t stride = (size_t)cn.max_n + 2;
  double *X = cn.ws, *X1 = X + stride, *Fg = X1 + stride, *aa = Fg + stride, *bb = aa + stride, *cc = bb + stride,
         *reis = cc + stride, *his = reis + stride, *gA1 = his + stride, *gA2 = gA1 + stride, *gA3 = gA2 + stride,
         *gB = gA3 + stride, *XP = gB + stride,     // XP: X^(5/3) for the function, X^(2/3) for the Jacobian
         *gV = XP + stride, *gW = gV + stride;
  const bool in_shared = (int)stride <= shared_rows;
  double *IM = chan_sm, *IX = IM + 2 * kChanMaxParts * (2 * kChanMaxParts + 1);       // interface system of the partition solve
  double *vec = chan_sm + kChanInterface;
  double *A1 = in_shared ? vec : gA1, *A2 = in_shared ? vec + stride : gA2, *A3 = in_shared ? vec + 2 * stride : gA3,
         *B = in_shared ? vec + 3 * stride : gB, *V = in_shared ? vec + 4 * stride : gV, *W = in_shared ? vec + 5 * stride : gW;
  double *G = A1, *FS = A2;         // after the solve: gradient and function values where the warp sums read them
  const double rif = cn.roughness;
  const double TOLF = 1.0e-4, TOLX = 1.0e-7, STPMX = 100.0, ALF = 1.0e-4;
  const double fk2 = 5. / 3., fk3 = 2. / 3., fk1 = 5. / 6.;
  int iters = 0, backtracks = 0, maxits = 0;
  long long cy_solve = 0, cy_ls = 0, cy_fs = 0;
  const long long cy_begin = clock64();

  for (int step = 0; step < n_steps; step++) {
    const double *qeff = qeff_steps + (size_t)step * cn.slots;
    double q_out = 0.0;
    for (int id = tid; id < cn.n_reach; id += T) cn.qstrm[cn.slot[cn.off[id + 1] - 1]] = 0.0;   // tKinemat.cpp:823-824
    __syncthreads();
    for (int id = 0; id < cn.n_reach; id++) {
      const int e0 = cn.off[id], n = cn.off[id + 1] - e0, m = n - 1, m1 = n - 2;
      const int32_t *sl = cn.slot + e0;
      const double *bis = cn.bis + e0, *siis = cn.siis + e0, *c = cn.C + e0, *y1 = cn.Y1 + e0, *y2 = cn.Y2 + e0, *y3 = cn.Y3 + e0;
      const bool basin_outlet = sl[m] == cn.slots - 1;
      for (int i = tid; i < n; i += T) {
        his[i] = i < m ? cn.hlevel[sl[i]] : cn.outlet_hlev[id];
        if (i < m) {                                               // AssignLateralInflux
          const double q = qeff[sl[i]];
          if (i < m1) reis[i] = (q + qeff[sl[i + 1]]) / 2.;
          else reis[i] = (basin_outlet ? q + qeff[cn.slots - 1] : q) / 2.;
        } else
          reis[i] = 0.0;
      }
      const double qin = cn.qstrm[sl[0]];                          // AssignQin
      const double qit = 0.5 * qin;
      const double hupp = pow(qin * rif / (bis[0] * sqrt(siis[0])), 3. / 5.);
      __syncthreads();
      if (n == 2) {                                                // SolveForTwoNodeReach
        if (tid == 0) {
          const double c1 = 0.5 * c[1], c2 = 0.5 * y2[0];
          const double c3 = 2 * y3[0] * (hupp - his[0]) - c2 * his[1] - qit - reis[0];
          sc[6] = c3 > 0.0 ? his[1] : chan_two_node_root(c1, c2, c3, 0., 40., 1.0E-6, hupp);
        }
        __syncthreads();
        if (tid == 0) { his[0] = hupp; his[1] = sc[6]; }
        __syncthreads();
      } else {                                                     // KinematWave
        for (int i = tid; i < m; i += T) { const double x = his[i + 1]; X[i] = X1[i] = x; A3[i] = x; XP[i] = pow(x, fk2); }
        __syncthreads();
        double fmx = 0.0;
        for (int i = tid; i < m; i += T) {
          const double f = chan_function(i, m, X1, XP, his, c, y1, y2, y3, reis, qit, hupp);
          Fg[i] = f; FS[i] = f; fmx = fmax(fmx, fabs(f));
        }
        const double ftest0 = chan_block_max(fmx, red);          // (barriers inside: FS and A3 are complete behind it)
        if (tid < 32) {
          const double sum = chan_warp_sum(m, [&](int i) { const double v = FS[i]; return v * v; });
          if (tid == 0) sc[0] = 0.5 * sum;
        } else if (tid < 64) {
          const double sx = chan_warp_sum(m, [&](int i) { const double v = A3[i]; return v * v; });
          if (tid == 32) sc[2] = STPMX * fmax(sqrt(sx), (double)m);
        }
        __syncthreads();
        double f = sc[0];
        const double stpmax = sc[2];
        bool done = ftest0 < 0.01 * TOLF;                          // the old levels already solve the system: nothing is written
        bool converged_levels = false;
        int iter = 1;
        for (; !done && iter <= 200; iter++) {
          iters++;
          // Jacobian (global copy for the gradient) and, in the same pass, the band matrix and right-hand side of the solve
          for (int i = tid; i < m; i += T) XP[i] = pow(X1[i], fk3);
          __syncthreads();
          for (int i = tid; i < m; i += T) {
            double b_i, c_i = 0.0, a_i = 0.0;
            if (i < m - 1) { b_i = y2[i]; c_i = fk1 * c[i + 2] * XP[i + 1] + y1[i]; }
            else b_i = fk1 * c[m] * XP[i] + y2[m - 1];
            if (i >= 1) a_i = -fk1 * c[i] * XP[i - 1] + y3[i];
            aa[i] = a_i; bb[i] = b_i; cc[i] = c_i;
            const int k = i + 1;                                    // 1-based row: sub-diagonal, diagonal, super-diagonal
            A1[k] = a_i; A2[k] = b_i; A3[k] = c_i;
            const double fv = Fg[i];
            B[k] = (fv == 0) ? fv : -fv;
          }
          __syncthreads();
          // partition solve (channel_solve.cuh): chunks side by side on the lanes of warp 0, the interface system on
          // thread 0, the assembly on everybody
          const int P = chan_part_count(m);
          const long long t0 = clock64();
          if (tid < P) {
            int first, last;
            chan_part_range(m, P, tid, &first, &last);
            chan_part_chunk(A1 + 1, A2 + 1, A3 + 1, A1, A2, A3, B, V, W, first, last, tid > 0, tid < P - 1);
          }
          __syncthreads();
          if (tid == 0) chan_part_interface(B, V, W, m, P, IM, IX);
          __syncthreads();
          for (int k = 1 + tid; k <= m; k += T) B[k] = chan_part_assemble(B, V, W, IX, m, P, k);
          __syncthreads();
          if (tid == 0) cy_solve += clock64() - t0;
          const double *p = B + 1;
          for (int i = tid; i < m; i += T) {                        // grad(f) = F*J, and the levels of the preceding iteration
            double g;
            if (i == 0) g = bb[0] * Fg[0] + aa[1] * Fg[1];
            else if (i == m1) g = bb[m1] * Fg[m1] + cc[m1 - 1] * Fg[m1 - 1];
            else g = cc[i - 1] * Fg[i - 1] + bb[i] * Fg[i] + aa[i + 1] * Fg[i + 1];
            G[i] = g;
            X[i] = X1[i];
          }
          __syncthreads();
          const double fold = f;
          // lnsrch
          double tloc = 0.0;
          // the two sums are independent: one warp each, side by side (a scaled step - never seen on the test basins -
          // redoes the slope after the scaling, as the reference's order of operations requires)
          if (tid < 32) {
            const long long t0 = clock64();
            const double sum = chan_warp_sum(m, [&](int i) { const double v = p[i]; return v * v; });
            if (tid == 0) { sc[4] = sqrt(sum); cy_ls += clock64() - t0; }
          } else if (tid < 64) {
            const double slope = chan_warp_sum(m, [&](int i) { return G[i] * p[i]; });
            if (tid == 32) sc[3] = slope;
          }
          __syncthreads();
          if (sc[4] > stpmax) {                                     // scale the step, then the slope of the scaled step
            const double scale = stpmax / sc[4];
            __syncthreads();
            for (int i = tid; i < m; i += T) B[i + 1] *= scale;
            __syncthreads();
            if (tid < 32) {
              const double slope = chan_warp_sum(m, [&](int i) { return G[i] * p[i]; });
              if (tid == 0) sc[3] = slope;
            }
            __syncthreads();
          }
          for (int i = tid; i < m; i += T) tloc = fmax(tloc, fabs(p[i]) / fmax(fabs(X[i]), 1.0));
          const double ptest = chan_block_max(tloc, red);
          const double slope = sc[3];
          if (slope >= 0.0) {                                       // the reference exits the process here
            if (tid == 0 && cn.status[0] == 0) cn.status[0] = 1;
            done = true;
            break;
          }
          const double alamin = TOLX / ptest;
          double alam = 1.0, alam2 = 0.0, f2 = 0.0;
          int check = 0;
          for (;;) {
            for (int i = tid; i < m; i += T) {
              double x = X[i] + alam * p[i];
              if (x < 0.0) x = 1.0E-6;
              X1[i] = x;
              XP[i] = pow(x, fk2);
            }
            __syncthreads();
            double tmx = 0.0;
            for (int i = tid; i < m; i += T) {
              const double fv = chan_function(i, m, X1, XP, his, c, y1, y2, y3, reis, qit, hupp);
              Fg[i] = fv; FS[i] = fv; tmx = fmax(tmx, fabs(fv));
            }
            const double trial_max = chan_block_max(tmx, red);
            if (tid < 32) {
              const long long t0 = clock64();
              const double sum = chan_warp_sum(m, [&](int i) { const double v = FS[i]; return v * v; });
              if (tid == 0) { sc[0] = 0.5 * sum; sc[1] = trial_max; cy_fs += clock64() - t0; }
            }
            __syncthreads();
            f = sc[0];
            if (alam < alamin) {
              for (int i = tid; i < m; i += T) X1[i] = X[i];
              check = 1;
              break;
            } else if (f <= fold + ALF * alam * slope)
              break;
            double tmplam;
            if (alam == 1.0)
              tmplam = -slope / (2.0 * (f - fold - slope));
            else {
              const double rhs1 = f - fold - alam * slope, rhs2 = f2 - fold - alam2 * slope;
              const double a = (rhs1 / (alam * alam) - rhs2 / (alam2 * alam2)) / (alam - alam2);
              const double b = (-alam2 * rhs1 / (alam * alam) + alam * rhs2 / (alam2 * alam2)) / (alam - alam2);
              if (a == 0.0) tmplam = -slope / (2.0 * b);
              else {
                const double disc = b * b - 3.0 * a * slope;
                if (disc < 0.0) tmplam = 0.5 * alam;
                else if (b <= 0.0) tmplam = (-b + sqrt(disc)) / (3.0 * a);
                else tmplam = -slope / (b + sqrt(disc));
              }
              if (tmplam > 0.5 * alam) tmplam = 0.5 * alam;
            }
            alam2 = alam; f2 = f;
            alam = fmax(tmplam, 0.1 * alam);
            backtracks++;
            __syncthreads();                                        // sc[] is rewritten by the next trial
          }
          __syncthreads();
          const double ftest = sc[1];                               // max |F| at the accepted (or restored) levels
          if (ftest < TOLF) { converged_levels = true; break; }
          if (check) { converged_levels = true; break; }            // gradient test only sets the returned flag
          double dloc = 0.0;
          for (int i = tid; i < m; i += T) dloc = fmax(dloc, fabs(X1[i] - X[i]) / fmax(fabs(X1[i]), 1.0));
          if (chan_block_max(dloc, red) < TOLX) { converged_levels = true; break; }
        }
        if (!done && !converged_levels && iter > 200) maxits++;     // "MAXITS exceeded in newt": levels stay
        __syncthreads();
        if (converged_levels) {                                     // UpdateHsShifted
          for (int i = tid; i < m; i += T) his[i + 1] = X1[i];
          if (tid == 0) his[0] = hupp;
        }
        __syncthreads();
      }
      // ComputeQout, UpdateStreamVars
      for (int i = tid; i < m; i += T) {
        const double h = his[i];
        cn.hlevel[sl[i]] = h;
        cn.qstrm[sl[i]] = (bis[i] * pow(h, (5. / 3.)) * sqrt(siis[i]) / rif);
        cn.velocity[sl[i]] = (pow(h, (2. / 3.)) * sqrt(siis[i]) / rif);
      }
      __syncthreads();                                              // the outlet's discharge is added after the nodes' values are set
      if (tid == 0) {
        const double h = his[m];
        if (basin_outlet) cn.hlevel[sl[m]] = h;
        cn.outlet_hlev[id] = h;
        const double q = bis[m] * pow(h, fk2) * sqrt(siis[m]) / rif;
        cn.qstrm[sl[m]] += q;
        sc[7] = q;
      }
      __syncthreads();
      q_out = sc[7];
    }
    if (tid == 0) qout[step] = q_out;
    __syncthreads();
  }
  if (tid == 0) {
    cn.status[1] += iters; cn.status[2] += backtracks; cn.status[3] += maxits;
    cn.cycles[0] += cy_solve; cn.cycles[1] += cy_ls; cn.cycles[2] += cy_fs; cn.cycles[3] += clock64() - cy_begin;
  }
}

// ------------------------------------------------------------------------------------ host side
struct ChannelState {
  ChanNet net{};
  int n_entries = 0;
  double *d_qout = nullptr;
  int qout_cap = 0;
  size_t smem = 0;
  int shared_rows = 0;
  int64_t launches = 0;
  double *snap = nullptr;                  // tribs_gpu_snapshot: 3 x slots + n_reach
};

static ChannelState *&channel_of(tribs_ctx *c) { return c->chan; }
static void tribs_free_channel(tribs_ctx *c) { delete c->chan; c->chan = nullptr; }   // device memory is in c->allocs

// ---- the channel state in tribs_gpu_snapshot / rollback and in the checkpoint file (checkpoint.cuh)
static bool channel_present(tribs_ctx *c) { return c->chan != nullptr; }
static int channel_copy_state(tribs_ctx *c, bool to_snapshot) {
  ChannelState *cs = c->chan;
  const size_t S = (size_t)cs->net.slots, R = (size_t)cs->net.n_reach;
  double *parts[4] = {cs->net.hlevel, cs->net.qstrm, cs->net.velocity, cs->net.outlet_hlev};
  const size_t counts[4] = {S, S, S, R};
  size_t off = 0;
  for (int k = 0; k < 4; k++) {
    CK(cudaMemcpyAsync(to_snapshot ? cs->snap + off : parts[k], to_snapshot ? parts[k] : cs->snap + off, counts[k] * sizeof(double),
                       cudaMemcpyDeviceToDevice, c->stream));
    off += counts[k];
  }
  return 0;
}
static int channel_snapshot(tribs_ctx *c) {
  if (!c->chan) return 0;
  if (!c->chan->snap && dev_alloc(c, &c->chan->snap, (size_t)3 * c->chan->net.slots + c->chan->net.n_reach)) return 1;
  return channel_copy_state(c, true);
}
static int channel_rollback(tribs_ctx *c) {
  if (!c->chan) return 0;
  if (!c->chan->snap) return fail("tribs_gpu_rollback: the channel network was attached after the snapshot");
  return channel_copy_state(c, false);
}
static bool channel_write(tribs_ctx *c, FILE *f) {
  if (!c->chan) return true;
  ChannelState *cs = c->chan;
  const int32_t hdr[2] = {cs->net.n_reach, cs->net.slots};
  if (fwrite("CHAN", 1, 4, f) != 4 || fwrite(hdr, sizeof(int32_t), 2, f) != 2) return false;
  const double *parts[4] = {cs->net.hlevel, cs->net.qstrm, cs->net.velocity, cs->net.outlet_hlev};
  const size_t counts[4] = {(size_t)cs->net.slots, (size_t)cs->net.slots, (size_t)cs->net.slots, (size_t)cs->net.n_reach};
  for (int k = 0; k < 4; k++) {
    std::vector<double> b(counts[k]);
    if (cudaMemcpy(b.data(), parts[k], counts[k] * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return false;
    if (fwrite(b.data(), sizeof(double), counts[k], f) != counts[k]) return false;
  }
  return true;
}
static int channel_read(tribs_ctx *c, FILE *f) {
  ChannelState *cs = c->chan;
  char magic[4];
  int32_t hdr[2];
  if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "CHAN", 4) != 0 || fread(hdr, sizeof(int32_t), 2, f) != 2)
    return fail("tribs_gpu_load_state: the channel section of the checkpoint is unreadable");
  if (hdr[0] != cs->net.n_reach || hdr[1] != cs->net.slots) return fail("tribs_gpu_load_state: the checkpoint's channel network is another one");
  double *parts[4] = {cs->net.hlevel, cs->net.qstrm, cs->net.velocity, cs->net.outlet_hlev};
  const size_t counts[4] = {(size_t)cs->net.slots, (size_t)cs->net.slots, (size_t)cs->net.slots, (size_t)cs->net.n_reach};
  for (int k = 0; k < 4; k++) {
    std::vector<double> b(counts[k]);
    if (fread(b.data(), sizeof(double), counts[k], f) != counts[k]) return fail("tribs_gpu_load_state: the channel section of the checkpoint is truncated");
    CK(cudaMemcpy(parts[k], b.data(), counts[k] * sizeof(double), cudaMemcpyHostToDevice));
  }
  return 0;
}

template <typename T>
static int chan_upload(tribs_ctx *c, T **dst, const std::vector<T> &h) {
  if (dev_alloc(c, dst, h.size())) return 1;
  CK(cudaMemcpy(*dst, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int tribs_gpu_channel_init(tribs_handle_t c, const tribs_channel_t *ch) {
  if (!c || !ch) return fail("tribs_gpu_channel_init: null argument");
  if (channel_of(c)) return fail("tribs_gpu_channel_init: the handle already has a channel network");
  if (ch->n_reach <= 0 || !ch->reach_off || !ch->reach_node || !ch->width || !ch->length || !ch->slope)
    return fail("tribs_gpu_channel_init: reach lists or geometry missing");
  if (!(ch->roughness > 0.0) || !(ch->dt_s > 0.0)) return fail("tribs_gpu_channel_init: roughness and dt_s must be positive");
  if (ch->percolation_option != 0 || ch->reservoir_option != 0)
    return fail("tribs_gpu_channel_init: OPTPERCOLATION / OPTRESERVOIR other than 0 are not built on the device");
  CK(cudaSetDevice(c->device));
  const int R = ch->n_reach, E = ch->reach_off[R], S = c->n_stream;
  if (ch->reach_off[0] != 0 || E <= 0) return fail("tribs_gpu_channel_init: reach_off must start at 0 and grow");
  std::unordered_map<int32_t, int32_t> slot_of;
  slot_of.reserve((size_t)S * 2);
  for (int k = 0; k < S; k++) slot_of[c->stream_cells_sweep[k]] = k;
  std::vector<int32_t> off(ch->reach_off, ch->reach_off + R + 1), slot(E);
  std::vector<double> bis(E), siis(E), C(E), Y1(E, 0.0), Y2(E, 0.0), Y3(E, 0.0);
  int max_n = 0;
  for (int id = 0; id < R; id++) {
    const int e0 = off[id], n = off[id + 1] - e0, m = n - 1, m1 = n - 2;
    if (n < 2) return fail("tribs_gpu_channel_init: a reach needs at least two nodes (head and outlet)");
    max_n = std::max(max_n, n);
    for (int i = 0; i < n; i++) {
      const int32_t cell = ch->reach_node[e0 + i];
      if (cell == c->n) {
        if (i != m) return fail("tribs_gpu_channel_init: the basin outlet can only end a reach");
        slot[e0 + i] = S;
      } else {
        auto it = slot_of.find(cell);
        if (it == slot_of.end()) return fail("tribs_gpu_channel_init: reach node " + std::to_string(cell) + " is not a stream cell");
        slot[e0 + i] = it->second;
      }
      bis[e0 + i] = ch->uniform_width > 0. ? ch->uniform_width : ch->width[e0 + i];
      if (!(bis[e0 + i] > 0.0)) return fail("tribs_gpu_channel_init: channel widths must be positive");
    }
    // InitializeStreamReach (tKinemat.cpp:1157-1190): node 0 takes the slope of its own link, node i+1 that of link i
    double *s = siis.data() + e0;
    s[0] = ch->slope[e0] <= 0 ? 0.0050667 : ch->slope[e0];
    for (int i = 0; i < m; i++) s[i + 1] = ch->slope[e0 + i] <= 0 ? 0.0050667 : ch->slope[e0 + i];
    // ComputeCoefficientArrays (tKinemat.cpp:1267-1285), the reference's expressions on the host's FP64
    const double *ais = ch->length + e0, *b = bis.data() + e0, dt = ch->dt_s, rif = ch->roughness;
    for (int i = 0; i < n; i++) {
      C[e0 + i] = 1. / rif * sqrt(s[i]) * b[i];
      if (i < m1) {
        Y1[e0 + i] = 1. / 6. * ais[i + 1] * b[i + 2] / dt;
        Y2[e0 + i] = 1. / 3. * (ais[i] + ais[i + 1]) * b[i + 1] / dt;
        Y3[e0 + i] = 1. / 6. * ais[i] * b[i] / dt;
      } else if (i == m1) {
        Y2[e0 + m1] = 1. / 3. * ais[m1] * b[m] / dt;
        Y3[e0 + m1] = 1. / 6. * ais[m1] * b[m1] / dt;
      }
    }
  }
  ChannelState *cs = new ChannelState();
  ChanNet &nt = cs->net;
  nt.n_reach = R; nt.max_n = max_n; nt.slots = S + 1; nt.roughness = ch->roughness;
  cs->n_entries = E;
  int32_t *d_off = nullptr, *d_slot = nullptr;
  double *d_bis = nullptr, *d_siis = nullptr, *d_C = nullptr, *d_Y1 = nullptr, *d_Y2 = nullptr, *d_Y3 = nullptr;
  if (chan_upload(c, &d_off, off) || chan_upload(c, &d_slot, slot) || chan_upload(c, &d_bis, bis) || chan_upload(c, &d_siis, siis) ||
      chan_upload(c, &d_C, C) || chan_upload(c, &d_Y1, Y1) || chan_upload(c, &d_Y2, Y2) || chan_upload(c, &d_Y3, Y3)) {
    delete cs;
    return 1;
  }
  nt.off = d_off; nt.slot = d_slot; nt.bis = d_bis; nt.siis = d_siis; nt.C = d_C; nt.Y1 = d_Y1; nt.Y2 = d_Y2; nt.Y3 = d_Y3;
  std::vector<double> zeros((size_t)std::max(S + 1, R), 0.0);
  auto state = [&](double **dst, const double *init, int count, bool by_cell) -> int {
    std::vector<double> h((size_t)count, 0.0);
    if (init) {
      if (by_cell) {                        // per caller cell index, slot n_cells = the basin outlet
        for (int k = 0; k < S; k++) h[k] = init[c->stream_cells_sweep[k]];
        h[S] = init[c->n];
      } else
        for (int k = 0; k < count; k++) h[k] = init[k];
    }
    return chan_upload(c, dst, h);
  };
  if (state(&nt.hlevel, ch->hlevel0, S + 1, true) || state(&nt.qstrm, ch->qstrm0, S + 1, true) || state(&nt.velocity, nullptr, S + 1, true) ||
      state(&nt.outlet_hlev, ch->outlet_hlev0, R, false) || dev_alloc(c, &nt.ws, (size_t)kChanWsVecs * (max_n + 2)) ||
      dev_alloc(c, &nt.status, 4) || dev_alloc(c, &nt.cycles, 4)) {
    delete cs;
    return 1;
  }
  CK(cudaMemset(nt.ws, 0, (size_t)kChanWsVecs * (max_n + 2) * sizeof(double)));
  CK(cudaMemset(nt.status, 0, 4 * sizeof(int32_t)));
  CK(cudaMemset(nt.cycles, 0, 4 * sizeof(long long)));
  // the six vectors of the linear solve live in shared memory when the longest reach fits (kChanSharedRows - 2 = 4 072 nodes), else in the
  // workspace in global memory (slower, same results)
  cs->shared_rows = (max_n + 2 <= kChanSharedRows && !getenv("TRIBS_CHAN_GLOBAL_VECTORS")) ? max_n + 2 : 0;   // the variable: test hook
  cs->smem = ((size_t)kChanInterface + (size_t)kChanSharedVecs * cs->shared_rows) * sizeof(double);
  CK(cudaFuncSetAttribute(channel_route_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)(((size_t)kChanInterface + (size_t)kChanSharedVecs * kChanSharedRows) * sizeof(double))));
  // nothing is allocated or loaded inside a timed batch: the outlet series buffer exists, the kernel's code is resident
  cs->qout_cap = 1024;
  if (dev_alloc(c, &cs->d_qout, (size_t)cs->qout_cap)) { delete cs; return 1; }
  channel_route_kernel<<<1, kChanThreads, cs->smem, c->stream>>>(cs->net, c->d_qeff_now, 0, cs->d_qout, cs->shared_rows);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  channel_of(c) = cs;
  return 0;
}

// Routes the channels over the steps of the last batch (n_steps >= 1: the per-step lateral inflows tribs_gpu_run left on
// the device) or over the single step whose ROUTE phase ran last (n_steps == 0); qout[k] = tKinemat::Qout after step k.
extern "C" int tribs_gpu_channel_route(tribs_handle_t c, int n_steps, double *qout) {
  if (!c || !qout || n_steps < 0) return fail("tribs_gpu_channel_route: bad argument");
  ChannelState *cs = channel_of(c);
  if (!cs) return fail("tribs_gpu_channel_route: tribs_gpu_channel_init has not been called");
  CK(cudaSetDevice(c->device));
  if (check_sweep_abort(c)) return 1;
  const double *src = c->d_qeff_now;
  const int steps = std::max(n_steps, 1);
  if (n_steps >= 1) {
    PipeState *ps = pipe_of(c);
    if (!ps || n_steps > ps->last_steps) return fail("tribs_gpu_channel_route: more steps than the last tribs_gpu_run held");
    src = ps->d_qeff_out;
  }
  if (steps > cs->qout_cap) {
    if (dev_alloc(c, &cs->d_qout, (size_t)steps)) return 1;
    cs->qout_cap = steps;
  }
  channel_route_kernel<<<1, kChanThreads, cs->smem, c->stream>>>(cs->net, src, steps, cs->d_qout, cs->shared_rows);
  CK(cudaGetLastError());
  cs->launches++; c->launches++;
  int32_t status[4];
  CK(cudaMemcpyAsync(qout, cs->d_qout, (size_t)steps * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(status, cs->net.status, sizeof status, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (status[0] == 1) return fail("tribs_gpu_channel_route: the Newton direction of a reach is not a descent direction (the reference exits here: 'Roundoff problem in lnsrch')");
  return 0;
}

// hlevel / qstrm / velocity: per stream slot (tribs_gpu_stream_cells order) + the basin outlet; outlet_hlev: per reach;
// counters[7]: Newton iterations, line-search backtracks, reaches that hit MAXITS since init, then thread 0's clock
// cycles in the band solve, the line-search sums, the function sums, and the whole kernel. Any pointer may be NULL.
extern "C" int tribs_gpu_channel_sync(tribs_handle_t c, double *hlevel, double *qstrm, double *velocity, double *outlet_hlev,
                                      int64_t *counters) {
  if (!c) return fail("tribs_gpu_channel_sync: null handle");
  ChannelState *cs = channel_of(c);
  if (!cs) return fail("tribs_gpu_channel_sync: tribs_gpu_channel_init has not been called");
  CK(cudaSetDevice(c->device));
  const size_t slots = (size_t)cs->net.slots * sizeof(double);
  if (hlevel) CK(cudaMemcpyAsync(hlevel, cs->net.hlevel, slots, cudaMemcpyDeviceToHost, c->stream));
  if (qstrm) CK(cudaMemcpyAsync(qstrm, cs->net.qstrm, slots, cudaMemcpyDeviceToHost, c->stream));
  if (velocity) CK(cudaMemcpyAsync(velocity, cs->net.velocity, slots, cudaMemcpyDeviceToHost, c->stream));
  if (outlet_hlev) CK(cudaMemcpyAsync(outlet_hlev, cs->net.outlet_hlev, (size_t)cs->net.n_reach * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  int32_t status[4] = {0, 0, 0, 0};
  if (counters) CK(cudaMemcpyAsync(status, cs->net.status, sizeof status, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (counters) {
    counters[0] = status[1]; counters[1] = status[2]; counters[2] = status[3];
    long long cy[4];
    CK(cudaMemcpy(cy, cs->net.cycles, sizeof cy, cudaMemcpyDeviceToHost));
    for (int k = 0; k < 4; k++) counters[3 + k] = cy[k];
  }
  return 0;
}
