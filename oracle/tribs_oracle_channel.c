/* TEST INFRASTRUCTURE - CPU restatement of the reference's kinematic-wave channel router, used only by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the checker of the device path.
 *
 * Follows tKinemat::RunRoutingModel (src/tFlowNet/tKinemat.cpp:803-897) for OPTPERCOLATION 0 and
 * OPTRESERVOIR 0, one call per time step, after the hillslope routing of that step:
 *   InitializeStreamReach 1136-1252, ComputeCoefficientArrays 1267-1285, AssignLateralInflux 1296-1494
 *   (the per-node values RetrieveQeff 1504-1551 returned for this step come in as an array),
 *   AssignQin 1561-1573, SolveForTwoNodeReach 1736-1765 (rtsafe + polynomialH, src/Mathutil/mathutil.cpp:1682-1759),
 *   KinematWave 1788-1931 (globally convergent Newton: ComputeFunction 1950-1998, fmin 2007-2013, ComputeJacobian
 *   2052-2072, SolveLinearSystem 2085-2145 = bandec/banbks 2154-2230 with partial pivoting, lnsrch 2321-2412),
 *   ComputeQout 1582-1586, UpdateStreamVars 1617-1672.
 * Pinned against the reference's own per-step outlet discharge and node levels (tests/test_channel.py). */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "tribs_oracle_channel.h"

#define ERR_SLOPE 0.0050667 /* tKinemat.cpp:1158, 1188 */

static double fmax2(double a, double b) { return a > b ? a : b; }

/* mathutil.cpp:1754-1759 */
static void polynomial_h(double x, double c1, double c2, double c3, double *fn, double *df) {
  double fk2 = 5. / 3., fk3 = 2. / 3.;
  *fn = c1 * pow(x, fk2) + c2 * x + c3;
  *df = fk2 * c1 * pow(x, fk3) + c2;
}

/* mathutil.cpp:1682-1743 */
static double rtsafe_h(double c1, double c2, double c3, double x1, double x2, double xacc, double xguess) {
  int j;
  double df, dx, dxold, f, fh, fl, temp, xh, xl, rts;
  polynomial_h(x1, c1, c2, c3, &fl, &df);
  polynomial_h(x2, c1, c2, c3, &fh, &df);
  if (fl == 0.0) return x1;
  if (fh == 0.0) return x2;
  if (fl < 0.0) { xl = x1; xh = x2; } else { xh = x1; xl = x2; }
  rts = xguess;
  dxold = fabs(x2 - x1);
  dx = dxold;
  polynomial_h(rts, c1, c2, c3, &f, &df);
  for (j = 1; j <= 200; j++) {
    if ((((rts - xh) * df - f) * ((rts - xl) * df - f) > 0.0) || (fabs(2.0 * f) > fabs(dxold * df))) {
      dxold = dx;
      dx = 0.5 * (xh - xl);
      rts = xl + dx;
      if (xl == rts) return rts;
    } else {
      dxold = dx;
      dx = f / df;
      temp = rts;
      rts -= dx;
      if (temp == rts) return rts;
    }
    if (fabs(dx) < xacc) return rts;
    polynomial_h(rts, c1, c2, c3, &f, &df);
    if (f < 0.0) xl = rts; else xh = rts;
  }
  return 0.0;
}

typedef struct {
  int n, m, m1;
  double *his, *bis, *ais, *siis, *rifis, *reis, *C, *Y1, *Y2, *Y3;
  double *X, *X1, *XR, *F, *gradf, *aa, *bb, *cc, *xfk2;
  double *a, *al;       /* compact band matrix (N+1) x 4 and its multipliers (N+1) x 2 */
  unsigned long *indx;
} reach_ws;

/* tKinemat.cpp:1950-1998; X holds the levels of nodes 1..m at (t+1), HLev those of nodes 0..m at t */
static void compute_function(reach_ws *w, double *F, const double *X, double qit, double hupp) {
  int i, M = w->m, M1 = w->m1;
  const double *c = w->C, *y1 = w->Y1, *y2 = w->Y2, *y3 = w->Y3, *Reff = w->reis, *HLev = w->his;
  double fk2 = 5. / 3.;
  for (i = 0; i < M; i++) w->xfk2[i] = pow(X[i], fk2);
  for (i = 0; i < M; i++) {
    if (i == 0)
      F[0] = 0.5 * c[2] * w->xfk2[1] - qit + y3[0] * (hupp - HLev[0]) + y2[0] * (X[0] - HLev[1]) + y1[0] * (X[1] - HLev[2]) - Reff[0];
    else if (i == M1)
      F[M1] = 0.5 * c[M] * w->xfk2[M1] - 0.5 * c[M1] * w->xfk2[M1 - 1] + y3[M1] * (X[M1 - 1] - HLev[M1]) + y2[M1] * (X[M1] - HLev[M]) - Reff[M1];
    else
      F[i] = 0.5 * c[i + 2] * w->xfk2[i + 1] - 0.5 * c[i] * w->xfk2[i - 1] + y3[i] * (X[i - 1] - HLev[i]) + y2[i] * (X[i] - HLev[i + 1])
             + y1[i] * (X[i + 1] - HLev[i + 2]) - Reff[i];
  }
}

/* tKinemat.cpp:2007-2013 */
static double half_sum_sq(const double *f, int n) {
  int i;
  double sum = 0.0;
  for (i = 0; i < n; i++) sum += f[i] * f[i];
  return 0.5 * sum;
}

/* tKinemat.cpp:2052-2072 */
static void compute_jacobian(reach_ws *w, const double *X, int N) {
  int i;
  double f1, f2, fk1 = 5. / 6., fk3 = 2. / 3.;
  const double *c = w->C;
  f1 = pow(X[0], fk3);
  for (i = 0; i < N - 1; i++) {
    f2 = pow(X[i + 1], fk3);
    w->bb[i] = w->Y2[i];
    w->cc[i] = fk1 * c[i + 2] * f2 + w->Y1[i];
    w->aa[i + 1] = -fk1 * c[i + 1] * f1 + w->Y3[i + 1];
    f1 = f2;
  }
  w->bb[N - 1] = fk1 * c[N] * pow(X[N - 1], fk3) + w->Y2[N - 1];
}

#define A_(i, j) a[(size_t)(i) * 4 + (j)]
#define AL_(i, j) al[(size_t)(i) * 2 + (j)]
/* tKinemat.cpp:2085-2145 with bandec 2154-2196 and banbks 2203-2230 (M1 = M2 = 1); XR is 1-based on entry, 0-based on exit */
static void solve_linear_system(reach_ws *w, double *XR, int N) {
  double *a = w->a, *al = w->al, dum;
  unsigned long *indx = w->indx, i, j, k, l, NN = (unsigned long)N;
  const int mm = 3;
  for (i = 0; i < 4; i++) A_(0, i) = 999.;
  A_(1, 0) = 999.; A_(1, 1) = 999.; A_(1, 2) = w->bb[0]; A_(1, 3) = w->cc[0];
  for (i = 2; i < NN; i++) { A_(i, 0) = 999; A_(i, 1) = w->aa[i - 1]; A_(i, 2) = w->bb[i - 1]; A_(i, 3) = w->cc[i - 1]; }
  A_(NN, 0) = 999; A_(NN, 1) = w->aa[NN - 1]; A_(NN, 2) = w->bb[NN - 1]; A_(NN, 3) = 999.;
  /* bandec */
  l = 1;
  for (i = 1; i <= 1; i++) {
    for (j = 1 + 2 - i; j <= (unsigned long)mm; j++) A_(i, j - l) = A_(i, j);
    l--;
    for (j = mm - l; j <= (unsigned long)mm; j++) A_(i, j) = 0.0;
  }
  l = 1;
  for (k = 1; k <= NN; k++) {
    dum = A_(k, 1);
    i = k;
    if (l < NN) l++;
    for (j = k + 1; j <= l; j++)
      if (fabs(A_(j, 1)) > fabs(dum)) { dum = A_(j, 1); i = j; }
    indx[k] = i;
    if (dum == 0.0) A_(k, 1) = 1.0E-20;
    if (i != k)
      for (j = 1; j <= (unsigned long)mm; j++) { double t = A_(k, j); A_(k, j) = A_(i, j); A_(i, j) = t; }
    for (i = k + 1; i <= l; i++) {
      dum = A_(i, 1) / A_(k, 1);
      AL_(k, i - k) = dum;
      for (j = 2; j <= (unsigned long)mm; j++) A_(i, j - 1) = A_(i, j) - dum * A_(k, j);
      A_(i, mm) = 0.0;
    }
  }
  /* banbks */
  l = 1;
  for (k = 1; k <= NN; k++) {
    i = indx[k];
    if (i != k) { double t = XR[k]; XR[k] = XR[i]; XR[i] = t; }
    if (l < NN) l++;
    for (i = k + 1; i <= l; i++) XR[i] -= AL_(k, i - k) * XR[k];
  }
  l = 1;
  for (i = NN; i >= 1; i--) {
    dum = XR[i];
    for (k = 2; k <= l; k++) dum -= A_(i, k) * XR[k + i - 1];
    XR[i] = dum / A_(i, 1);
    if (l < (unsigned long)mm) l++;
  }
  for (i = 0; i < NN; i++) XR[i] = XR[i + 1];
}

/* tKinemat.cpp:2321-2412; returns 0, or 2 where the reference would exit on a non-descent direction */
static int lnsrch(reach_ws *w, int N, const double *xold, double fold, const double *g, double *p, double *x, double *f,
                  double stpmax, int *check, double *F, double qit, double hupp) {
  const double ALF = 1.0e-4, TOLX = 1.0e-7;
  int i;
  double a, alam, alam2, alamin, b, disc, f2, rhs1, rhs2, slope, sum, temp, test, tmplam = 0.0;
  f2 = alam2 = 0.0;
  *check = 0;
  for (sum = 0.0, i = 0; i < N; i++) sum += p[i] * p[i];
  sum = sqrt(sum);
  if (sum > stpmax)
    for (i = 0; i < N; i++) p[i] *= stpmax / sum;
  for (slope = 0.0, i = 0; i < N; i++) slope += g[i] * p[i];
  if (slope >= 0.0) return 2;
  test = 0.0;
  for (i = 0; i < N; i++) {
    temp = fabs(p[i]) / fmax2(fabs(xold[i]), 1.0);
    if (temp > test) test = temp;
  }
  alamin = TOLX / test;
  alam = 1.0;
  for (;;) {
    for (i = 0; i < N; i++) {
      x[i] = xold[i] + alam * p[i];
      if (x[i] < 0.0) x[i] = 1.0E-6;
    }
    compute_function(w, F, x, qit, hupp);
    *f = half_sum_sq(F, N);
    if (alam < alamin) {
      for (i = 0; i < N; i++) x[i] = xold[i];
      *check = 1;
      return 0;
    } else if (*f <= fold + ALF * alam * slope)
      return 0;
    else {
      if (alam == 1.0)
        tmplam = -slope / (2.0 * (*f - fold - slope));
      else {
        rhs1 = *f - fold - alam * slope;
        rhs2 = f2 - fold - alam2 * slope;
        a = (rhs1 / (alam * alam) - rhs2 / (alam2 * alam2)) / (alam - alam2);
        b = (-alam2 * rhs1 / (alam * alam) + alam * rhs2 / (alam2 * alam2)) / (alam - alam2);
        if (a == 0.0)
          tmplam = -slope / (2.0 * b);
        else {
          disc = b * b - 3.0 * a * slope;
          if (disc < 0.0) tmplam = 0.5 * alam;
          else if (b <= 0.0) tmplam = (-b + sqrt(disc)) / (3.0 * a);
          else tmplam = -slope / (b + sqrt(disc));
        }
        if (tmplam > 0.5 * alam) tmplam = 0.5 * alam;
      }
    }
    alam2 = alam;
    f2 = *f;
    alam = fmax2(tmplam, 0.1 * alam);
  }
}

static void update_hs_shifted(const double *xnew, double *xold, double hupp, int N) {
  int i;
  xold[0] = hupp;
  for (i = 0; i < N; i++) xold[i + 1] = xnew[i];
}

/* tKinemat.cpp:1788-1931; returns the number of Newton iterations taken (0: the old levels already solve the system),
   -1 where MAXITS was exceeded, -2 for lnsrch's exit */
static int kinemat_wave(reach_ws *w, double qit, double hupp, int *check) {
  const double TOLF = 1.0e-4, TOLMIN = 1.0e-6, TOLX = 1.0e-7, STPMX = 100.0;
  int iter, i, m = w->m, m1 = w->m1;
  double *X = w->X, *X1 = w->X1, *XR = w->XR, *F = w->F, *gradf = w->gradf, *aa = w->aa, *bb = w->bb, *cc = w->cc;
  double *HLev = w->his;
  double den, f, fold, stpmax, sum, temp, test;
  for (i = 0; i < m; i++) X[i] = X1[i] = HLev[i + 1];
  compute_function(w, F, X1, qit, hupp);
  f = half_sum_sq(F, m);
  test = 0.0;
  for (i = 0; i < m; i++)
    if (fabs(F[i]) > test) test = fabs(F[i]);
  if (test < 0.01 * TOLF) { *check = 0; return 0; }
  for (sum = 0.0, i = 0; i < m; i++) sum += X[i] * X[i];
  stpmax = STPMX * fmax2(sqrt(sum), (double)m);
  for (iter = 1; iter <= 200; iter++) {
    compute_jacobian(w, X1, m);
    gradf[0] = bb[0] * F[0] + aa[1] * F[1];
    for (i = 1; i < m1; i++) gradf[i] = cc[i - 1] * F[i - 1] + bb[i] * F[i] + aa[i + 1] * F[i + 1];
    gradf[m1] = bb[m1] * F[m1] + cc[m1 - 1] * F[m1 - 1];
    for (i = 0; i < m; i++) X[i] = X1[i];
    fold = f;
    for (i = 0; i < m; i++) XR[i + 1] = (F[i] == 0) ? F[i] : -F[i];
    solve_linear_system(w, XR, m);
    if (lnsrch(w, m, X, fold, gradf, XR, X1, &f, stpmax, check, F, qit, hupp)) return -2;
    test = 0.0;
    for (i = 0; i < m; i++)
      if (fabs(F[i]) > test) test = fabs(F[i]);
    if (test < TOLF) {
      *check = 0;
      update_hs_shifted(X1, HLev, hupp, m);
      return iter;
    }
    if (*check) {
      test = 0.0;
      den = fmax2(f, 0.5 * m);
      for (i = 0; i < m; i++) {
        temp = fabs(gradf[i]) * fmax2(fabs(X1[i]), 1.0) / den;
        if (temp > test) test = temp;
      }
      *check = (test < TOLMIN ? 1 : 0);
      update_hs_shifted(X1, HLev, hupp, m);
      return iter;
    }
    test = 0.0;
    for (i = 0; i < m; i++) {
      temp = (fabs(X1[i] - X[i])) / fmax2(fabs(X1[i]), 1.0);
      if (temp > test) test = temp;
    }
    if (test < TOLX) {
      update_hs_shifted(X1, HLev, hupp, m);
      return iter;
    }
  }
  return -1;
}

static double node_q(const reach_ws *w, int i) { return (w->bis[i] * pow(w->his[i], (5. / 3.)) * sqrt(w->siis[i]) / w->rifis[i]); }
static double node_vel(const reach_ws *w, int i) { return (pow(w->his[i], (2. / 3.)) * sqrt(w->siis[i]) / w->rifis[i]); }

static void ws_alloc(reach_ws *w, int n) {
  size_t nn = (size_t)n + 2;
  double **v[] = {&w->his, &w->bis, &w->ais, &w->siis, &w->rifis, &w->reis, &w->C, &w->Y1, &w->Y2, &w->Y3, &w->X, &w->X1,
                  &w->XR, &w->F, &w->gradf, &w->aa, &w->bb, &w->cc, &w->xfk2};
  size_t k;
  for (k = 0; k < sizeof v / sizeof v[0]; k++) *v[k] = (double *)calloc(nn, sizeof(double));
  w->a = (double *)calloc(nn * 4, sizeof(double));
  w->al = (double *)calloc(nn * 2, sizeof(double));
  w->indx = (unsigned long *)calloc(nn, sizeof(unsigned long));
}
static void ws_free(reach_ws *w) {
  double *v[] = {w->his, w->bis, w->ais, w->siis, w->rifis, w->reis, w->C, w->Y1, w->Y2, w->Y3, w->X, w->X1,
                 w->XR, w->F, w->gradf, w->aa, w->bb, w->cc, w->xfk2, w->a, w->al};
  size_t k;
  for (k = 0; k < sizeof v / sizeof v[0]; k++) free(v[k]);
  free(w->indx);
}

/* One time step of the channel router over every reach in the reference's list order. `qeff` holds, per cell index
   (slot n_cells = the basin outlet), what RetrieveQeff returned for this step. Returns 0, or a negative code of the
   reach solver; *qout = tKinemat::Qout after the last reach, *iters the Newton iterations taken over all reaches. */
int orc_channel_step(const orc_channel *ch, orc_channel_state *s, const double *qeff, double dt, double *qout, int *iters) {
  int id, i, rc = 0, total_iters = 0;
  const int basin_outlet = ch->n_cells;
  double q_out = 0.0;
  for (id = 0; id < ch->n_reach; id++) s->qstrm[ch->node[ch->off[id + 1] - 1]] = 0.0; /* tKinemat.cpp:823-824 */
  for (id = 0; id < ch->n_reach; id++) {
    const int e0 = ch->off[id], n = ch->off[id + 1] - e0, m = n - 1, m1 = n - 2;
    const int *node = ch->node + e0;
    const int ends_at_basin_outlet = node[m] == basin_outlet;
    reach_ws w;
    double qin, qit, h0;
    int check = 0;
    memset(&w, 0, sizeof w);
    ws_alloc(&w, n);
    w.n = n; w.m = m; w.m1 = m1;
    /* InitializeStreamReach */
    w.siis[0] = ch->slope[e0];
    if (w.siis[0] <= 0) w.siis[0] = ERR_SLOPE;
    for (i = 0; i < m; i++) {
      w.his[i] = s->hlevel[node[i]];
      w.bis[i] = ch->width[e0 + i];
      w.ais[i] = ch->length[e0 + i];
      w.siis[i + 1] = ch->slope[e0 + i] <= 0 ? ERR_SLOPE : ch->slope[e0 + i];
    }
    w.his[m] = s->outlet_hlev[id];
    w.bis[m] = ch->width[e0 + m];
    for (i = 0; i < n; i++) {
      w.rifis[i] = ch->roughness;
      if (ch->uniform_width > 0.) w.bis[i] = ch->uniform_width;
    }
    /* ComputeCoefficientArrays */
    for (i = 0; i < n; i++) {
      w.C[i] = 1. / w.rifis[i] * sqrt(w.siis[i]) * w.bis[i];
      if (i < m1) {
        w.Y1[i] = 1. / 6. * w.ais[i + 1] * w.bis[i + 2] / dt;
        w.Y2[i] = 1. / 3. * (w.ais[i] + w.ais[i + 1]) * w.bis[i + 1] / dt;
        w.Y3[i] = 1. / 6. * w.ais[i] * w.bis[i] / dt;
      } else if (i == m1) {
        w.Y2[m1] = 1. / 3. * w.ais[m1] * w.bis[m] / dt;
        w.Y3[m1] = 1. / 6. * w.ais[m1] * w.bis[m1] / dt;
      }
    }
    /* AssignLateralInflux */
    for (i = 0; i < n; i++) w.reis[i] = 0.0;
    for (i = 0; i < m; i++) w.reis[i] = qeff[node[i]];
    if (ends_at_basin_outlet) w.reis[m] = qeff[basin_outlet];
    for (i = 0; i < m; i++) {
      if (i < m1)
        w.reis[i] = (w.reis[i] + w.reis[i + 1]) / 2.;
      else if (i == m1) {
        if (ends_at_basin_outlet) w.reis[i] += w.reis[m];
        w.reis[i] /= 2.;
      }
    }
    /* AssignQin */
    qin = s->qstrm[node[0]];
    qit = 0.5 * qin;
    h0 = pow(qin * w.rifis[0] / (w.bis[0] * sqrt(w.siis[0])), 3. / 5.);
    if (n == 2) { /* SolveForTwoNodeReach */
      double c1 = 0.5 * w.C[1], c2 = 0.5 * w.Y2[0], xx;
      double c3 = 2 * w.Y3[0] * (h0 - w.his[0]) - c2 * w.his[1] - qit - w.reis[0];
      if (c3 > 0.0) xx = w.his[1];
      else xx = rtsafe_h(c1, c2, c3, 0., 40., 1.0E-6, h0);
      w.his[0] = h0;
      w.his[1] = xx;
    } else {
      int it = kinemat_wave(&w, qit, h0, &check);
      if (it < 0) rc = it; else total_iters += it;
    }
    /* ComputeQout, UpdateStreamVars */
    q_out = w.bis[m] * pow(w.his[m], 5. / 3.) * sqrt(w.siis[m]) / w.rifis[m];
    for (i = 0; i < m; i++) {
      s->hlevel[node[i]] = w.his[i];
      s->qstrm[node[i]] = node_q(&w, i);
      s->velocity[node[i]] = node_vel(&w, i);
    }
    if (ends_at_basin_outlet) s->hlevel[node[m]] = w.his[m];
    s->outlet_hlev[id] = w.his[m];
    s->qstrm[node[m]] += node_q(&w, m);
    ws_free(&w);
  }
  *qout = q_out;
  if (iters) *iters = total_iters;
  return rc;
}
