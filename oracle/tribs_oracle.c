/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See tribs_oracle.h for scope and pinning.
 *
 * Restatement notes
 *  - The reference keeps one cell's working values in tHydroModel data members and its
 *    helpers read/write them implicitly. Here that working set is the explicit
 *    struct `cw` threaded through every helper; side effects are reproduced on purpose
 *    (e.g. total_moist() zeroing nwt1, tHydroModel.cpp:2252-2259).
 *  - Build with -O2 -ffp-contract=off (the reference's g++ -O2 line has no FMA contraction),
 *    so that this file is bit-identical to the reference on the same libm.
 *  - PI is the reference's 8-digit constant (src/Headers/Definitions.h:54); LAMBEPS from
 *    src/tHydro/tHydroModel.h:32.
 */
#include "tribs_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define REF_PI 3.1415926
#define LAMBEPS 2.2204E-16
#define RADAR_MM 0.1

typedef struct {
  /* soil */
  double ks, ths, thr, lam, psib, f, ar, uar, eps, dbr;
  /* state: 0 = Old, 1 = New */
  double nwt0, nwt1, mu0, mu1, mi0, mi1, nf0, nf1, nt0, nt1, ru0, ru1, ri0, ri1;
  double qpin, qpout, isv;
  double r, r1, cosv, sinv;
  double srf, hsrf, esrf, psrf, satsrf, sbsrf;
  double g, sein, se0, thri, thre;
} cw;

/* ------------------------------------------------------------------------------------
 * Brooks-Corey profile integrals (reference: tHydroModel.cpp:2240-2315)
 * ---------------------------------------------------------------------------------- */
static int lam_is_one(const cw *c) { return c->lam > (1.0 - 1.0E-6) && c->lam < (1.0 + 1.0E-6); }

static double total_moist(cw *c, double nwt) {
  double dm = 0.0;
  if (nwt >= fabs(c->psib)) {
    if (lam_is_one(c))
      dm = c->thr * (nwt + c->psib) - fabs(c->psib) * (c->ths - c->thr) * log((-c->psib) / nwt) - c->ths * c->psib;
    else
      dm = c->thr * (nwt + c->psib) - c->ths * c->psib -
           (c->ths - c->thr) / (c->lam - 1.0) * (c->psib + nwt * pow((-c->psib / nwt), c->lam));
  } else if (nwt == 0.0) {
    dm = 0.0;
  } else {
    c->nwt1 = 0.0; /* 2258: water table inside the capillary fringe is moved to the surface */
  }
  return dm;
}

static double upper_moist(cw *c, double nf, double nwt) {
  double dm;
  if (nf <= (nwt + c->psib)) {
    if (lam_is_one(c))
      dm = c->thr * nf + fabs(c->psib) * (c->ths - c->thr) * log(nwt / (nwt - nf));
    else
      dm = c->thr * nf - (c->ths - c->thr) / (c->lam - 1.0) *
                             ((nf - nwt) * pow((c->psib / (nf - nwt)), c->lam) + nwt * pow((-c->psib / nwt), c->lam));
  } else {
    if (lam_is_one(c))
      dm = total_moist(c, nwt) + c->ths * c->psib + c->ths * (nf - (nwt + c->psib));
    else
      dm = c->thr * (nwt + c->psib) -
           (c->ths - c->thr) / (c->lam - 1.0) * (c->psib + nwt * pow((-c->psib / nwt), c->lam)) +
           c->ths * (nf - (nwt + c->psib));
  }
  return dm;
}

static double lower_moist(const cw *c, double nf, double nwt) {
  double dm;
  if (nf <= (nwt + c->psib)) {
    if (lam_is_one(c))
      dm = c->thr * (nwt + c->psib - nf) - fabs(c->psib) * (c->ths - c->thr) * log((-c->psib) / (nwt - nf)) -
           c->ths * c->psib;
    else
      dm = c->thr * (nwt + c->psib - nf) - c->ths * c->psib -
           (c->ths - c->thr) / (c->lam - 1.0) * (c->psib + (nwt - nf) * pow((c->psib / (nf - nwt)), c->lam));
  } else
    dm = c->ths * (nwt - nf);
  return dm;
}

/* 2324-2342: point moisture in the initial profile (uses Old water table) / in the wedge */
static double init_moist_z(const cw *c, double nf) {
  if (nf >= (c->nwt0 + c->psib)) return c->ths;
  return (c->thr + (c->ths - c->thr) * pow((c->psib / (nf - c->nwt0)), c->lam));
}
static double edge_moist_z(const cw *c, double nf) {
  return (c->thr + exp(c->f * nf / c->eps) * (c->ths - c->thr) * pow((c->ru1 / c->ks), (1.0 / c->eps)));
}

/* 2353-2364: Green-Ampt capillary drive */
static void suction(cw *c, double nf) {
  if (c->thre == c->ths) {
    c->sein = pow(((c->thri - c->thr) / (c->ths - c->thr)), (3. + 1. / (c->lam * exp(-c->f * nf / 2.))));
    c->g = -c->psib * (1. - c->sein) / (3 * c->lam * exp(-c->f * nf / 2.) + 1.);
  } else {
    c->sein = pow(((c->thri - c->thr) / (c->ths - c->thr)), (3. + 1. / (c->lam * exp(-c->f * nf / 2.))));
    c->se0 = pow(((c->thre - c->thr) / (c->ths - c->thr)), (3. + 1. / (c->lam * exp(-c->f * nf / 2.))));
    c->g = -c->psib * (c->se0 - c->sein) / (3.0 * c->lam * exp(-c->f * nf / 2.) + 1.);
  }
}

/* 2371-2426 */
static double recharge_rate(const cw *c, double dm, double nf) {
  double bb = (dm - c->thr * nf) / (c->ths - c->thr);
  bb = bb * c->f / (c->eps * expm1(c->f * nf / c->eps));
  return (c->ks * pow(bb, c->eps));
}
static double unsat_lat(const cw *c, double nf, double ru, double vel) { return (nf * ru * (c->uar - 1.0) * vel); }
static double sat_lat(const cw *c, double nt, double nf, double ru, double vel) {
  double bb;
  if (nt == 0.0)
    bb = (c->ks * c->uar / c->f * (1. - exp(-c->f * nf)) - c->ks * c->f * nf * nf / expm1(c->f * nf)) * vel;
  else {
    bb = nt * ru * (c->uar - 1.0) + c->ks * c->uar / c->f * (exp(-c->f * nt) - exp(-c->f * nf));
    bb = (bb - c->ks * c->f * (nf - nt) * (nf - nt) / (exp(c->f * nf) - exp(c->f * nt))) * vel;
  }
  return bb;
}
static double transm_fin(const cw *c, double nwt) { return (c->ar * c->ks * (exp(-c->f * nwt) - exp(-c->f * c->dbr)) / c->f); }

/* 3594-3642: mean moisture of the top Z mm from the New state */
static double surf_soil_moist(cw *c, double z) {
  double dm = 0, nstar = 0;
  if (c->nwt1 == 0.0)
    dm = z * c->ths;
  else if ((c->nwt1 > 0.0) && (c->nf1 == 0.0 || c->nf1 == c->nwt1))
    dm = upper_moist(c, z, c->nwt1);
  else if (c->nf1 > 0.0 && c->nt1 == c->nf1 && c->nf1 < c->nwt1) {
    if (fabs(c->ru1) > 1.0E-6) {
      nstar = log(c->ks / c->ru1) / c->f;
      if (c->nf1 >= z)
        dm = c->eps / c->f * (c->ths - c->thr) * (exp(c->f * (z - nstar) / c->eps) - exp(-c->f * nstar / c->eps)) +
             c->thr * z;
      else {
        dm = upper_moist(c, z, c->nwt1);
        dm += (c->mu1 - c->mi1);
      }
    } else
      dm = upper_moist(c, z, c->nwt1);
  } else if (c->nf1 > 0.0 && c->nt1 < c->nf1 && c->nt1 > 0.0) {
    nstar = c->nt1;
    if (c->nt1 >= z)
      dm = c->eps / c->f * (c->ths - c->thr) * (exp(c->f * (z - nstar) / c->eps) - exp(-c->f * nstar / c->eps)) +
           c->thr * z;
    else {
      if (c->nf1 >= z) {
        dm = c->eps / c->f * (c->ths - c->thr) * (1.0 - exp(-c->f * c->nt1 / c->eps)) + c->thr * c->nt1;
        dm += ((z - c->nt1) * c->ths);
      } else if (c->nf1 < z) {
        dm = upper_moist(c, z, c->nwt1);
        dm += (c->mu1 - c->mi1);
      }
    }
  } else if (c->nf1 > 0.0 && c->nt1 == 0.0) {
    if (c->nf1 >= z)
      dm = z * c->ths;
    else if (c->nf1 < z) {
      dm = upper_moist(c, z, c->nwt1);
      dm += (c->mu1 - c->mi1);
    }
  }
  return (dm / z);
}

/* 3654-3686 */
static double z1z2_moist(const cw *c, double z1, double z2, double nwt) {
  double dm;
  if (z2 <= (nwt + c->psib + 1.0E-3)) {
    if (lam_is_one(c))
      dm = c->thr * (z2 - z1) - fabs(c->psib) * (c->ths - c->thr) * log((nwt - z2) / (nwt - z1));
    else
      dm = c->thr * (z2 - z1) - (c->ths - c->thr) / (c->lam - 1.0) *
                                   ((z2 - nwt) * pow((-c->psib / (nwt - z2)), c->lam) -
                                    (z1 - nwt) * pow((-c->psib / (nwt - z1)), c->lam));
  } else if (z2 > (nwt + c->psib + 1.0E-3) && z1 < (nwt + c->psib + 1.0E-3)) {
    dm = z1z2_moist(c, z1, nwt + c->psib, nwt);
    dm += (z2 - (nwt + c->psib)) * c->ths;
  } else
    dm = (z2 - z1) * c->ths;
  return dm;
}

/* ------------------------------------------------------------------------------------
 * Lambert W, principal branch, Halley iteration on a complex iterate
 * (reference: tHydroModel.cpp:3831-3915; complex ops src/Mathutil/mathutil.cpp:1549-1670).
 * Quirks kept: the loop condition's && / || precedence (3893) and the float-typed
 * temporaries inside the complex division (mathutil.cpp:1601-1617).
 * ---------------------------------------------------------------------------------- */
typedef struct { double r, i; } cplx;
static cplx c_sub(cplx a, cplx b) { cplx c; c.r = a.r - b.r; c.i = a.i - b.i; return c; }
static cplx c_mul(cplx a, cplx b) { cplx c; c.r = a.r * b.r - a.i * b.i; c.i = a.i * b.r + a.r * b.i; return c; }
static cplx c_addr(cplx a, double d) { cplx c; c.r = a.r + d; c.i = a.i; return c; }
static cplx c_scale(double x, cplx a) { cplx c; c.r = x * a.r; c.i = x * a.i; return c; }
static cplx c_div(cplx a, cplx b) {
  cplx c;
  float r, den;
  if (fabs(b.r) >= fabs(b.i)) {
    r = b.i / b.r;
    den = b.r + r * b.i;
    c.r = (a.r + r * a.i) / den;
    c.i = (a.i - r * a.r) / den;
  } else {
    r = b.r / b.i;
    den = b.i + r * b.r;
    c.r = (a.r * r + a.i) / den;
    c.i = (a.i * r - a.r) / den;
  }
  return c;
}

static double lambertw(double z) {
  cplx w, v, tt, ttt, t, p;
  double xx, xy = 0.0, c, f;
  int n;
  if (z == 0.0) return 0.0;
  xx = 2.0 * exp(1.0) * z + 2.0;
  if (xx < 0.0) { xx = fabs(xx); xx = sqrt(xx); w.r = -1.0; w.i = xx; }
  else { xx = sqrt(xx) - 1.0; w.r = xx; w.i = 0; }
  xx = fabs(z);
  xx = log(xx);
  if (z < 0.0) { v.r = xx; v.i = REF_PI; } else { v.r = xx; v.i = 0; }
  tt.r = tt.i = 0; /* reference leaves tt unset until assigned; always assigned before use */
  if (z > 1.0) {
    xy = log(xx);
    tt.r = xy; tt.i = 0;
    v = c_sub(v, tt);
  } else if (z == 1.0) {
    tt.r = 0; tt.i = 0;
  } else if (z < 1.0 && z > 0.0) {
    xy = log(fabs(xx));
    tt.r = xy; tt.i = REF_PI;
    v = c_sub(v, tt);
  } else if (z < 0.0 && z > -1.0) {
    xx = atan(v.i / v.r);
    xy = log(fabs(v.r / cos(xx)));
    xx += REF_PI;
    tt.r = xy; tt.i = xx;
    v = c_sub(v, tt);
  }
  c = fabs(z + 1.0 / exp(1.0));
  if (c > 1.45) { c = 1.0; w = v; } else c = 0.0;
  n = 0;
  t.r = 1000; t.i = 1000;
  while (((n < 15) && (fabs(t.r) > xy)) || (fabs(t.i) > xx)) {
    xx = exp(w.r);
    p.r = xx * cos(w.i); p.i = xx * sin(w.i);
    t = c_mul(w, p);
    t.r = t.r - z;
    if (w.i == 0 && w.r == -1) f = 0.0; else f = 1.0;
    tt = c_mul(p, c_addr(w, f));
    ttt = c_sub(tt, c_div(c_mul(t, c_scale(0.5, c_addr(w, 2.0))), c_addr(w, f)));
    t = c_div(c_scale(f, t), ttt);
    w = c_sub(w, t);
    xy = (2.48 * LAMBEPS) * (1.0 + fabs(w.r));
    xx = (2.48 * LAMBEPS) * (1.0 + fabs(w.i));
    n++;
  }
  return w.r;
}

/* ------------------------------------------------------------------------------------
 * Root finders (reference: tHydroModel.cpp:3937-4192)
 * ---------------------------------------------------------------------------------- */
static void polyn1(const cw *c, double x, double *fv, double *dv, double c1, double c2, double c3) {
  *fv = c1 * pow(x, c->lam) + c3 * pow(x, (c->lam - 1.0)) + c2;
  *dv = c1 * c->lam * pow(x, (c->lam - 1.0)) + c3 * (c->lam - 1.0) * pow(x, (c->lam - 2.0));
}
static void polyn2(const cw *c, double x, double nwt, double *fv, double *dv, double c1, double c2, double c3) {
  *fv = c1 * x * pow((nwt - x), (c->lam - 1.0)) + c3 * pow((nwt - x), (c->lam - 1.0)) - c2;
  *dv = c1 * pow((nwt - x), (c->lam - 1.0)) - c1 * (c->lam - 1.0) * x * pow((nwt - x), (c->lam - 2.0)) -
        c3 * (c->lam - 1.0) * pow((nwt - x), (c->lam - 2.0));
}

static double rtsafe_mod(const cw *c, double c1, double c2, double c3, double x1, double x2, double xacc,
                         double xguess) {
  int j;
  double df = 0.0, dx, dxold, f, fh = 0.0, fl = 0.0, temp, xh, xl, rts;
  polyn1(c, x1, &fl, &df, c1, c2, c3);
  polyn1(c, x2, &fh, &df, c1, c2, c3);
  if ((fl > 0.0 && fh > 0.0) || (fl < 0.0 && fh < 0.0)) return xguess;
  if (fl == 0.0) return x1;
  if (fh == 0.0) return x2;
  if (fl < 0.0) { xl = x1; xh = x2; } else { xh = x1; xl = x2; }
  rts = xguess;
  dxold = fabs(x2 - x1);
  dx = dxold;
  polyn1(c, rts, &f, &df, c1, c2, c3);
  for (j = 1; j <= 30; j++) {
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
    polyn1(c, rts, &f, &df, c1, c2, c3);
    if (f < 0.0) xl = rts; else xh = rts;
  }
  return xguess;
}

/* Newton #1: water-table depth from a moisture deficit (3937-4020), non-verbose path */
static double newton1(cw *c, double dm, double nwt) {
  const int ITMAX = 30;
  const double DMIN = 1.0E-5, DX_TOL = 1.0E-5;
  int i;
  double c1, c2, c3, fvalue = 0, fdvalue = 0, x = 0, dx = 0, xinit, xup, est = 0;
  c1 = c->ths - c->thr;
  c2 = c1 * pow((-c->psib), c->lam) / (c->lam - 1.0);
  c3 = c1 * c->psib / (c->lam - 1.0) + c->psib * c1 - dm;
  if (nwt == 0.0) {
    xinit = dm * (c->lam - 1.0) / ((c->ths - c->thr) * c->lam) - c->psib;
    xinit += 50.0;
    if (xinit <= fabs(c->psib)) xinit = fabs(c->psib) + 50.0;
  } else
    xinit = nwt;
  if (c->lam < 5.0) {
    if ((c->ths * nwt - total_moist(c, nwt)) > dm)
      xup = ceil(nwt);
    else
      xup = c->dbr + 1.0;
    x = rtsafe_mod(c, c1, c2, c3, dm / (c->ths - c->thr) + fabs(c->psib), xup, DX_TOL, xinit);
  }
  if (c->lam >= 5.0 || fabs(x - xinit) <= 1.0E-7) {
    for (x = xinit, i = 0; i < ITMAX; x -= dx, i++) {
      polyn1(c, x, &fvalue, &fdvalue, c1, c2, c3);
      if (fabs(fdvalue) < DMIN) { est = x; break; }
      dx = fvalue / fdvalue;
      if (fabs(dx) < DX_TOL) { est = x; break; }
    }
    if ((fabs(fdvalue) < DMIN) || (fabs(dx) < DX_TOL))
      x = est;
    else
      x = nwt;
  }
  if (x > c->dbr) x = c->dbr;
  return x;
}

/* Newton #2: relocate the wetting front so that dM is removed/added (4037-4082) */
static double newton2(const cw *c, double dm, double nwt, double nf, int flag) {
  const int ITMAX = 30;
  const double DX_TOL = 1.0E-5;
  int i;
  double c1, c2, c3, fvalue, fdvalue, x, dx = 0, xinit;
  if (!flag) c1 = c->ths - c->thr; else c1 = -(c->ths - c->thr);
  c2 = c1 * pow((-c->psib), c->lam) / (c->lam - 1);
  c3 = c2 / pow((nwt - nf), (c->lam - 1)) - c1 * nf + dm;
  xinit = nf;
  for (x = xinit, i = 0; i < ITMAX; x -= dx, i++) {
    polyn2(c, x, nwt, &fvalue, &fdvalue, c1, c2, c3);
    /* small-derivative early return only happens in verbose mode (4059-4065) */
    dx = fvalue / fdvalue;
    if (fabs(dx) < DX_TOL) return x;
  }
  return nf;
}

/* ------------------------------------------------------------------------------------
 * helpers shared by both zones
 * ---------------------------------------------------------------------------------- */
static void load_soil(const orc_basin *b, int i, cw *c) {
  c->ks = b->sf[ORC_Ks][i];
  c->ths = b->sf[ORC_ThetaS][i];
  c->thr = b->sf[ORC_ThetaR][i];
  c->lam = b->sf[ORC_PoreSize][i];
  c->psib = b->sf[ORC_AirEBubPres][i];
  c->f = b->sf[ORC_DecayF][i];
  c->ar = b->sf[ORC_SatAnRatio][i];
  c->uar = b->sf[ORC_UnsatAnRatio][i];
  c->eps = 3 + 2 / c->lam;
}
static void load_old(const orc_state *s, int i, cw *c) {
  c->nwt0 = c->nwt1 = s->f[ORCS_NwtOld][i];
  c->mu0 = c->mu1 = s->f[ORCS_MuOld][i];
  c->mi0 = c->mi1 = s->f[ORCS_MiOld][i];
  c->nt0 = c->nt1 = s->f[ORCS_NtOld][i];
  c->nf0 = c->nf1 = s->f[ORCS_NfOld][i];
  c->ru0 = c->ru1 = s->f[ORCS_RuOld][i];
  c->ri0 = c->ri1 = s->f[ORCS_RiOld][i];
}
static void store_new(orc_state *s, int i, const cw *c) {
  s->f[ORCS_NwtNew][i] = c->nwt1;
  s->f[ORCS_MuNew][i] = c->mu1;
  s->f[ORCS_MiNew][i] = c->mi1;
  s->f[ORCS_NfNew][i] = c->nf1;
  s->f[ORCS_NtNew][i] = c->nt1;
  s->f[ORCS_RiNew][i] = c->ri1;
  s->f[ORCS_RuNew][i] = c->ru1;
}

/* soil-moisture diagnostics written by both zones (2034-2057 and 3476-3501) */
static void moisture_diag(const orc_basin *b, orc_state *s, int i, cw *c, double current_time, double *thsurf_io) {
  double thsurf, aa, bb, md;
  thsurf = surf_soil_moist(c, b->surfaceSoilDepth);
  s->f[ORCS_SoilMoisture][i] = thsurf;
  s->f[ORCS_SoilMoistureSC][i] = thsurf / c->ths;
  if (c->nwt1)
    s->f[ORCS_SoilMoistureUNSC][i] = c->mu1 / c->nwt1 / c->ths;
  else
    s->f[ORCS_SoilMoistureUNSC][i] = 1.0;
  aa = (double)((int)(current_time / b->tstep)); /* tRunTimer::getElapsedSteps, tRunTimer.cpp:436 */
  bb = floor(s->f[ORCS_AvSoilMoisture][i]) * 1.0E-4;
  md = (s->f[ORCS_AvSoilMoisture][i] - floor(s->f[ORCS_AvSoilMoisture][i])) * 1.0E+1;
  s->f[ORCS_AvSoilMoisture][i] = floor((bb * aa + thsurf / c->ths) / (aa + 1) * 1.0E+4);
  thsurf = surf_soil_moist(c, b->rootZoneDepth);
  s->f[ORCS_RootMoisture][i] = thsurf;
  s->f[ORCS_RootMoistureSC][i] = thsurf / c->ths;
  s->f[ORCS_AvSoilMoisture][i] += (md * aa + thsurf / c->ths) / (aa + 1.0) * 1.0E-1;
  *thsurf_io = thsurf;
}

/* ------------------------------------------------------------------------------------
 * Unsaturated zone sweep (reference: tHydroModel::UnSaturatedZone, tHydroModel.cpp:720-2223,
 * SetupNodeUSZ 652-703). OPTRUNON is not restated (off by default).
 * ---------------------------------------------------------------------------------- */
static long long g_state_counts[16];
void orc_state_counts(long long *out, int reset) {
  int k;
  for (k = 0; k < 16; k++) { out[k] = g_state_counts[k]; if (reset) g_state_counts[k] = 0; }
}

enum { ST_STORM_EVOL, ST_WT_STAYS, ST_WT_DROPS, ST_WT_GETS, ST_STORM_UNSAT, ST_PERCHED,
       ST_PERCHED_SURFSAT, ST_STORM2INTER, ST_EXACT_INIT, ST_INTSTORM_BELOW };

void orc_unsat(const orc_basin *b, orc_state *s, double dt, double current_time) {
  cw cc, *c = &cc;
  /* these persist across cells in the reference (declared outside its while loop) */
  double kunsat = 0, ractual = 0, xxsrf = 0, mperch = 0, mdelt = 0, mdva = 0, aa = 0, bb = 0;
  double qn = 0, nstar = 0, nwtnext = 0, nfnext = 0, thsurf = 0, evsoi, evveg;
  int i, state;
  int hourly_reset = !(fmod((current_time - b->tstep), b->etistep));
  memset(c, 0, sizeof *c);

  for (i = 0; i < b->n; i++) {
    double varea = b->sf[ORC_varea][i];
    double alpha, vel_mm, qpout0;
    int dst = b->si[ORCI_flow_dest][i];

    /* SetupNodeUSZ */
    load_soil(b, i, c);
    load_old(s, i, c);
    c->qpout = s->f[ORCS_Qpout][i];
    c->qpin = s->f[ORCS_Qpin][i];
    c->qpout = c->qpout * 1.0E-6 / varea;
    c->qpin = c->qpin * 1.0E-6 / varea;
    c->isv = s->f[ORCS_intstorm][i];
    if (hourly_reset) s->f[ORCS_srf_hr][i] = 0.0;

    c->srf = c->hsrf = c->esrf = c->psrf = c->satsrf = c->sbsrf = 0.0;
    alpha = atan(b->sf[ORC_flow_slope][i]);
    if (alpha > 0.0) { c->cosv = fabs(cos(alpha)); c->sinv = fabs(sin(alpha)); }
    else { c->cosv = 1.0; c->sinv = 0.0; }
    c->dbr = b->sf[ORC_bedrock][i];
    vel_mm = b->sf[ORC_flow_vedglen][i] * 1000.;

    evsoi = s->f[ORCS_EvapSoil][i];
    evveg = s->f[ORCS_EvapDryCanopy][i];
    if (evsoi < 0.0) evsoi = 0.0;
    if (evveg < 0.0) evveg = 0.0;
    if (b->Ioption == 0 && b->EToption == 0)
      ractual = s->f[ORCS_Rain][i];
    else if (b->Ioption == 0 && b->EToption != 0)
      ractual = s->f[ORCS_Rain][i] - evsoi - evveg;
    else if (b->Ioption != 0 && b->EToption == 0) {
      if (b->Ioption != 1) ractual = s->f[ORCS_Rain][i];
      else ractual = s->f[ORCS_NetPrecipitation][i];
    } else
      ractual = s->f[ORCS_NetPrecipitation][i] - evsoi - evveg;

    s->glob[ORCG_TotRain] += (ractual * dt * varea / 1000.);

    if (b->SnOpt) {
      double snwe = s->f[ORCS_LiqWE][i] + s->f[ORCS_IceWE][i];
      double routewe = s->f[ORCS_LiqRouted][i];
      if ((snwe > 1e-3) || (routewe > 0.)) ractual = 10.0 * routewe - evveg;
    }

    c->r1 = (ractual + (c->qpin - c->qpout)) * c->cosv + 0.0;
    c->r = ractual;

    if (c->ks != 0.0) thsurf = init_moist_z(c, 0.);
    if (c->ru0 == 0.0)
      kunsat = c->ks * pow(((thsurf - c->thr) / (c->ths - c->thr)), c->eps);
    else
      kunsat = c->ru0;

    /* classification, 876-965; later tests override earlier ones */
    state = -1000;
    if ((c->nwt0 == 0.) && (c->r1 >= 0.)) state = ST_WT_STAYS;
    if ((c->nwt0 == 0.) && (c->r1 < 0.)) state = ST_WT_DROPS;
    if ((c->nwt0 > 0.) && ((c->r1 * dt) >= (c->nwt0 * c->ths - c->mu0)) && (c->nf0 == 0. || c->nf0 == c->nwt0))
      state = ST_WT_GETS;
    if (state == -1000) {
      c->mu1 = c->mu0 + dt * c->r1;
      if (c->mu1 < c->mi0) state = ST_INTSTORM_BELOW;
      if (fabs(c->mu1 - c->mi0) < 1.0E-6) state = ST_EXACT_INIT;
      if (c->mu1 > c->mi0) {
        if (c->r1 > 0.0) {
          if (c->r > RADAR_MM) {
            if (c->isv >= b->IntStormMAX && c->nf0 > 0.0) {
              if (c->mu1 >= c->nwt0 * c->ths) {
                c->nwt1 = 0.0;
                c->mi1 = 0.0;
                c->sbsrf = (c->mu1 - c->nwt0 * c->ths) / dt;
              } else {
                mdelt = c->ths * c->nwt0 - c->mu0;
                c->nwt1 = newton1(c, mdelt, c->nwt0);
                c->mi1 = total_moist(c, c->nwt1);
              }
              c->nwt0 = c->nwt1;
              c->mi0 = c->mu0 = c->mu1 = c->mi1;
              c->nt0 = c->nt1 = 0.0;
              c->nf0 = c->nf1 = 0.0;
              c->ri0 = c->ri1 = 0.0;
              c->ru0 = c->ru1 = 0.0;
              c->isv = 0.0;
            }
          }
          if (c->nt0 == c->nf0) state = ST_STORM_UNSAT;
          if ((c->nt0 < c->nf0) && (c->nt0 != 0.0)) state = ST_PERCHED;
          if ((c->nt0 == 0.0) && (c->nf0 > 0.0)) {
            c->thri = init_moist_z(c, c->nf0);
            c->thre = c->ths;
            suction(c, c->nf0);
            qn = c->ks * c->f * c->nf0 / expm1(c->f * c->nf0) * (c->cosv + c->g / c->nf0);
            xxsrf = qn - c->ri0 * c->cosv;
            if (c->r1 >= xxsrf) state = ST_PERCHED_SURFSAT;
            else state = ST_STORM2INTER;
          }
          if ((c->nf0 == c->nwt0) || (c->r <= RADAR_MM && c->isv > 0 && c->nf0 == 0.0)) state = ST_STORM_EVOL;
        }
        if (c->r1 <= 0.0) state = ST_STORM2INTER;
      }
      if (c->ks == 0.0) state = ST_PERCHED_SURFSAT;
    }

    if (state >= 0) g_state_counts[state]++;
    switch (state) {
    case ST_WT_STAYS: /* 977-1007 */
      if (c->r > 0.0) c->sbsrf = c->r * c->cosv;
      if (c->qpin > c->qpout) c->psrf = (c->qpin - c->qpout) * c->cosv;
      else c->sbsrf -= (c->qpout - c->qpin) * c->cosv;
      c->mi1 = c->mu1 = c->ri1 = c->ru1 = c->nf1 = c->nt1 = c->nwt1 = 0.0;
      c->qpout = 0.0;
      if (c->r <= RADAR_MM) c->isv += dt;
      else if (c->isv > 0.0) c->isv = 0.0;
      break;

    case ST_WT_GETS: /* 1011-1040 */
      if (c->qpin > c->qpout) {
        c->psrf = (c->qpin - c->qpout) * c->cosv + c->mu0 / dt - c->nwt0 * c->ths / dt;
        if (c->psrf < 0.0) { c->sbsrf = c->r * c->cosv + c->psrf; c->psrf = 0.0; }
        else if (c->r > 0.0) c->sbsrf = c->r * c->cosv;
      } else
        c->sbsrf = (c->r + (c->qpin - c->qpout)) * c->cosv + c->mu0 / dt - c->nwt0 * c->ths / dt;
      c->mu1 = c->ri1 = c->ru1 = c->nf1 = c->nt1 = c->nwt1 = c->mi1 = 0.0;
      c->qpout = 0.0;
      if (c->isv > 0.0) c->isv = 0.0;
      break;

    case ST_WT_DROPS: /* 1044-1066 */
      c->r1 = (c->r + (c->qpin - c->qpout)) * c->cosv;
      if (c->r1 < 0.0) c->nwt1 = newton1(c, fabs(c->r1 * dt), 0.0);
      c->mi1 = total_moist(c, c->nwt1);
      c->mu1 = c->mi1;
      c->ri1 = 0.0;
      c->ru1 = 0.0;
      c->qpout = 0.0;
      c->isv += dt;
      if (c->isv >= b->IntStormMAX) { c->nf1 = 0.0; c->nt1 = 0.0; }
      else { c->nf1 = c->nwt1; c->nt1 = c->nwt1; }
      break;

    case ST_EXACT_INIT: /* 1070-1098 */
      c->mi1 = c->mi0;
      c->mu1 = c->mi1;
      c->nwt1 = c->nwt0;
      c->ri1 = 0.0;
      c->ru1 = 0.0;
      c->qpout = 0.0;
      if (c->r <= 0.0) c->isv += dt;
      if (c->nf0 == c->nwt0) {
        if (c->isv >= b->IntStormMAX) { c->nf1 = 0.0; c->nt1 = 0.0; }
        else { c->nf1 = c->nwt1; c->nt1 = c->nwt1; }
      } else { c->nf1 = 0.0; c->nt1 = 0.0; }
      break;

    case ST_INTSTORM_BELOW: /* 1102-1135 */
      c->r1 = (c->r + (c->qpin - c->qpout)) * c->cosv;
      mdelt = c->ths * c->nwt0 - (c->mu0 + c->r1 * dt);
      c->nwt1 = newton1(c, mdelt, c->nwt0);
      c->ri1 = 0.0;
      c->ru1 = 0.0;
      c->mi1 = total_moist(c, c->nwt1);
      c->mu1 = c->mi1;
      c->isv += dt;
      if ((c->nf0 < c->nwt0) && (c->nf0 > 0.0)) {
        c->nf1 = c->nt1 = 0.0;
      } else {
        if (c->isv < b->IntStormMAX) {
          if (c->nf0 == 0.0) c->nf1 = c->nt1 = 0.0;
          else c->nf1 = c->nt1 = c->nwt1;
        } else
          c->nf1 = c->nt1 = 0.0;
      }
      c->qpout = 0.0;
      break;

    case ST_STORM2INTER: /* 1139-1381 */
      if (c->r <= RADAR_MM) c->isv += dt; else c->isv = 0.0;
      c->r1 = (c->r + (c->qpin - c->qpout)) * c->cosv;
      c->mu1 = c->mu0 + c->r1 * dt;
      c->nwt1 = c->nwt0;
      c->mi1 = c->mi0;
      c->ri1 = c->ri0;
      if (c->isv >= b->IntStormMAX && b->RdstrOption) {
        if (c->mu1 >= c->nwt0 * c->ths) {
          c->nwt1 = 0.0;
          c->sbsrf = (c->mu1 - c->nwt0 * c->ths) / dt;
          c->mi1 = 0.0;
        } else {
          mdelt = c->ths * c->nwt0 - c->mu1;
          c->nwt1 = newton1(c, mdelt, c->nwt0);
          c->mi1 = total_moist(c, c->nwt1);
        }
        c->mu1 = c->mi1;
        c->nt1 = 0.0; c->nf1 = 0.0; c->ri1 = 0.0; c->ru1 = 0.0;
      } else {
        mdelt = lower_moist(c, c->nf0, c->nwt1);
        mdelt = c->mu1 - mdelt;
        aa = c->eps / c->f * (c->ths - c->thr) * (1.0 - exp(-c->f * c->nf0 / c->eps)) + c->thr * c->nf0;
        if (mdelt > aa && c->nt0 == c->nf0) mdelt = aa - 1.0E-4;

        if (mdelt <= aa) { /* unsaturated wedge, 1196-1278 */
          if (c->r1 == 0.0) {
            c->ru1 = c->ru0;
            nstar = (log(c->ks / c->ru1)) / c->f;
          } else {
            c->ru1 = recharge_rate(c, mdelt, c->nf0);
            nstar = (log(c->ks / c->ru1)) / c->f;
          }
          if ((100. * (c->mu1 - c->mi0) / c->mi0) <= 0.05) {
            c->nf1 = c->nt1 = 0.0;
            if ((c->mu1 - c->mi0) > 1.0E-4) {
              mdelt = c->ths * c->nwt0 - c->mu1;
              c->nwt1 = newton1(c, mdelt, c->nwt0);
            }
            c->mi1 = total_moist(c, c->nwt1);
            c->mu1 = c->mi1;
            c->ri1 = c->ru1 = 0.0;
          } else {
            c->thri = init_moist_z(c, c->nf0);
            c->thre = edge_moist_z(c, c->nf0);
            if (c->thre <= c->thri) {
              mdelt = c->ths * c->nwt0 - (c->mu0 + c->r1 * dt);
              c->nwt1 = newton1(c, mdelt, c->nwt0);
              c->nt1 = c->nf1 = 0.0;
              c->ru1 = c->ri1 = 0.0;
              c->mi1 = total_moist(c, c->nwt1);
              c->mu1 = c->mi1;
              c->qpout = 0.0;
            } else {
              suction(c, c->nf0);
              qn = c->ru1 * c->cosv + c->ks * exp(-c->f * c->nf0) * c->g / c->nf0;
              c->nf1 = c->nf0 + dt * (qn - c->ri1 * c->cosv) / (c->thre - c->thri);
              c->nt1 = c->nf1;
              if (c->nf1 < (c->nwt1 + c->psib)) {
                mdelt = lower_moist(c, c->nf1, c->nwt1);
                mdelt = c->mu1 - mdelt;
                c->ru1 = recharge_rate(c, mdelt, c->nf1);
                aa = c->eps / c->f * (c->ths - c->thr) * (1.0 - exp(-c->f * c->nf1 / c->eps)) + c->thr * c->nf1;
                if (mdelt > aa) {
                  c->nt1 = (log(c->ks / c->ru1)) / c->f;
                  c->ru1 = c->ks * exp(-c->f * c->nt1);
                }
              } else if (c->nf1 >= (c->nwt1 + c->psib)) {
                if ((c->mu1 - c->mi0) > 1.0E-4) {
                  mdelt = c->ths * c->nwt0 - c->mu1;
                  c->nwt1 = newton1(c, mdelt, c->nwt0);
                }
                c->mi1 = total_moist(c, c->nwt1);
                c->mu1 = c->mi1;
                c->ri1 = c->ru1 = 0.0;
                /* downslope cell's Old values (1266-1273); the outlet node is never swept,
                   its Old values stay at their constructor zeros */
                nwtnext = dst >= 0 ? s->f[ORCS_NwtOld][dst] : 0.0;
                nfnext = dst >= 0 ? s->f[ORCS_NfOld][dst] : 0.0;
                if ((nfnext == nwtnext) && (c->isv < b->IntStormMAX)) c->nf1 = c->nt1 = c->nwt1;
                else c->nf1 = c->nt1 = 0.0;
                c->ru1 = 0.0;
              }
            }
          }
        } else { /* perched / surface-saturated wedge, 1283-1369 */
          c->thri = init_moist_z(c, c->nf0);
          c->thre = c->ths;
          suction(c, c->nf0);
          qn = c->ks * c->f * (c->nf0 - c->nt0) / (exp(c->f * c->nf0) - exp(c->f * c->nt0));
          qn *= (c->cosv + c->g / c->nf0);
          c->nf1 = c->nf0 + dt * (qn - c->ri1 * c->cosv) / (c->ths - c->thri);
          if (fabs(c->nf1 - (c->nwt1 + c->psib)) <= 1.0E-3) {
            c->nwt1 = c->nt0 - c->psib;
            nwtnext = dst >= 0 ? s->f[ORCS_NwtOld][dst] : 0.0;
            nfnext = dst >= 0 ? s->f[ORCS_NfOld][dst] : 0.0;
            if ((nfnext == nwtnext) && (c->isv < b->IntStormMAX)) c->nf1 = c->nwt1;
            else c->nf1 = 0.0;
          } else if (c->nf1 > (c->nwt1 + c->psib)) {
            c->nwt1 = c->nt0 - c->psib;
            nwtnext = dst >= 0 ? s->f[ORCS_NwtOld][dst] : 0.0;
            nfnext = dst >= 0 ? s->f[ORCS_NfOld][dst] : 0.0;
            if ((nfnext == nwtnext) && (c->isv < b->IntStormMAX)) c->nf1 = c->nwt1;
            else c->nf1 = 0;
          }
          if ((c->nf1 == 0.0) || (c->nf1 == c->nwt1)) {
            if (c->mu1 >= c->nwt0 * c->ths) {
              c->nwt1 = 0.0;
              c->nt1 = c->nf1 = 0.0;
              c->sbsrf = (c->mu1 - c->nwt0 * c->ths) / dt;
              c->mu1 = c->mi1 = 0.0;
              c->ru1 = c->ri1 = 0.0;
            } else {
              mdelt = c->ths * c->nwt0 - c->mu1;
              c->nwt1 = newton1(c, mdelt, c->nwt0);
              c->ri1 = c->ru1 = 0.0;
              c->mi1 = total_moist(c, c->nwt1);
              c->mu1 = c->mi1;
              if (c->nf1 == c->nwt1) c->nf1 = c->nt1 = c->nwt1;
              else c->nf1 = c->nt1 = 0.0;
            }
          } else if ((c->nf1 > 0.0) && (c->nf1 != c->nwt1)) {
            mdelt = c->eps / c->f * (c->ths - c->thr) * (1.0 - exp(-c->f * c->nt0 / c->eps)) + c->thr * c->nt0;
            mdva = c->eps / c->f * (c->ths - c->thr) * (1.0 - exp(-c->f * c->nf1 / c->eps)) + c->thr * c->nf1;
            aa = (mdelt + (c->nf1 - c->nt0) * c->ths - mdva);
            if (c->r1 < (qn - aa / dt)) {
              mperch = lower_moist(c, c->nf1, c->nwt1);
              mperch = c->mu1 - mperch;
              c->ru1 = recharge_rate(c, mperch, c->nf1);
              c->nt1 = c->nf1;
              if (mdelt >= mdva) c->mu1 -= mdelt - mdva + 1.0E-4;
            } else {
              bb = -dt * (qn - c->r1) / (c->ths - c->thr) - c->eps / c->f * exp(-c->f * c->nt0 / c->eps);
              aa = -exp(c->f * (bb - c->nt0) / c->eps);
              c->nt1 = c->nt0 + (c->eps / c->f * lambertw(aa) - bb);
              c->ru1 = c->ks * exp(-c->f * c->nt1);
            }
          }
        }
        c->qpout = 0.0;
        if ((c->nt1 == c->nf1) && (c->nf1 != c->nwt1) && (c->nf1 > 10.))
          c->qpout = unsat_lat(c, c->nf1, c->ru1, vel_mm);
        if ((c->nf1 - c->nt1) > 1.0) c->qpout = sat_lat(c, c->nt1, c->nf1, c->ru1, vel_mm);
        c->qpout *= c->sinv;
      }
      break;

    case ST_STORM_EVOL: /* 1385-1413 */
      c->r1 = (c->r + (c->qpin - c->qpout)) * c->cosv;
      mdelt = c->ths * c->nwt0 - (c->mu0 + c->r1 * dt);
      c->nwt1 = newton1(c, mdelt, c->nwt0);
      c->ri1 = c->ru1 = 0.0;
      c->mi1 = total_moist(c, c->nwt1);
      c->mu1 = c->mi1;
      c->qpout = 0.0;
      if (c->r <= RADAR_MM) {
        c->isv += dt;
        if (c->nf0 == c->nwt0) {
          if (c->isv >= b->IntStormMAX) c->nf1 = c->nt1 = 0.0;
          else c->nf1 = c->nt1 = c->nwt1;
        } else
          c->nf1 = c->nt1 = 0.0;
      } else {
        if (c->isv > 0.0) c->isv = 0.0;
        c->nf1 = c->nt1 = c->nwt1;
      }
      break;

    case ST_STORM_UNSAT: /* 1417-1666 */
      if (c->r <= RADAR_MM) c->isv += dt; else c->isv = 0.0;
      c->r1 = (c->r + (c->qpin - c->qpout)) * c->cosv;
      c->nwt1 = c->nwt0;
      c->mu1 = c->mu0 + c->r1 * dt;
      c->mi1 = c->mi0;
      c->ri1 = c->ri0 = 0.0;
      if (c->nf0 == 0.0 && c->nt0 == 0.0) { /* no front yet, 1433-1523 */
        c->thri = init_moist_z(c, 0.0);
        if (c->r < kunsat && c->r1 / c->cosv < kunsat) {
          c->thre = pow((c->r1 / c->ks), (1 / c->eps)) * (c->ths - c->thr) + c->thr;
          qn = c->r1;
          c->nf1 = dt * qn / (c->thre - c->thri);
        } else {
          if (c->r1 < c->ks) {
            c->thre = c->thr + (c->ths - c->thr) * pow((c->r1 / c->ks), (1 / c->eps));
            suction(c, 0.0);
            if (fabs(c->thre - c->thri) < 1.0E-4) qn = c->r1;
            else qn = c->ks * exp(-c->f * c->r1 / (c->thre - c->thri)) * c->g * (c->thre - c->thri) / c->r1 + c->r1;
            c->nf1 = dt * qn / (c->ths - c->thri);
          } else {
            c->thre = c->ths;
            suction(c, 0.0);
            qn = c->ks * c->g * (c->thre - c->thri) / c->ks + c->r1;
            c->nf1 = dt * qn / (c->thre - c->thri);
          }
        }
        if (c->nf1 > 0.0 && c->nf1 < 32000.) {
          if (c->nf1 < 1.0) c->nf1 = 1.0;
          c->nt1 = c->nf1;
        } else
          c->nt1 = c->nf1 = c->nwt0 + 100.;

        if (c->nf1 < c->nwt0) {
          mdelt = lower_moist(c, c->nf1, c->nwt1);
          mdelt = c->mu1 - mdelt;
          aa = c->eps / c->f * (c->ths - c->thr) * (1.0 - exp(-c->f * c->nf1 / c->eps)) + c->thr * c->nf1;
          if (mdelt >= aa) {
            if (fabs(mdelt - aa) <= 1.0E-6) {
              c->nt1 = c->nf1;
              c->nf1 += 1.0E-5;
              c->ru1 = c->ks * exp(-c->f * c->nt1);
            } else {
              if (mdelt >= c->nf1 * c->ths) {
                c->thri = init_moist_z(c, c->nf1);
                c->nf1 += (mdelt - c->nf1 * c->ths) / (c->ths - c->thri);
                if (c->nf1 >= (c->nwt1 + c->psib)) {
                  c->nf1 = c->nt1 = c->nwt1 = 0.0;
                  c->mu1 = c->mi1 = 0.0;
                  c->ru1 = c->ri1 = 0.0;
                } else {
                  c->nt1 = 0.0;
                  c->ru1 = c->ks * c->f * c->nf1 / expm1(c->f * c->nf1);
                  aa = lower_moist(c, c->nf1, c->nwt1);
                  bb = c->nf1 * c->ths;
                  c->mu1 = aa + bb;
                }
              } else if (mdelt < c->nf1 * c->ths) {
                c->ru1 = recharge_rate(c, mdelt, c->nf1);
                c->nt1 = (log(c->ks / c->ru1)) / c->f;
                c->nf1 += 1.0E-5;
              }
            }
          } else
            c->ru1 = recharge_rate(c, mdelt, c->nf1);
          if ((c->ri1 - c->ru1) < 1.0E-2 && (c->ri1 - c->ru1) > 0.0) c->ru1 = c->ri1;
        }
      } else if (c->nf0 > 0.0 && c->nt0 > 0.0) { /* fronts below the surface, 1528-1645 */
        mdelt = lower_moist(c, c->nf0, c->nwt1);
        mdelt = c->mu1 - mdelt;
        if (mdelt >= c->nf0 * c->ths) {
          c->thri = init_moist_z(c, c->nf0);
          c->thre = c->ths;
          suction(c, c->nf0);
          qn = c->ks * c->f * c->nf0 / expm1(c->f * c->nf0);
          if ((qn * (c->cosv + c->g / c->nf0) - c->ri0 * c->cosv) <= c->r1) {
            qn *= (c->cosv + c->g / c->nf0);
            if (c->r1 >= (qn - c->ri0 * c->cosv)) {
              c->hsrf += (c->r1 - (qn - c->ri0 * c->cosv));
              c->r1 -= c->hsrf;
            }
            c->nt1 = 0.0;
            c->mu1 = c->mu0 + c->r1 * dt;
            c->nf1 = c->nf0 + dt * (qn - c->ri1 * c->cosv) / (c->ths - c->thri);
            c->ru1 = c->ks * c->f * c->nf1 / expm1(c->f * c->nf1);
            if ((fabs(c->nf1 - (c->nwt1 + c->psib)) <= 1.0E-3) || (c->nf1 > (c->nwt1 + c->psib))) {
              c->nwt1 = c->mi1 = 0.0;
              c->ru1 = c->ri1 = 0.0;
              mperch = c->mu1 - c->nwt0 * c->ths;
              if (mperch > 0.0) c->sbsrf += mperch / dt;
              c->mu1 = 0.0;
              c->nt1 = c->nf1 = 0.0;
            }
          } else {
            c->thre = edge_moist_z(c, c->nf0);
            suction(c, c->nf0);
            qn = c->ru0 * c->cosv + c->ks * exp(-c->f * c->nf0) * c->g / c->nf0;
            c->nf1 = c->nf0 + dt * (qn - c->ri1 * c->cosv) / (c->ths - c->thri);
            c->nt1 = c->nf1;
            if (c->nf1 < (c->nwt1 + c->psib)) {
              mdelt = lower_moist(c, c->nf1, c->nwt1);
              mdelt = c->mu1 - mdelt;
              c->ru1 = recharge_rate(c, mdelt, c->nf1);
            } else {
              if (c->mu1 > c->ths * c->nwt1) {
                c->sbsrf += (c->mu1 - c->ths * c->nwt1);
                c->nf1 = c->nt1 = c->nwt1 = 0.0;
                c->mu1 = c->mi1 = 0.0;
                c->ru1 = c->ri1 = 0.0;
              }
            }
          }
        } else {
          c->ru1 = recharge_rate(c, mdelt, c->nf0);
          nstar = (log(c->ks / c->ru1)) / c->f;
          if (nstar <= c->nf0) {
            c->nt1 = nstar;
            c->nf1 = c->nf0 + 1.0E-5;
          } else {
            c->thri = init_moist_z(c, c->nf0);
            c->thre = edge_moist_z(c, c->nf0);
            if (c->thre <= c->thri) {
              mdelt = c->ths * c->nwt0 - (c->mu0 + c->r1 * dt);
              c->nwt1 = newton1(c, mdelt, c->nwt0);
              c->ru1 = c->ri1 = 0.0;
              c->mi1 = total_moist(c, c->nwt1);
              c->mu1 = c->mi1;
              c->qpout = 0.0;
              if (c->nwt0 < 2.0 * fabs(c->psib)) c->nt1 = c->nf1 = c->nwt1;
              else c->nt1 = c->nf1 = 0.0;
            } else {
              suction(c, c->nf0);
              qn = c->ru1 * c->cosv + c->ks * exp(-c->f * c->nf0) * c->g / c->nf0;
              c->nf1 = c->nf0 + dt * (qn - c->ri1 * c->cosv) / (c->thre - c->thri);
              c->nt1 = c->nf1;
              if (c->nf1 < (c->nwt1 + c->psib)) {
                mdelt = lower_moist(c, c->nf1, c->nwt1);
                mdelt = c->mu1 - mdelt;
                c->ru1 = recharge_rate(c, mdelt, c->nf1);
              }
            }
          }
        }
      }
      if ((c->nf1 > (c->nwt1 + c->psib)) && (c->nf1 != c->nwt1)) { /* 1648-1655 */
        mdelt = c->ths * c->nwt1 - c->mu1;
        c->nwt1 = newton1(c, mdelt, c->nwt0);
        c->ri1 = c->ru1 = 0.0;
        c->mu1 = c->mi1 = total_moist(c, c->nwt1);
        c->nf1 = c->nt1 = c->nwt1;
      }
      c->qpout = 0.0;
      if ((c->nt1 == c->nf1) && (c->nf1 != c->nwt1) && (c->nf1 > 10.0))
        c->qpout = unsat_lat(c, c->nf1, c->ru1, vel_mm);
      if ((c->nf1 - c->nt1) > 1.0) c->qpout = sat_lat(c, c->nt1, c->nf1, c->ru1, vel_mm);
      c->qpout *= c->sinv;
      break;

    case ST_PERCHED: /* 1670-1806 */
      if (c->r <= RADAR_MM) c->isv += dt; else c->isv = 0.0;
      c->r1 = (c->r + (c->qpin - c->qpout)) * c->cosv;
      c->nwt1 = c->nwt0;
      c->mu1 = c->mu0 + c->r1 * dt;
      c->mi1 = c->mi0;
      c->ri1 = c->ri0;
      c->thri = init_moist_z(c, c->nf0);
      c->thre = c->ths;
      suction(c, c->nf0);
      qn = c->ks * c->f * (c->nf0 - c->nt0) / (exp(c->f * c->nf0) - exp(c->f * c->nt0));
      qn *= (c->cosv + c->g / c->nf0);
      c->nf1 = c->nf0 + dt * (qn - c->ri1 * c->cosv) / (c->ths - c->thri);
      if (fabs(c->nf1 - (c->nwt1 + c->psib)) <= 1.0E-3) {
        c->nwt1 = c->nt0 - c->psib;
        c->nf1 = c->nwt1;
      } else if (c->nf1 > (c->nwt1 + c->psib)) {
        c->nwt1 = c->nt0 - c->psib;
        c->nf1 = c->nwt1;
        c->r1 += qn - ((c->nwt0 + c->psib - c->nf0) * (c->ths - c->thri) / dt + c->ri1 * c->cosv);
      }
      if (c->nf1 == c->nwt1) {
        if (c->mu1 >= c->nwt0 * c->ths) {
          c->nwt1 = c->nt1 = c->nf1 = 0.0;
          c->sbsrf = (c->mu1 - c->nwt0 * c->ths) / dt;
          c->mu1 = c->mi1 = 0.0;
          c->ru1 = c->ri1 = 0.0;
        } else {
          mdelt = c->ths * c->nwt0 - c->mu1;
          c->nwt1 = newton1(c, mdelt, c->nwt0);
          c->ri1 = c->ru1 = 0.0;
          c->mi1 = total_moist(c, c->nwt1);
          c->mu1 = c->mi1;
          c->nt1 = c->nf1 = c->nwt1;
        }
      } else if (c->nf1 != c->nwt1) {
        mdelt = c->eps / c->f * (c->ths - c->thr) * (1.0 - exp(-c->f * c->nt0 / c->eps)) + c->thr * c->nt0;
        mdva = c->eps / c->f * (c->ths - c->thr) * (1.0 - exp(-c->f * c->nf1 / c->eps)) + c->thr * c->nf1;
        aa = (mdelt + (c->nf1 - c->nt0) * c->ths - mdva);
        if ((c->r1 + c->ri1 * c->cosv) < (qn - aa / dt)) {
          mperch = lower_moist(c, c->nf1, c->nwt1);
          mperch = c->mu1 - mperch;
          c->ru1 = recharge_rate(c, mperch, c->nf1);
          c->nt1 = c->nf1;
        } else {
          xxsrf = (c->r1 - qn) * dt + mdelt;
          if (xxsrf >= c->nt0 * c->ths) {
            c->nt1 = 0.0;
            c->hsrf += (xxsrf - c->nt0 * c->ths) / dt;
            c->mu1 -= c->hsrf * dt;
            c->ru1 = c->ks * c->f * c->nf1 / expm1(c->f * c->nf1);
          } else {
            bb = -dt * (qn - c->r1) / (c->ths - c->thr) - c->eps / c->f * exp(-c->f * c->nt0 / c->eps);
            aa = -exp(c->f * (bb - c->nt0) / c->eps);
            c->nt1 = c->nt0 + (c->eps / c->f * lambertw(aa) - bb);
            c->ru1 = c->ks * exp(-c->f * c->nt1);
          }
          if (c->nwt1 != 0.0) {
            aa = lower_moist(c, c->nf1, c->nwt1);
            bb = (c->nf1 - c->nt1) * c->ths + c->eps / c->f * (c->ths - c->thr) * (1.0 - exp(-c->f * c->nt1 / c->eps)) +
                 c->thr * c->nt1;
            if (c->mu1 > (aa + bb)) {
              c->thri = init_moist_z(c, c->nf1);
              c->nf1 += (c->mu1 - (aa + bb)) / (c->ths - c->thri);
              if (c->nf1 >= (c->nwt1 + c->psib)) {
                if (c->nt1 > 0.0) {
                  c->nwt1 = c->nt1 - c->psib;
                  c->nf1 = c->nt1 = c->nwt1;
                  c->mu1 = c->mi1 = total_moist(c, c->nwt1);
                  c->ri1 = c->ru1 = 0.0;
                } else if (c->nt1 == 0.0) {
                  c->nwt1 = c->nf1 = c->nt1 = 0.0;
                  c->mu1 = c->mi1 = 0.0;
                  c->ri1 = c->ru1 = 0.0;
                }
              } else {
                c->mu1 = aa + bb;
                if (c->nt1 == 0.0) c->ru1 = c->ks * c->f * c->nf1 / expm1(c->f * c->nf1);
              }
            }
          }
        }
      }
      c->qpout = 0.0;
      if ((c->nt1 == c->nf1) && (c->nf1 != c->nwt1) && (c->nf1 > 10.0))
        c->qpout = unsat_lat(c, c->nf1, c->ru1, vel_mm);
      if ((c->nf1 - c->nt1) > 1.0) c->qpout = sat_lat(c, c->nt1, c->nf1, c->ru1, vel_mm);
      c->qpout *= c->sinv;
      break;

    case ST_PERCHED_SURFSAT: /* 1810-1911 */
      if (c->r <= RADAR_MM) c->isv += dt; else c->isv = 0.0;
      if (c->ks == 0.0) {
        c->mu1 = c->mi1 = 0.0;
        c->ri1 = c->ru1 = 0.0;
        c->nf1 = c->nt1 = 0.0;
        c->nwt1 = c->nwt0;
        c->qpout = 0.0;
        c->hsrf += c->r * c->cosv;
      } else {
        c->nwt1 = c->nwt0;
        if ((c->nt0 == 0.0) && (c->nf0 > 0.0) && (c->nf0 != c->nwt0)) {
          c->thri = init_moist_z(c, c->nf0);
          c->thre = c->ths;
          suction(c, c->nf0);
          qn = c->ks * c->f * c->nf0 / expm1(c->f * c->nf0);
          qn *= (c->cosv + c->g / c->nf0);
          c->r1 = (c->r + (c->qpin - c->qpout)) * c->cosv;
          if (c->r1 >= qn) {
            c->hsrf += (c->r1 - qn);
            c->r1 -= c->hsrf;
          }
          c->nt1 = 0.0;
          c->ri1 = c->ri0;
          c->mi1 = c->mi0;
          c->mu1 = c->mu0 + c->r1 * dt;
          c->nf1 = c->nf0 + dt * (qn - c->ri1 * c->cosv) / (c->ths - c->thri);
        }
        if (c->nf1 >= 1.0) {
          c->ru1 = c->ks * c->f * c->nf1 / expm1(c->f * c->nf1);
          if (c->ru1 > c->ks) c->ru1 = c->ks;
        } else
          c->ru1 = c->ks;
        if ((fabs(c->nf1 - (c->nwt1 + c->psib)) <= 1.0E-3) || (c->nf1 > (c->nwt1 + c->psib))) {
          mperch = c->mu1 - c->nwt0 * c->ths;
          if (mperch >= 0.0) {
            c->sbsrf += mperch / dt;
            c->nf1 = c->nt1 = c->nwt1 = 0.0;
            c->mi1 = c->mu1 = 0.0;
            c->ri1 = c->ru1 = 0.0;
          } else {
            c->nwt1 = newton1(c, fabs(mperch), 0.0);
            c->mu1 = c->mi1 = total_moist(c, c->nwt1);
            c->ri1 = c->ru1 = 0.0;
            c->nf1 = c->nt1 = c->nwt1;
          }
        }
        if (c->nwt1 != 0.0 && c->nf1 >= 1.0 && c->nf1 != c->nwt1) {
          aa = lower_moist(c, c->nf1, c->nwt1);
          bb = c->nf1 * c->ths;
          if (c->mu1 >= (aa + bb)) {
            c->thri = init_moist_z(c, c->nf1);
            c->nf1 += (c->mu1 - (aa + bb)) / (c->ths - c->thri);
            if (c->nf1 >= (c->nwt1 + c->psib)) {
              if (c->mu1 >= (c->nwt0 * c->ths)) c->sbsrf += (c->mu1 - c->nwt0 * c->ths);
              c->nwt1 = 0.0;
              c->ri1 = c->ru1 = 0.0;
              c->mi1 = c->mu1 = 0.0;
              c->nf1 = c->nt1 = 0.0;
            } else {
              c->mu1 = aa + bb;
              c->ru1 = c->ks * c->f * c->nf1 / expm1(c->f * c->nf1);
            }
          }
        }
        c->qpout = 0.0;
        if ((c->nf1 - c->nt1) > 10.0) {
          c->qpout = sat_lat(c, c->nt1, c->nf1, c->ru1, vel_mm);
          c->qpout *= c->sinv;
        }
      }
      break;
    default:
      break;
    }

    /* runoff onto a horizontal plane, 1995-2006 */
    c->esrf = c->psrf;
    c->esrf = c->esrf / c->cosv;
    c->srf = c->hsrf + c->sbsrf + c->psrf;
    c->srf /= c->cosv;
    c->hsrf /= c->cosv;
    c->sbsrf /= c->cosv;
    c->psrf /= c->cosv;

    {
      double thsurf0 = s->f[ORCS_SoilMoistureSC][i];
      double throot0 = s->f[ORCS_RootMoistureSC][i];
      double areaf, fs = 0, ft = 0, fc = 0;
      qpout0 = s->f[ORCS_Qpout][i];

      store_new(s, i, c);
      s->f[ORCS_Qpout][i] = c->qpout;
      s->f[ORCS_intstorm][i] = c->isv;
      s->f[ORCS_srf_hr][i] += c->srf * dt;
      s->f[ORCS_srf][i] = c->srf * dt;
      s->f[ORCS_cumsrf][i] += c->srf * dt;
      s->f[ORCS_sbsrf][i] = c->sbsrf * dt;
      s->f[ORCS_hsrf][i] = c->hsrf * dt;
      s->f[ORCS_psrf][i] = c->psrf * dt;
      s->f[ORCS_esrf][i] = c->esrf * dt;

      moisture_diag(b, s, i, c, current_time, &thsurf);

      s->f[ORCS_Recharge][i] = (c->nwt1 - c->nwt0) * c->ths / (c->cosv * dt);
      s->f[ORCS_UnSatFlowOut][i] = c->qpout * 1.0E-6 / varea;
      s->f[ORCS_UnSatFlowIn][i] = c->qpin * 1.0E-6 / varea;

      if (c->hsrf > 0.0) s->f[ORCS_hsrfOccur][i] = s->f[ORCS_hsrfOccur][i] + floor(c->hsrf * 1.0E+3) + 1.0E-6;
      if (c->sbsrf > 0.0) s->f[ORCS_sbsrfOccur][i] = s->f[ORCS_sbsrfOccur][i] + floor(c->sbsrf * 1.0E+3) + 1.0E-6;
      if (c->psrf > 0.0) s->f[ORCS_psrfOccur][i] = s->f[ORCS_psrfOccur][i] + floor(c->psrf * 1.0E+3) + 1.0E-6;
      if (surf_soil_moist(c, 1.0) / c->ths > 0.999) s->f[ORCS_satOccur][i] = s->f[ORCS_satOccur][i] + 1;
      s->f[ORCS_RechDisch][i] =
          s->f[ORCS_RechDisch][i] + ((c->nwt0 - c->nwt1) + c->sbsrf * dt * c->cosv / c->ths) * 1.0E-3;

      /* scatter to the downslope cell, consumed later in this same sweep (2078-2079) */
      if (dst >= 0) s->f[ORCS_Qpin][dst] += c->qpout;

      s->glob[ORCG_Stok] += c->srf * dt * varea / 1000.0;
      s->glob[ORCG_TotGWchange] += (c->nwt1 - c->nwt0) * c->ths * varea / (c->cosv * 1000.0);
      s->glob[ORCG_TotMoist] += (c->mu1 - c->mu0) * varea / (c->cosv * 1000.0);

      /* relative-forcing diagnostics, 2103-2152 */
      qpout0 *= ((1.0E-6) / varea);
      if (c->nwt1 < c->nwt0) {
        fs = fabs(c->mu1 + c->ths * (c->nwt0 - c->nwt1) - c->mu0) / dt - (c->qpin - qpout0) * c->cosv + (evsoi + evveg);
        ft = fabs(c->qpin - qpout0) * c->cosv;
        fc = fabs(evsoi + evveg);
      } else if (c->nwt1 > c->nwt0) {
        if (c->ru0 > 0) fs = c->ru0; else fs = s->f[ORCS_NetPrecipitation][i];
        ft = fabs(c->qpin - qpout0) * c->cosv;
        fc = fabs(evsoi + evveg);
      } else if (c->nwt0 == 0.0 && c->nwt1 == 0.0) {
        ft = (c->qpin - qpout0) * c->cosv;
        fc = fabs(evsoi + evveg);
        if (ft - fc > 0.0) fs = 0.0; else fs = fabs(ft + fc);
        ft = fabs(ft);
      } else if (fabs(c->nwt1 - c->nwt0) < 1.0E-9) {
        fs = c->ru1;
        ft = fabs(c->qpin - qpout0) * c->cosv;
        fc = fabs(evsoi + evveg);
      }
      areaf = varea / b->BasArea;
      fs *= areaf; ft *= areaf; fc *= areaf;
      s->glob[ORCG_fSoi100] += fs;
      s->glob[ORCG_fTop100] += ft;
      s->glob[ORCG_fClm100] += fc;
      s->glob[ORCG_dM100] += (s->f[ORCS_SoilMoistureSC][i] - thsurf0) * areaf * c->ths * 100.;
      s->glob[ORCG_dMRt] += (s->f[ORCS_RootMoistureSC][i] - throot0) * areaf * c->ths * 1000.;
      s->glob[ORCG_mTh100] += (s->f[ORCS_SoilMoistureSC][i] + thsurf0) / 2.0 * areaf;
      s->glob[ORCG_mThRt] += (s->f[ORCS_RootMoistureSC][i] + throot0) / 2.0 * areaf;
    }
    s->glob[ORCG_psrf_last] = c->psrf; /* the member the saturated zone later reads (3439) */
  }
}

/* ------------------------------------------------------------------------------------
 * Saturated zone (reference: ResetGW 2470-2488, ComputeFluxesEdgesND 2663-2823,
 * SaturatedZone 2998-3573, tCNode::getGwaterChng tCNode.cpp:356-370)
 * ---------------------------------------------------------------------------------- */
void orc_reset_gw(const orc_basin *b, orc_state *s) {
  int i;
  for (i = 0; i < b->n; i++) {
    s->f[ORCS_NwtOld][i] = s->f[ORCS_NwtNew][i];
    s->f[ORCS_MuOld][i] = s->f[ORCS_MuNew][i];
    s->f[ORCS_MiOld][i] = s->f[ORCS_MiNew][i];
    s->f[ORCS_NfOld][i] = s->f[ORCS_NfNew][i];
    s->f[ORCS_NtOld][i] = s->f[ORCS_NtNew][i];
    s->f[ORCS_RuOld][i] = s->f[ORCS_RuNew][i];
    s->f[ORCS_RiOld][i] = s->f[ORCS_RiNew][i];
    s->f[ORCS_esrf][i] = 0.0;
    s->f[ORCS_satsrf][i] = 0.0;
    s->f[ORCS_gwchange][i] = 0.0;
  }
}

static int cmp_dbl(const void *a, const void *b) {
  double x = *(const double *)a, y = *(const double *)b;
  return (x > y) - (x < y);
}

/* flux over the directed edge org->dst (2681-2808); returns 1 when the "positive flux" branch
   was taken, *q = QOut, *tr = the transmissivity the reference stores on the origin */
static int edge_flux(const orc_basin *b, const orc_state *s, int org, int dst, double len, double vedglen,
                     double *q, double *tr) {
  double cos1 = cos(atan(b->sf[ORC_flow_slope][org]));
  double cos2 = cos(atan(b->sf[ORC_flow_slope][dst]));
  double nwo = s->f[ORCS_NwtOld][org], nwd = s->f[ORCS_NwtOld][dst];
  double psib, this_e, next_e, dbr, wts, t, width, deficit;
  psib = b->sf[ORC_AirEBubPres][org];
  if (nwo == 0.0) this_e = b->sf[ORC_z][org] - ((-psib) / (cos1 * 1000.0));
  else this_e = b->sf[ORC_z][org] - (nwo / (cos1 * 1000.0));
  psib = b->sf[ORC_AirEBubPres][dst];
  if (nwd == 0.0) next_e = b->sf[ORC_z][dst] - ((-psib) / (cos2 * 1000.0));
  else next_e = b->sf[ORC_z][dst] - (nwd / (cos2 * 1000.0));
  dbr = b->sf[ORC_bedrock][org];
  *q = 0.0;
  if (this_e > next_e && nwo <= dbr && nwd <= dbr) {
    double ks = b->sf[ORC_Ks][org], f = b->sf[ORC_DecayF][org], ar = b->sf[ORC_SatAnRatio][org];
    wts = (this_e - next_e) / len;
    t = (ar * ks * (exp(-f * nwo) - exp(-f * dbr)) / f);
    width = vedglen * 1000.0;
    if (wts > 1.0) wts = 1.0;
    *q = t * width * wts;
    if (width) {
      deficit = b->sf[ORC_varea][dst] / (width * width * 10.0E-6);
      if (deficit <= 0.1) *q *= deficit;
    }
    *tr = t;
    return 1;
  } else {
    double ks = b->sf[ORC_Ks][org], f = b->sf[ORC_DecayF][org], ar = b->sf[ORC_SatAnRatio][org];
    *tr = ar * ks * exp(-f * nwo) / f;
    return 0;
  }
}

void orc_sat(const orc_basin *b, orc_state *s, double dtgw, double current_time) {
  cw cc, *c = &cc;
  int i, k, n = b->n;
  double gwdm = 0.0, mth100 = 0.0, mthrt = 0.0, areagw = 0.0, thsurf = 0.0;
  int gwcnt = 0;
  double *contrib;
  int *cnt, *cap_off;
  memset(c, 0, sizeof *c);

  /* ComputeFluxesEdgesND: every cell receives +Q for each of its outgoing edges and -Q for each
     incoming one; at most 2*degree contributions */
  cap_off = (int *)malloc(sizeof(int) * (size_t)(n + 1));
  cnt = (int *)calloc((size_t)n, sizeof(int));
  cap_off[0] = 0;
  for (i = 0; i < n; i++) cap_off[i + 1] = cap_off[i] + 2 * (b->nbr_rowptr[i + 1] - b->nbr_rowptr[i]);
  contrib = (double *)malloc(sizeof(double) * (size_t)(cap_off[n] + 1));
  for (i = 0; i < n; i++) {
    int best = -1;
    double besttr = 0.0;
    for (k = b->nbr_rowptr[i]; k < b->nbr_rowptr[i + 1]; k++) {
      int j = b->ei[ORCEI_nbr_col][k];
      double q, tr;
      if (j < 0) continue; /* outlet / closed boundary nodes take no part (2681-2684) */
      edge_flux(b, s, i, j, b->ef[ORCE_nbr_len][k], b->ef[ORCE_nbr_vedglen][k], &q, &tr);
      if (q > 0.0 || q < 0.0) {
        contrib[cap_off[i] + cnt[i]++] = q;
        contrib[cap_off[j] + cnt[j]++] = -q;
      }
      if (b->ei[ORCEI_nbr_listpos][k] > best) { best = b->ei[ORCEI_nbr_listpos][k]; besttr = tr; }
    }
    if (best >= 0) s->f[ORCS_Transmissivity][i] = besttr; /* last edge in list order wins (2807, 2821) */
  }
  for (i = 0; i < n; i++) {
    double g = s->f[ORCS_gwchange][i];
    if (cnt[i] > 1) qsort(contrib + cap_off[i], (size_t)cnt[i], sizeof(double), cmp_dbl);
    for (k = 0; k < cnt[i]; k++) g += contrib[cap_off[i] + k];
    s->f[ORCS_gwchange][i] = g;
  }
  free(contrib); free(cnt); free(cap_off);

  for (i = 0; i < n; i++) {
    double alpha, area, gw = s->f[ORCS_gwchange][i];
    double mdelt, dm1, dm2, dm3, nstar, thsurf0, throot0, areaf;
    int st;
    enum { GW_EXFIL, GW_INTSTORM, GW_INITIAL, GW_POSITIVE };

    load_soil(b, i, c);
    load_old(s, i, c);
    c->srf = c->satsrf = 0.0;
    alpha = atan(b->sf[ORC_flow_slope][i]);
    if (alpha > 0.0) c->cosv = fabs(cos(alpha)); else c->cosv = 1.0;
    area = b->sf[ORC_varea][i];
    c->dbr = b->sf[ORC_bedrock][i];

    c->nwt1 = c->nwt0 + dtgw * (gw * 1.0E-6) * c->cosv / (area * c->ths);
    if (c->nwt1 < 0.0) {
      c->satsrf = fabs(c->nwt1 * c->ths) / dtgw;
      c->nwt1 = 0.0;
    }
    st = -1000;
    if (c->nwt1 == c->nwt0) st = GW_INITIAL;
    else if (c->nwt1 - c->nwt0 > 1.0E-3) st = GW_INTSTORM;
    else if (c->nwt0 - c->nwt1 > 1.0E-3) st = GW_POSITIVE;
    if (c->mu0 > c->nwt1 * c->ths) st = GW_EXFIL;
    if (st == -1000) st = GW_INITIAL;

    g_state_counts[10 + st]++;
    switch (st) {
    case GW_EXFIL:
      c->satsrf += (c->mu0 - c->nwt1 * c->ths) / dtgw;
      c->nwt1 = 0;
      c->mu1 = c->mi1 = 0.0;
      c->ri1 = c->ru1 = 0.0;
      c->nf1 = c->nt1 = 0.0;
      break;
    case GW_INITIAL:
      c->nwt1 = c->nwt0; c->mu1 = c->mu0; c->mi1 = c->mi0; c->nf1 = c->nf0;
      c->nt1 = c->nt0; c->ru1 = c->ru0; c->ri1 = c->ri0;
      break;
    case GW_INTSTORM: /* 3105-3180 */
      mdelt = (c->nwt1 - c->nwt0) * c->ths;
      if (c->nwt0 > 0.0) c->nwt1 = newton1(c, (c->ths * c->nwt0 - (c->mi0 - mdelt)), c->nwt1);
      else c->nwt1 = newton1(c, (c->ths * c->nwt0 - (c->mi0 - mdelt)), 0.0);
      c->mi1 = total_moist(c, c->nwt1);
      c->ri1 = 0.0;
      if ((c->nf0 == 0.0) || (c->nf0 == c->nwt0)) {
        c->ru1 = 0.0;
        c->mu1 = c->mi1;
        if (c->nf0 == 0.0) c->nf1 = c->nt1 = 0.0;
        else c->nf1 = c->nt1 = c->nwt1;
      } else {
        dm1 = upper_moist(c, c->nf0, c->nwt1);
        dm2 = c->mu0 - c->mi0;
        mdelt = dm1 + dm2;
        dm3 = c->eps / c->f * (c->ths - c->thr) * (1.0 - exp(-c->f * c->nf0 / c->eps)) + c->thr * c->nf0;
        c->mu1 = c->mi1 + dm2;
        if (mdelt < dm3) {
          c->nt1 = c->nf1 = c->nf0;
          c->ru1 = recharge_rate(c, mdelt, c->nf0);
        } else {
          mdelt = upper_moist(c, c->nf0, c->nwt0);
          mdelt -= dm1;
          if ((c->nt0 > 0.0) && (c->nt0 < c->nf0)) {
            c->nf1 = newton2(c, mdelt, c->nwt1, c->nf0, 0);
            if (c->nf1 <= c->nt0) {
              if ((c->nt0 - c->nf1) >= 10.0) { /* warning only */ }
              else c->nf1 = c->nt0 + 0.001;
            }
            c->nt1 = c->nt0;
            c->ru1 = c->ru0;
          } else if (c->nt0 == 0.0 && c->nf0 > 0.0) {
            c->nf1 = newton2(c, mdelt, c->nwt1, c->nf0, 0);
            c->nt1 = c->nt0;
            c->ru1 = c->ks * c->f * c->nf1 / expm1(c->f * c->nf1);
          }
        }
      }
      break;
    case GW_POSITIVE: /* 3184-3363 */
      if (((c->nwt1 + c->psib) > c->nf0) || (c->nf0 == c->nwt0)) {
        c->mi1 = total_moist(c, c->nwt1);
        dm1 = lower_moist(c, c->nwt1, c->nwt0);
        dm2 = c->mi0 - dm1;
        mdelt = dm1 - (c->mi1 - dm2);
        c->nwt1 = newton1(c, (c->ths * c->nwt1 - (c->mi1 + mdelt)), c->nwt1);
        c->mi1 = total_moist(c, c->nwt1);
        if ((c->nf0 == 0.0) || (c->nf0 == c->nwt0)) {
          c->mu1 = c->mi1;
          c->ru1 = c->ri1 = 0.0;
          if (c->nf0 == c->nwt0) c->nf1 = c->nt1 = c->nwt1;
          else c->nf1 = c->nt1 = 0.0;
        } else if ((c->nf0 > 0.0) && (c->nf0 != c->nwt0)) {
          if ((c->nwt1 + c->psib) > c->nf0) {
            mdelt = upper_moist(c, c->nf0, c->nwt1);
            mdelt += (c->mu0 - c->mi0);
            if (c->nt0 == c->nf0) {
              if (mdelt >= c->nf0 * c->ths) {
                mdelt = -dtgw * (gw * 1.0E-6) * c->cosv / area;
                c->nwt1 = newton1(c, (c->ths * c->nwt0 - (c->mu0 + mdelt)), (c->nt0 - c->psib));
                c->mu1 = c->mi1 = total_moist(c, c->nwt1);
                c->ri1 = c->ru1 = 0.0;
                c->nf1 = c->nt1 = c->nwt1;
              } else {
                c->ru1 = recharge_rate(c, mdelt, c->nf0);
                nstar = (log(c->ks / c->ru1)) / c->f;
                if (nstar <= c->nf0) { c->nt1 = nstar; c->nf1 = c->nf0 + 1.0E-5; }
                else c->nf1 = c->nt1 = c->nf0;
                c->mu1 = c->mi1 + (c->mu0 - c->mi0);
                c->ri1 = 0.0;
              }
            } else if (c->nt0 < c->nf0) {
              if (mdelt >= c->nf0 * c->ths) {
                mdelt -= c->nf0 * c->ths;
                dm3 = c->nwt0;
                c->nwt0 = c->nwt1; /* the reference overwrites its Old scratch here (3261-3262) */
                dm1 = c->ths * (c->nwt1 + c->psib - c->nf1);
                dm2 = z1z2_moist(c, c->nf1, c->nwt1 + c->psib, c->nwt1);
                if ((dm1 - dm2 - mdelt) > 1.0E-6) c->nf1 = newton2(c, mdelt, c->nwt1, c->nf0, 1);
                else c->nf1 = c->nwt1 + c->psib + 1.0E-6;
                c->nt1 = c->nt0;
              } else {
                dm1 = upper_moist(c, c->nf0, c->nwt1);
                dm2 = upper_moist(c, c->nf0, c->nwt0);
                mdelt = dm1 - dm2;
                dm3 = c->nwt0;
                c->nwt0 = c->nwt1;
                dm1 = c->ths * (c->nwt1 + c->psib - c->nf1);
                dm2 = z1z2_moist(c, c->nf1, c->nwt1 + c->psib, c->nwt1);
                if ((dm1 - dm2 - mdelt) > 1.0E-6) c->nf1 = newton2(c, mdelt, c->nwt1, c->nf0, 1);
                else c->nf1 = c->nwt1 + c->psib + 1.0E-6;
                c->nt1 = c->nt0;
              }
              if (c->nf1 >= (c->nwt1 + c->psib)) {
                if (c->nt0 > 0.0) {
                  mdelt = -dtgw * (gw * 1.0E-6) * c->cosv / area;
                  c->nwt0 = dm3;
                  c->nwt1 = newton1(c, (c->ths * c->nwt0 - (c->mu0 + mdelt)), (c->nt0 - c->psib));
                  c->mu1 = c->mi1 = total_moist(c, c->nwt1);
                  c->ri1 = c->ru1 = 0.0;
                  c->nf1 = c->nt1 = c->nwt1;
                } else if (c->nt1 == 0.0) {
                  c->nwt1 = c->nf1 = c->nt1 = 0.0;
                  c->mu1 = c->mi1 = 0.0;
                  c->ri1 = c->ru1 = 0.0;
                }
              } else {
                c->mu1 = c->mi1 + (c->mu0 - c->mi0);
                c->ri1 = 0.0;
                if (c->nt1 == 0.0) c->ru1 = c->ks * c->f * c->nf1 / expm1(c->f * c->nf1);
                else c->ru1 = c->ru0;
              }
            }
          } else {
            mdelt = -dtgw * (gw * 1.0E-6) * c->cosv / area;
            c->nwt1 = newton1(c, (c->ths * c->nwt0 - (c->mu0 + mdelt)), c->nwt0);
            c->mu1 = c->mi1 = total_moist(c, c->nwt1);
            c->ri1 = c->ru1 = 0.0;
            c->nf1 = c->nt1 = c->nwt1;
          }
        }
      } else if (((c->nwt1 + c->psib) <= c->nf0) && (c->nf0 != c->nwt0)) {
        c->mi1 = total_moist(c, c->nwt1);
        dm1 = lower_moist(c, c->nwt1, c->nwt0);
        dm2 = c->mi0 - dm1;
        mdelt = dm1 - (c->mi1 - dm2);
        if (c->nt0 == c->nf0)
          c->nwt1 = newton1(c, (c->ths * c->nwt0 - (c->mu0 + c->ths * (c->nwt0 - c->nwt1))), c->nwt1);
        else if (c->nt0 < c->nf0)
          c->nwt1 = newton1(c, (c->ths * c->nwt0 - (c->mu0 + c->ths * (c->nwt0 - c->nwt1))), (c->nt0 - c->psib));
        c->mu1 = c->mi1 = total_moist(c, c->nwt1);
        c->ru1 = c->ri1 = 0.0;
        c->nt1 = c->nwt1;
        c->nf1 = c->nwt1;
      }
      break;
    }

    /* 3438-3533 */
    c->esrf = s->glob[ORCG_psrf_last] + c->satsrf; /* stale psrf member of the last unsat cell */
    c->esrf = c->esrf / c->cosv;
    c->satsrf /= c->cosv;
    c->srf = c->satsrf;

    thsurf0 = s->f[ORCS_SoilMoistureSC][i];
    throot0 = s->f[ORCS_RootMoistureSC][i];
    store_new(s, i, c);
    s->f[ORCS_srf_hr][i] += c->srf * dtgw;
    s->f[ORCS_srf][i] = s->f[ORCS_srf][i] + c->srf * dtgw;
    s->f[ORCS_cumsrf][i] += s->f[ORCS_srf][i]; /* double counting kept (3460-3461) */
    s->f[ORCS_satsrf][i] = c->satsrf * dtgw;
    s->f[ORCS_esrf][i] = c->esrf * dtgw;
    if (c->satsrf > 0.0) s->f[ORCS_satsrfOccur][i] = s->f[ORCS_satsrfOccur][i] + floor(c->satsrf * 1.0E+3) + 1.0E-6;
    if (surf_soil_moist(c, 1.0) / c->ths > 0.999) s->f[ORCS_satOccur][i] = s->f[ORCS_satOccur][i] + 1;
    s->f[ORCS_RechDisch][i] =
        s->f[ORCS_RechDisch][i] + ((c->nwt0 - c->nwt1) + c->satsrf * dtgw * c->cosv / c->ths) * 1.0E-3;
    moisture_diag(b, s, i, c, current_time, &thsurf);
    s->f[ORCS_Recharge][i] = (c->nwt1 - c->nwt0) * c->ths / (c->cosv * dtgw);
    s->glob[ORCG_Stok] += c->srf * dtgw * area / 1000.0;
    s->glob[ORCG_TotMoist] += (c->mu1 - c->mu0) * area / (c->cosv * 1000.0);
    s->glob[ORCG_TotGWchange] += (c->nwt1 - c->nwt0) * c->ths * area / (c->cosv * 1000.0);
    if (fabs(s->f[ORCS_SoilMoistureSC][i] - thsurf0) > 1.0E-6) {
      s->glob[ORCG_fGW100] += fabs(dtgw * (gw * 1.0E-6) / area) * 1.0 * area;
      gwdm += (s->f[ORCS_SoilMoistureSC][i] - thsurf0) * c->ths * 100.0 * area;
      areagw += area;
      gwcnt++;
    }
    areaf = area / b->BasArea;
    s->glob[ORCG_dMRt] += (s->f[ORCS_RootMoistureSC][i] - throot0) * c->ths * 1000.0 * areaf;
    mth100 += (s->f[ORCS_SoilMoistureSC][i] + thsurf0) / 2.0 * areaf;
    mthrt += (s->f[ORCS_RootMoistureSC][i] + throot0) / 2.0 * areaf;
  }
  (void)gwdm; (void)mth100; (void)mthrt; (void)areagw; (void)gwcnt;
  /* 3536-3563: the eight relative-forcing sums are normalised, (not) printed, and zeroed */
  s->glob[ORCG_fSoi100] = s->glob[ORCG_fTop100] = s->glob[ORCG_fClm100] = s->glob[ORCG_fGW100] = 0.0;
  s->glob[ORCG_dM100] = s->glob[ORCG_dMRt] = s->glob[ORCG_mTh100] = s->glob[ORCG_mThRt] = 0.0;
}

/* ------------------------------------------------------------------------------------
 * Hillslope routing + basin accumulations
 * (reference: tKinemat::RunHydrologicRouting tKinemat.cpp:920-1089, setTravelVelocityKin
 *  1100-1125, tFlowNet::SurfaceFlow tFlowNet.cpp:735-867, tFlowResults.cpp:857-1161)
 * ---------------------------------------------------------------------------------- */
static void stack_add(orc_stack *st, int step, double v) {
  int k, pos;
  if (st->n == st->cap) {
    st->cap = st->cap ? 2 * st->cap : 8;
    st->tind = (int32_t *)realloc(st->tind, sizeof(int32_t) * (size_t)st->cap);
    st->q = (double *)realloc(st->q, sizeof(double) * (size_t)st->cap);
  }
  if (st->n == 0 || st->tind[st->n - 1] < step) { st->tind[st->n] = step; st->q[st->n] = v; st->n++; return; }
  pos = 0;
  while (pos < st->n && st->tind[pos] < step) pos++;
  if (st->tind[pos] == step) { st->q[pos] += v; return; }
  for (k = st->n; k > pos; k--) { st->tind[k] = st->tind[k - 1]; st->q[k] = st->q[k - 1]; }
  st->tind[pos] = step; st->q[pos] = v; st->n++;
}

static int res_step(const orc_basin *b, double t, double now) { return (int)floor((t + now) / b->output_interval); }

static void store_volume(const orc_basin *b, double *arr, int limit, double now, double time, double value) {
  /* value is already divided by getOutputIntervalSec by the caller-specific add_m_volume;
     here arr receives value/(output_interval*3600) per the reference's add_m_volume */
  double dcalc = b->tstep, dres = b->output_interval, dint;
  double sec = b->output_interval * 3600.0;
  int init = res_step(b, time + .01 * dcalc, now), end = res_step(b, time + .99 * dcalc, now), ii;
  if (init < 0 || end >= limit) return;
  if (init == end) arr[init] += value / sec;
  else if (end - init == 1) {
    dint = (init + 1) * dres - (now + time);
    arr[init] += (value * dint / dcalc) / sec;
    arr[end] += (value * (dcalc - dint) / dcalc) / sec;
  } else {
    dint = (init + 1) * dres - (now + time);
    arr[init] += (value * dint / dcalc) / sec;
    for (ii = init + 1; ii < end; ii++) arr[ii] += (value * dres / dcalc) / sec;
    dint = (now + time + dcalc) - end * dres;
    arr[end] += (value * dint / dcalc) / sec;
  }
}

void orc_route(const orc_basin *b, orc_state *s, orc_results *r, double now) {
  int i, n = b->n;
  double hillvel = b->velcoef / b->velratio; /* value left by tFlowNet::setTravelVelocity (531-532) */
  double dtreff = b->dtReff;
  int cur_step = (int)ceil(now / dtreff);
  int at_end = ((int)ceil(now / dtreff) == (int)floor(now / dtreff));

  for (i = 0; i < n; i++) {
    double srf = s->f[ORCS_srf][i];
    int bnd = b->si[ORCI_boundary][i];
    if (bnd == 0) { /* setTravelVelocityKin for every hillslope cell */
      int sn = b->si[ORCI_stream_node][i];
      double q, carea, flowout;
      if (sn < 0) { /* drains straight into the outlet node (cn->getStreamNode() == OutletNode) */
        q = s->glob[ORCG_Qstrm_outlet];
        carea = b->outlet_contr_area;
      } else {
        int dn = b->si[ORCI_flow_dest][sn];
        carea = b->sf[ORC_contr_area][sn];
        q = (s->f[ORCS_Qstrm][sn] + (dn >= 0 ? s->f[ORCS_Qstrm][dn] : s->glob[ORCG_Qstrm_outlet])) / 2.;
      }
      if (!q) flowout = b->baseflow * carea / b->outlet_contr_area; else flowout = q;
      if (b->flowexp > 0.0) hillvel = b->kincoef * pow(flowout / carea, b->flowexp);
      else hillvel = b->velcoef / b->velratio;
    }
    if (srf > 0.0) {
      double vr = srf * b->sf[ORC_varea][i] / 1000.;
      int nss;
      orc_stack *st;
      if (srf > 999999) vr = 0.;
      if (bnd == 0) {
        double tt;
        s->f[ORCS_TTime][i] = b->sf[ORC_hillpath][i] / hillvel / 3600.;
        tt = s->f[ORCS_TTime][i];
        nss = (int)ceil((now + tt) / dtreff);
        if (cur_step == nss) {
          if (at_end) vr /= (b->tstep * 3600.);
          else vr /= ((nss * dtreff - now + b->tstep) * 3600.);
        } else
          vr /= (dtreff * 3600.);
        st = &r->stacks[b->si[ORCI_stream_node][i] >= 0 ? b->si[ORCI_stream_node][i] : n];
      } else {
        nss = cur_step;
        if (at_end) vr /= (b->tstep * 3600.);
        else vr /= ((nss * dtreff - now + b->tstep) * 3600.);
        st = &r->stacks[i];
      }
      stack_add(st, nss, vr);
    }
    s->f[ORCS_FlowVelocity][i] = hillvel;
  }
  s->glob[ORCG_hillvel_last] = hillvel;

  /* tFlowNet::SurfaceFlow */
  {
    int bin0 = res_step(b, 0.0 - .01 * b->tstep, now), bin1 = res_step(b, 0.0 - .99 * b->tstep, now);
    double w = b->tstep / b->output_interval;
    double hv = b->velcoef / b->velratio, sv = b->velcoef;
    for (i = 0; i < n; i++) {
      double area = b->sf[ORC_varea][i], areaf = area / b->BasArea_flow, srf = s->f[ORCS_srf][i];
      double rain = s->f[ORCS_Rain][i];
      if (srf > 0.0) {
        double tt;
        if (b->flowexp > 0.0) s->f[ORCS_TTime][i] = (b->sf[ORC_hillpath][i] / hv + b->sf[ORC_streampath][i] / sv) / 3600.0;
        tt = s->f[ORCS_TTime][i];
        store_volume(b, r->a[ORCR_mhydro], r->limit, now, tt, srf * area / 1000.0);
        if (s->f[ORCS_hsrf][i]) store_volume(b, r->a[ORCR_HsrfRout], r->limit, now, tt, s->f[ORCS_hsrf][i] * area / 1000.0);
        if (s->f[ORCS_sbsrf][i]) store_volume(b, r->a[ORCR_SbsrfRout], r->limit, now, tt, s->f[ORCS_sbsrf][i] * area / 1000.0);
        if (s->f[ORCS_psrf][i]) store_volume(b, r->a[ORCR_PsrfRout], r->limit, now, tt, s->f[ORCS_psrf][i] * area / 1000.0);
        if (s->f[ORCS_satsrf][i]) store_volume(b, r->a[ORCR_SatsrfRout], r->limit, now, tt, s->f[ORCS_satsrf][i] * area / 1000.0);
      }
      if (bin0 == bin1 && bin0 >= 0 && bin0 < r->limit) {
        double cs = cos(atan(b->sf[ORC_flow_slope][i]));
        double nwtv;
        if (cs < 1E-9) cs = 1.E-9;
        nwtv = s->f[ORCS_NwtNew][i] / cs;
        r->a[ORCR_crr][bin0] += areaf * rain * b->tstep / b->output_interval;
        if (rain > r->a[ORCR_rmax][bin0]) r->a[ORCR_rmax][bin0] = rain;
        if (rain <= r->a[ORCR_rmin][bin0]) r->a[ORCR_rmin][bin0] = rain;
        if (rain > 0.0) r->a[ORCR_frac][bin0] += areaf * w;
        r->a[ORCR_msm][bin0] += areaf * s->f[ORCS_SoilMoistureSC][i] * w;
        r->a[ORCR_msmRt][bin0] += areaf * s->f[ORCS_RootMoistureSC][i] * w;
        r->a[ORCR_msmU][bin0] += areaf * s->f[ORCS_SoilMoistureUNSC][i] * w;
        if (s->f[ORCS_SoilMoistureSC][i] >= 1) r->a[ORCR_sat][bin0] += areaf * w;
        r->a[ORCR_mgw][bin0] += areaf * nwtv * w;
        r->a[ORCR_qunsat][bin0] += areaf * (s->f[ORCS_Qpout][i] - s->f[ORCS_Qpin][i]) * 1.E-6 / area * w;
      }
    }
  }
}

/* What the channel router sees: tKinemat::RetrieveQeff (tKinemat.cpp:1504-1551) applied once to
   every stream cell (and the outlet slot n); out has n+1 entries */
void orc_retrieve_qeff(const orc_basin *b, orc_results *r, double now, double *out) {
  int i, k, n = b->n;
  int cur = (int)ceil(now / b->dtReff);
  int at_end = ((int)ceil(now / b->dtReff) == (int)floor(now / b->dtReff));
  for (i = 0; i <= n; i++) {
    orc_stack *st = &r->stacks[i];
    out[i] = 0.0;
    if (i < n && b->si[ORCI_boundary][i] != 3) continue;
    if (st->n > 0 && st->tind[0] == cur) {
      out[i] = st->q[0];
      if (at_end) {
        for (k = 1; k < st->n; k++) { st->tind[k - 1] = st->tind[k]; st->q[k - 1] = st->q[k]; }
        st->n--;
      }
    }
  }
}

void orc_reset(const orc_basin *b, orc_state *s) {
  int i;
  for (i = 0; i < b->n; i++) {
    s->f[ORCS_NwtOld][i] = s->f[ORCS_NwtNew][i];
    s->f[ORCS_MuOld][i] = s->f[ORCS_MuNew][i];
    s->f[ORCS_MiOld][i] = s->f[ORCS_MiNew][i];
    s->f[ORCS_NfOld][i] = s->f[ORCS_NfNew][i];
    s->f[ORCS_NtOld][i] = s->f[ORCS_NtNew][i];
    s->f[ORCS_RuOld][i] = s->f[ORCS_RuNew][i];
    s->f[ORCS_RiOld][i] = s->f[ORCS_RiNew][i];
    s->f[ORCS_Qpin][i] = 0.0;
    s->f[ORCS_srf][i] = 0.0;
    s->f[ORCS_sbsrf][i] = 0.0;
    s->f[ORCS_hsrf][i] = 0.0;
    s->f[ORCS_psrf][i] = 0.0;
    s->f[ORCS_satsrf][i] = 0.0;
    s->f[ORCS_gwchange][i] = 0.0;
  }
}

void orc_balance(const orc_basin *b, orc_state *s, int gw_time, int met_time, double metstep_h,
                 const double *intercept_loss, const double *evap_wet_canopy, const double *evapotrans,
                 const double *can_storage, double *canopy_stor_vol, double *basin6) {
  const double unsStep = b->tstep * 60.0, satStep = b->gwstep * 60.0;   /* the keywords are in minutes */
  int i;
  {   /* UnSaturatedBalance, tWaterBalance.cpp:165-200 */
    const double dt = unsStep / 60.0;
    for (i = 0; i < b->n; i++) {
      double QUin = s->f[ORCS_UnSatFlowIn][i], QUout = s->f[ORCS_UnSatFlowOut][i], Rech = s->f[ORCS_Recharge][i];
      double Run = s->f[ORCS_srf][i] / unsStep, A = b->sf[ORC_varea][i], Inf = s->f[ORCS_NetPrecipitation][i];
      double Melt = s->f[ORCS_LiqRouted][i] * 10.0, ET, DelU;
      if (s->f[ORCS_LiqWE][i] + s->f[ORCS_IceWE][i] > 1e-4) Inf = Melt;
      else Inf = Inf + Melt;
      ET = s->f[ORCS_EvapSoil][i] + s->f[ORCS_EvapDryCanopy][i];
      DelU = Inf + QUin - QUout - Rech - ET - Run;
      s->f[ORCS_UnSaturatedStorage][i] = s->f[ORCS_UnSaturatedStorage][i] + DelU * dt * A / 1000.0;
    }
  }
  if (gw_time) {   /* SaturatedBalance, tWaterBalance.cpp:234-262; QgwIn/QgwOut are never written by the model */
    const double dt = satStep / 60.0;
    for (i = 0; i < b->n; i++) {
      double QSTin = 0.0 * 1.0E-9, QSTout = 0.0 * 1.0E-9, Rech = s->f[ORCS_Recharge][i];
      double Exf = s->f[ORCS_srf][i] / satStep, A = b->sf[ORC_varea][i];
      double DelG = QSTin - QSTout + (Rech - Exf) * A / 1000.0;
      s->f[ORCS_SaturatedStorage][i] = s->f[ORCS_SaturatedStorage][i] + DelG * dt;
    }
  }
  if (met_time && canopy_stor_vol) {   /* CanopyBalance, tWaterBalance.cpp:116-140 */
    for (i = 0; i < b->n; i++) {
      double DelI = intercept_loss[i] - evap_wet_canopy[i];
      canopy_stor_vol[i] = canopy_stor_vol[i] + DelI * metstep_h * b->sf[ORC_varea][i] / 1000.0;
    }
  }
  {   /* BasinStorage, tWaterBalance.cpp:283-337: three running sums; the other three keep the LAST cell's value */
    double rain = 0.0, evap = 0.0, run = 0.0, can = 0.0, uns = 0.0, sat = 0.0;
    for (i = 0; i < b->n; i++) {
      double a = b->sf[ORC_varea][i] / 1000.0;
      can = (can_storage ? can_storage[i] : 0.0) * a;
      rain += s->f[ORCS_Rain][i] * a * unsStep / 60.0;
      evap += (evapotrans ? evapotrans[i] : 0.0) * a * unsStep / 60.0;
      run += s->f[ORCS_srf][i] * a;
      uns = s->f[ORCS_MuNew][i] * a;
      sat = b->sf[ORC_ThetaS][i] * (b->sf[ORC_bedrock][i] - s->f[ORCS_NwtNew][i]) * a;
    }
    basin6[0] += rain; basin6[1] += evap; basin6[2] = can; basin6[3] = uns; basin6[4] = sat; basin6[5] += run;
  }
}

/* ------------------------------------------------------------------------------------ KATs */
static void kat_cell(cw *c, const double *soil) {
  memset(c, 0, sizeof *c);
  c->ks = soil[0]; c->ths = soil[1]; c->thr = soil[2]; c->lam = soil[3]; c->psib = soil[4];
  c->f = soil[5]; c->ar = soil[6]; c->uar = soil[7]; c->eps = 3 + 2 / c->lam; c->dbr = soil[8];
}
double orc_kat_lambertw(double z) { return lambertw(z); }
double orc_kat_total_moist(const double *soil, double nwt) { cw c; kat_cell(&c, soil); return total_moist(&c, nwt); }
double orc_kat_upper_moist(const double *soil, double nf, double nwt) { cw c; kat_cell(&c, soil); return upper_moist(&c, nf, nwt); }
double orc_kat_lower_moist(const double *soil, double nf, double nwt) { cw c; kat_cell(&c, soil); return lower_moist(&c, nf, nwt); }
double orc_kat_newton1(const double *soil, double dm, double nwt) { cw c; kat_cell(&c, soil); return newton1(&c, dm, nwt); }
double orc_kat_newton2(const double *soil, double dm, double nwt, double nf, int flag) { cw c; kat_cell(&c, soil); return newton2(&c, dm, nwt, nf, flag); }
double orc_kat_recharge(const double *soil, double dm, double nf) { cw c; kat_cell(&c, soil); return recharge_rate(&c, dm, nf); }
double orc_kat_unsat_lat(const double *soil, double nf, double ru, double vel) { cw c; kat_cell(&c, soil); return unsat_lat(&c, nf, ru, vel); }
double orc_kat_sat_lat(const double *soil, double nt, double nf, double ru, double vel) { cw c; kat_cell(&c, soil); return sat_lat(&c, nt, nf, ru, vel); }
double orc_kat_z1z2(const double *soil, double z1, double z2, double nwt) { cw c; kat_cell(&c, soil); return z1z2_moist(&c, z1, z2, nwt); }

int orc_sizes(int which) {
  switch (which) {
  case 0: return ORC_N_STATIC_F64;
  case 1: return ORC_N_STATIC_I32;
  case 2: return ORC_N_EDGE_F64;
  case 3: return ORC_N_EDGE_I32;
  case 4: return ORC_N_STATE_F64;
  case 5: return ORC_N_GLOBALS;
  case 6: return ORC_N_RESULTS;
  default: return (int)sizeof(orc_basin);
  }
}
