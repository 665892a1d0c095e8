// physics.cuh — per-column (one Voronoi cell) water-balance physics for the sm_100a kernels.
//
// One `Column` lives in the registers of one CUDA thread. It carries the soil parameters,
// the Old/New prognostic state and the Green-Ampt scratch terms that the reference keeps in
// tHydroModel data members, so every helper states its inputs explicitly.
//
// Reference behaviour restated here (file:line in /root/reference/src/tHydro/tHydroModel.cpp):
//   profile integrals 2240-2315, point moisture 2324-2342, suction 2353-2364,
//   recharge / lateral flows / transmissivity 2371-2426, surface moisture 3594-3642,
//   Z1-Z2 moisture 3654-3686, Lambert W 3831-3915 (+ Mathutil/mathutil.cpp:1549-1670),
//   root finders 3937-4192, unsaturated state machine 779-2075, saturated 3036-3503.
// Arithmetic is kept in the reference's operation order (FP64, no re-association) so the
// only differences to the CPU model are the device libm's last-ulp differences.
//
// The functions are __host__ __device__ only so that tests/hostcheck can compile this very
// header with g++ and compare it bit-for-bit with the oracle on a box without a GPU. The
// product library never executes them on the host.
#pragma once
#include <math.h>
#include <stdint.h>
#include "const_math.cuh"

#if defined(__CUDACC__)
#define TRIBS_HD __host__ __device__ __forceinline__
#define TRIBS_HD_NOINLINE __host__ __device__ __noinline__
#else
#define TRIBS_HD inline
#define TRIBS_HD_NOINLINE inline
#endif

namespace tribs {

// One out-of-line copy of each libm routine per kernel: the state machines call them from
// ~150 sites, and inlining every call makes the sweep kernel > 600 KB of SASS (instruction-cache
// misses on a latency-bound chain). A call costs ~20 cycles against several hundred of math.
#if defined(__CUDA_ARCH__)
#define TRIBS_MATH __device__ __noinline__
#else
#define TRIBS_MATH inline
#endif
#if defined(TRIBS_CONST_MATH) && (defined(__CUDA_ARCH__) || defined(TRIBS_CONST_MATH_HOST))
// (TRIBS_CONST_MATH_HOST: tests/exp_perturb.py runs the host build of these headers with the same routines)
// exp / log with their coefficients in the constant bank (const_math.cuh): fewer instructions in the
// routines that carry most of the sweep's instruction stream; x^y as exp(y ln x) as below.
TRIBS_MATH double m_exp(double x) { return cm::exp_c(x); }
TRIBS_MATH double m_log(double x) { return cm::log_c(x); }
TRIBS_MATH double m_pow(double x, double y) { return cm::pow_c(x, y); }
#else
TRIBS_MATH double m_exp(double x) { return exp(x); }
TRIBS_MATH double m_log(double x) { return log(x); }
#if defined(__CUDA_ARCH__) && defined(TRIBS_FAST_POW)
// x^y through exp(y*log x): ~3x fewer FP64 instructions than the double-double pow of the CUDA
// library; relative error ~|y ln x| * 1e-16, far inside the 1e-6 parity tolerance. Non-positive
// bases keep the library's semantics (0^y, NaN for negative bases with fractional exponents).
TRIBS_MATH double m_pow(double x, double y) { return (x > 0.0) ? exp(y * log(x)) : pow(x, y); }
#else
TRIBS_MATH double m_pow(double x, double y) { return pow(x, y); }
#endif
#endif
TRIBS_MATH double m_expm1(double x) { return expm1(x); }

constexpr double kRefPi = 3.1415926;       // src/Headers/Definitions.h:54 (8 digits, on purpose)
constexpr double kLambEps = 2.2204E-16;    // src/tHydro/tHydroModel.h:32
constexpr double kRadar = 0.1;             // mm/h, tHydroModel.cpp:751

enum UnsatState : int {
  kStormEvol = 0, kWtStaysAtSurf, kWtDropsFromSurf, kWtGetsToSurf, kStormUnsatEvol, kPerchedEvol,
  kPerchedSurfSat, kStormToInter, kExactInitial, kIntStormBelow
};
enum SatState : int { kGwExfiltrate = 0, kGwIntStormLike, kGwInitial, kGwPositiveBal };

struct Column {
  // soil
  double ks, ths, thr, lam, psib, f, ar, uar, eps, dbr;
  // prognostic state: 0 = value at step start, 1 = value being computed
  double nwt0, nwt1, mu0, mu1, mi0, mi1, nf0, nf1, nt0, nt1, ru0, ru1, ri0, ri1;
  double qpin, qpout, isv;
  double r, r1, cosv, sinv;
  double hsrf, psrf, sbsrf, satsrf;
  // Green-Ampt scratch
  double g, thri, thre;
  // one-entry memo of (-psib/nwt)^lam: the profile integrals of one update ask for it again
  // and again with the same water table (identical value, so results do not change)
  double memo_nwt, memo_val;
  int memo_ok;

  TRIBS_HD double fringe_pow(double nwt) {
    if (memo_ok && nwt == memo_nwt) return memo_val;
    memo_val = m_pow((-psib / nwt), lam);
    memo_nwt = nwt;
    memo_ok = 1;
    return memo_val;
  }

  TRIBS_HD bool lam_one() const { return lam > (1.0 - 1.0E-6) && lam < (1.0 + 1.0E-6); }

  // -------------------------------------------------------------- profile integrals
  TRIBS_HD double total_moist(double nwt) {
    double dm = 0.0;
    if (nwt >= fabs(psib)) {
      if (lam_one())
        dm = thr * (nwt + psib) - fabs(psib) * (ths - thr) * m_log((-psib) / nwt) - ths * psib;
      else
        dm = thr * (nwt + psib) - ths * psib - (ths - thr) / (lam - 1.0) * (psib + nwt * fringe_pow(nwt));
    } else if (nwt == 0.0) {
      dm = 0.0;
    } else {
      nwt1 = 0.0;  // table inside the capillary fringe -> at the surface (2258)
    }
    return dm;
  }

  TRIBS_HD double upper_moist(double nf, double nwt) {
    double dm;
    if (nf <= (nwt + psib)) {
      if (lam_one())
        dm = thr * nf + fabs(psib) * (ths - thr) * m_log(nwt / (nwt - nf));
      else
        dm = thr * nf - (ths - thr) / (lam - 1.0) *
                            ((nf - nwt) * m_pow((psib / (nf - nwt)), lam) + nwt * fringe_pow(nwt));
    } else {
      if (lam_one())
        dm = total_moist(nwt) + ths * psib + ths * (nf - (nwt + psib));
      else
        dm = thr * (nwt + psib) - (ths - thr) / (lam - 1.0) * (psib + nwt * fringe_pow(nwt)) +
             ths * (nf - (nwt + psib));
    }
    return dm;
  }

  TRIBS_HD double lower_moist(double nf, double nwt) const {
    double dm;
    if (nf <= (nwt + psib)) {
      if (lam_one())
        dm = thr * (nwt + psib - nf) - fabs(psib) * (ths - thr) * m_log((-psib) / (nwt - nf)) - ths * psib;
      else
        dm = thr * (nwt + psib - nf) - ths * psib -
             (ths - thr) / (lam - 1.0) * (psib + (nwt - nf) * m_pow((psib / (nf - nwt)), lam));
    } else
      dm = ths * (nwt - nf);
    return dm;
  }

  // moisture held by an unsaturated wedge that reaches depth d with the top at the surface
  TRIBS_HD double wedge_cap(double d) const {
    return eps / f * (ths - thr) * (1.0 - m_exp(-f * d / eps)) + thr * d;
  }

  TRIBS_HD double init_moist_z(double nf) const {
    if (nf >= (nwt0 + psib)) return ths;
    return (thr + (ths - thr) * m_pow((psib / (nf - nwt0)), lam));
  }
  TRIBS_HD double edge_moist_z(double nf) const {
    return (thr + m_exp(f * nf / eps) * (ths - thr) * m_pow((ru1 / ks), (1.0 / eps)));
  }

  TRIBS_HD void suction(double nf) {
    const double decay = m_exp(-f * nf / 2.);       // evaluated once; the reference repeats it
    const double expo = (3. + 1. / (lam * decay));
    if (thre == ths) {
      double sein = m_pow(((thri - thr) / (ths - thr)), expo);
      g = -psib * (1. - sein) / (3 * lam * decay + 1.);
    } else {
      double sein = m_pow(((thri - thr) / (ths - thr)), expo);
      double se0 = m_pow(((thre - thr) / (ths - thr)), expo);
      g = -psib * (se0 - sein) / (3.0 * lam * decay + 1.);
    }
  }

  TRIBS_HD double recharge_rate(double dm, double nf) const {
    double bb = (dm - thr * nf) / (ths - thr);
    bb = bb * f / (eps * m_expm1(f * nf / eps));
    return (ks * m_pow(bb, eps));
  }
  TRIBS_HD double unsat_lateral(double nf, double ru, double vel) const { return (nf * ru * (uar - 1.0) * vel); }
  TRIBS_HD double sat_lateral(double nt, double nf, double ru, double vel) const {
    double bb;
    if (nt == 0.0)
      bb = (ks * uar / f * (1. - m_exp(-f * nf)) - ks * f * nf * nf / m_expm1(f * nf)) * vel;
    else {
      bb = nt * ru * (uar - 1.0) + ks * uar / f * (m_exp(-f * nt) - m_exp(-f * nf));
      bb = (bb - ks * f * (nf - nt) * (nf - nt) / (m_exp(f * nf) - m_exp(f * nt))) * vel;
    }
    return bb;
  }
  // lateral outflow of the wedge through the flow edge, common tail of four states
  TRIBS_HD void wedge_outflow(double vel_mm) {
    qpout = 0.0;
    if ((nt1 == nf1) && (nf1 != nwt1) && (nf1 > 10.0)) qpout = unsat_lateral(nf1, ru1, vel_mm);
    if ((nf1 - nt1) > 1.0) qpout = sat_lateral(nt1, nf1, ru1, vel_mm);
    qpout *= sinv;
  }
  // surface-saturated conductance integral Ks*F*Nf/(e^{F Nf}-1)
  TRIBS_HD double ksat_mean(double nf) const { return ks * f * nf / m_expm1(f * nf); }

  // -------------------------------------------------------------- surface soil moisture
  TRIBS_HD double surf_soil_moist(double z) {
    double dm = 0, nstar = 0;
    if (nwt1 == 0.0)
      dm = z * ths;
    else if ((nwt1 > 0.0) && (nf1 == 0.0 || nf1 == nwt1))
      dm = upper_moist(z, nwt1);
    else if (nf1 > 0.0 && nt1 == nf1 && nf1 < nwt1) {
      if (fabs(ru1) > 1.0E-6) {
        nstar = m_log(ks / ru1) / f;
        if (nf1 >= z)
          dm = eps / f * (ths - thr) * (m_exp(f * (z - nstar) / eps) - m_exp(-f * nstar / eps)) + thr * z;
        else {
          dm = upper_moist(z, nwt1);
          dm += (mu1 - mi1);
        }
      } else
        dm = upper_moist(z, nwt1);
    } else if (nf1 > 0.0 && nt1 < nf1 && nt1 > 0.0) {
      nstar = nt1;
      if (nt1 >= z)
        dm = eps / f * (ths - thr) * (m_exp(f * (z - nstar) / eps) - m_exp(-f * nstar / eps)) + thr * z;
      else {
        if (nf1 >= z) {
          dm = eps / f * (ths - thr) * (1.0 - m_exp(-f * nt1 / eps)) + thr * nt1;
          dm += ((z - nt1) * ths);
        } else if (nf1 < z) {
          dm = upper_moist(z, nwt1);
          dm += (mu1 - mi1);
        }
      }
    } else if (nf1 > 0.0 && nt1 == 0.0) {
      if (nf1 >= z)
        dm = z * ths;
      else if (nf1 < z) {
        dm = upper_moist(z, nwt1);
        dm += (mu1 - mi1);
      }
    }
    return (dm / z);
  }

  TRIBS_HD double z1z2_base(double z1, double z2, double nwt) const {
    if (lam_one()) return thr * (z2 - z1) - fabs(psib) * (ths - thr) * m_log((nwt - z2) / (nwt - z1));
    return thr * (z2 - z1) - (ths - thr) / (lam - 1.0) *
                                 ((z2 - nwt) * m_pow((-psib / (nwt - z2)), lam) - (z1 - nwt) * m_pow((-psib / (nwt - z1)), lam));
  }
  // the reference recurses once at most (3676-3681); unrolled here
  TRIBS_HD double z1z2_moist(double z1, double z2, double nwt) const {
    double lim = nwt + psib + 1.0E-3;
    if (z2 <= lim) return z1z2_base(z1, z2, nwt);
    if (z2 > lim && z1 < lim) {
      double zb = nwt + psib, dm;
      if (zb <= lim) dm = z1z2_base(z1, zb, nwt);
      else dm = (zb - z1) * ths;  // unreachable (zb < lim always), kept for totality
      dm += (z2 - (nwt + psib)) * ths;
      return dm;
    }
    return (z2 - z1) * ths;
  }

  // -------------------------------------------------------------- root finders (bodies below the struct)
  // water-table depth whose initial profile has moisture deficit dm (3937-4020); see solve_wt_soil
  TRIBS_HD double solve_wt(double dm, double nwt) const;
  // wetting-front depth after dm is taken from (flag 0) / added to (flag 1) the wedge (4037-4082)
  TRIBS_HD double solve_front(double dm, double nwt, double nf, int flag) const;

  // re-initialise the profile for water table nwt1: Mi = Mu = total moisture, no wedge rates
  TRIBS_HD void settle_profile() {
    mi1 = total_moist(nwt1);
    mu1 = mi1;
  }
};

// ------------------------------------------------------------------ root finders, out of line
// The two solvers are called from ~20 sites of the state machines. They take the five soil numbers
// they need BY VALUE and return the root, so that no address of the caller's Column escapes: the
// column then lives in registers instead of a stack frame that every helper call would force the
// compiler to keep current (the round-1 kernel carried 1 KB of stack per thread for this).
struct SoilK { double ths, thr, psib, lam, dbr; };

TRIBS_HD void soil_poly_wt(const SoilK &k, double x, double &fv, double &dv, double c1, double c2, double c3) {
  fv = c1 * m_pow(x, k.lam) + c3 * m_pow(x, (k.lam - 1.0)) + c2;
  dv = c1 * k.lam * m_pow(x, (k.lam - 1.0)) + c3 * (k.lam - 1.0) * m_pow(x, (k.lam - 2.0));
}

// safeguarded Newton / bisection in a bracket (4094-4167), 30 iterations at most
TRIBS_HD double soil_bracketed_root(const SoilK &k, double c1, double c2, double c3, double x1, double x2, double xacc, double xguess) {
  double df = 0.0, dx, dxold, fx, fh = 0.0, fl = 0.0, temp, xh, xl, rts;
  soil_poly_wt(k, x1, fl, df, c1, c2, c3);
  soil_poly_wt(k, x2, fh, df, c1, c2, c3);
  if ((fl > 0.0 && fh > 0.0) || (fl < 0.0 && fh < 0.0)) return xguess;
  if (fl == 0.0) return x1;
  if (fh == 0.0) return x2;
  if (fl < 0.0) { xl = x1; xh = x2; } else { xh = x1; xl = x2; }
  rts = xguess;
  dxold = fabs(x2 - x1);
  dx = dxold;
  soil_poly_wt(k, rts, fx, df, c1, c2, c3);
  for (int j = 1; j <= 30; j++) {
    if ((((rts - xh) * df - fx) * ((rts - xl) * df - fx) > 0.0) || (fabs(2.0 * fx) > fabs(dxold * df))) {
      dxold = dx;
      dx = 0.5 * (xh - xl);
      rts = xl + dx;
      if (xl == rts) return rts;
    } else {
      dxold = dx;
      dx = fx / df;
      temp = rts;
      rts -= dx;
      if (temp == rts) return rts;
    }
    if (fabs(dx) < xacc) return rts;
    soil_poly_wt(k, rts, fx, df, c1, c2, c3);
    if (fx < 0.0) xl = rts; else xh = rts;
  }
  return xguess;
}

// total moisture above a water table at depth nwt (tHydroModel::get_Total_Moist 2240-2262) without the
// member side effect (the reference zeroes NwtNew when the table lies inside the capillary fringe; every
// caller of the solver assigns NwtNew from its result right afterwards, so the side effect never shows)
TRIBS_HD double soil_total_moist(const SoilK &k, double nwt) {
  double dm = 0.0;
  if (nwt >= fabs(k.psib)) {
    if (k.lam > (1.0 - 1.0E-6) && k.lam < (1.0 + 1.0E-6))
      dm = k.thr * (nwt + k.psib) - fabs(k.psib) * (k.ths - k.thr) * m_log((-k.psib) / nwt) - k.ths * k.psib;
    else
      dm = k.thr * (nwt + k.psib) - k.ths * k.psib - (k.ths - k.thr) / (k.lam - 1.0) * (k.psib + nwt * m_pow((-k.psib / nwt), k.lam));
  }
  return dm;
}

// water-table depth whose initial profile has moisture deficit dm (3937-4020)
TRIBS_HD_NOINLINE double solve_wt_soil(SoilK k, double dm, double nwt) {
  const double kDmin = 1.0E-5, kDxTol = 1.0E-5;
  double c1, c2, c3, fvalue = 0, fdvalue = 0, x = 0, dx = 0, xinit, xup, est = 0;
  c1 = k.ths - k.thr;
  c2 = c1 * m_pow((-k.psib), k.lam) / (k.lam - 1.0);
  c3 = c1 * k.psib / (k.lam - 1.0) + k.psib * c1 - dm;
  if (nwt == 0.0) {
    xinit = dm * (k.lam - 1.0) / ((k.ths - k.thr) * k.lam) - k.psib;
    xinit += 50.0;
    if (xinit <= fabs(k.psib)) xinit = fabs(k.psib) + 50.0;
  } else
    xinit = nwt;
  if (k.lam < 5.0) {
    if ((k.ths * nwt - soil_total_moist(k, nwt)) > dm) xup = ceil(nwt);
    else xup = k.dbr + 1.0;
    x = soil_bracketed_root(k, c1, c2, c3, dm / (k.ths - k.thr) + fabs(k.psib), xup, kDxTol, xinit);
  }
  if (k.lam >= 5.0 || fabs(x - xinit) <= 1.0E-7) {
    int i;
    for (x = xinit, i = 0; i < 30; x -= dx, i++) {
      soil_poly_wt(k, x, fvalue, fdvalue, c1, c2, c3);
      if (fabs(fdvalue) < kDmin) { est = x; break; }
      dx = fvalue / fdvalue;
      if (fabs(dx) < kDxTol) { est = x; break; }
    }
    if ((fabs(fdvalue) < kDmin) || (fabs(dx) < kDxTol)) x = est;
    else x = nwt;
  }
  if (x > k.dbr) x = k.dbr;
  return x;
}

// wetting-front depth after dm is taken from (flag 0) / added to (flag 1) the wedge (4037-4082)
TRIBS_HD_NOINLINE double solve_front_soil(SoilK k, double dm, double nwt, double nf, int flag) {
  double c1, c2, c3, fvalue, fdvalue, x, dx = 0;
  if (!flag) c1 = k.ths - k.thr; else c1 = -(k.ths - k.thr);
  c2 = c1 * m_pow((-k.psib), k.lam) / (k.lam - 1);
  c3 = c2 / m_pow((nwt - nf), (k.lam - 1)) - c1 * nf + dm;
  int i;
  for (x = nf, i = 0; i < 30; x -= dx, i++) {
    fvalue = c1 * x * m_pow((nwt - x), (k.lam - 1.0)) + c3 * m_pow((nwt - x), (k.lam - 1.0)) - c2;
    fdvalue = c1 * m_pow((nwt - x), (k.lam - 1.0)) - c1 * (k.lam - 1.0) * x * m_pow((nwt - x), (k.lam - 2.0)) -
              c3 * (k.lam - 1.0) * m_pow((nwt - x), (k.lam - 2.0));
    dx = fvalue / fdvalue;
    if (fabs(dx) < 1.0E-5) return x;
  }
  return nf;
}

TRIBS_HD double Column::solve_wt(double dm, double nwt) const {
  return solve_wt_soil(SoilK{ths, thr, psib, lam, dbr}, dm, nwt);
}
TRIBS_HD double Column::solve_front(double dm, double nwt, double nf, int flag) const {
  return solve_front_soil(SoilK{ths, thr, psib, lam, dbr}, dm, nwt, nf, flag);
}

// ------------------------------------------------------------------ Lambert W (principal branch)
struct Cx { double r, i; };
TRIBS_HD Cx cx_sub(Cx a, Cx b) { return Cx{a.r - b.r, a.i - b.i}; }
TRIBS_HD Cx cx_mul(Cx a, Cx b) { return Cx{a.r * b.r - a.i * b.i, a.i * b.r + a.r * b.i}; }
TRIBS_HD Cx cx_addr(Cx a, double d) { return Cx{a.r + d, a.i}; }
TRIBS_HD Cx cx_scale(double x, Cx a) { return Cx{x * a.r, x * a.i}; }
// the reference's quotient rounds its two temporaries to float (mathutil.cpp:1603-1616)
TRIBS_HD Cx cx_div(Cx a, Cx b) {
  Cx c;
  float r, den;
  if (fabs(b.r) >= fabs(b.i)) {
    r = (float)(b.i / b.r);
    den = (float)(b.r + r * b.i);
    c.r = (a.r + r * a.i) / den;
    c.i = (a.i - r * a.r) / den;
  } else {
    r = (float)(b.r / b.i);
    den = (float)(b.i + r * b.r);
    c.r = (a.r * r + a.i) / den;
    c.i = (a.i * r - a.r) / den;
  }
  return c;
}

TRIBS_HD_NOINLINE double lambert_w(double z) {
  Cx w, v, tt, ttt, t, p;
  double xx, xy = 0.0, c, fq;
  if (z == 0.0) return 0.0;
  xx = 2.0 * m_exp(1.0) * z + 2.0;
  if (xx < 0.0) { xx = fabs(xx); xx = sqrt(xx); w = Cx{-1.0, xx}; }
  else { xx = sqrt(xx) - 1.0; w = Cx{xx, 0}; }
  xx = fabs(z);
  xx = m_log(xx);
  if (z < 0.0) v = Cx{xx, kRefPi}; else v = Cx{xx, 0};
  if (z > 1.0) {
    xy = m_log(xx);
    tt = Cx{xy, 0};
    v = cx_sub(v, tt);
  } else if (z == 1.0) {
    tt = Cx{0, 0};
  } else if (z < 1.0 && z > 0.0) {
    xy = m_log(fabs(xx));
    tt = Cx{xy, kRefPi};
    v = cx_sub(v, tt);
  } else if (z < 0.0 && z > -1.0) {
    xx = atan(v.i / v.r);
    xy = m_log(fabs(v.r / cos(xx)));
    xx += kRefPi;
    tt = Cx{xy, xx};
    v = cx_sub(v, tt);
  }
  c = fabs(z + 1.0 / m_exp(1.0));
  if (c > 1.45) w = v;
  int n = 0;
  t = Cx{1000, 1000};
  // loop test keeps the reference's && / || precedence (3893)
  while (((n < 15) && (fabs(t.r) > xy)) || (fabs(t.i) > xx)) {
    xx = m_exp(w.r);
    p = Cx{xx * cos(w.i), xx * sin(w.i)};
    t = cx_mul(w, p);
    t.r = t.r - z;
    if (w.i == 0 && w.r == -1) fq = 0.0; else fq = 1.0;
    tt = cx_mul(p, cx_addr(w, fq));
    ttt = cx_sub(tt, cx_div(cx_mul(t, cx_scale(0.5, cx_addr(w, 2.0))), cx_addr(w, fq)));
    t = cx_div(cx_scale(fq, t), ttt);
    w = cx_sub(w, t);
    xy = (2.48 * kLambEps) * (1.0 + fabs(w.r));
    xx = (2.48 * kLambEps) * (1.0 + fabs(w.i));
    n++;
    if (n > 4096) break;  // the reference can spin here on |Im t|; bounded on the device
  }
  return w.r;
}

// ------------------------------------------------------------------ unsaturated zone, one column
struct UnsatEnv {
  double dt;            // h
  double int_storm_max; // h
  int rdstr_option;
  double vel_mm;        // Voronoi width of the flow edge, mm
  double dest_nwt_old;  // downslope cell, values at step start (0 for the outlet node)
  double dest_nf_old;
};

// Classifies the column and advances it by dt. On entry: soil, Old state (both copies equal),
// qpin/qpout in mm/h, isv, r (= Ractual), r1, cosv/sinv. On exit: New state, qpout
// (mm^3/h, already * sin), isv, hsrf/sbsrf/psrf on the slope plane (not yet / cos).
TRIBS_HD int unsat_advance(Column &c, const UnsatEnv &e) {
  const double dt = e.dt;
  // inputs of the surface conductivity Kunsat (reference 860-869) as they are on entry; it is
  // only needed when a new front is born, so it is evaluated there
  const double ru0_in = c.ru0, nwt0_in = c.nwt0;
  double qn = 0, xxsrf, mdelt, mdva, aa, bb, mperch, nstar;
  int state = -1000;

  if ((c.nwt0 == 0.) && (c.r1 >= 0.)) state = kWtStaysAtSurf;
  if ((c.nwt0 == 0.) && (c.r1 < 0.)) state = kWtDropsFromSurf;
  if ((c.nwt0 > 0.) && ((c.r1 * dt) >= (c.nwt0 * c.ths - c.mu0)) && (c.nf0 == 0. || c.nf0 == c.nwt0))
    state = kWtGetsToSurf;
  if (state == -1000) {
    c.mu1 = c.mu0 + dt * c.r1;
    if (c.mu1 < c.mi0) state = kIntStormBelow;
    if (fabs(c.mu1 - c.mi0) < 1.0E-6) state = kExactInitial;
    if (c.mu1 > c.mi0) {
      if (c.r1 > 0.0) {
        if (c.r > kRadar) {
          // a long dry spell ended: dissolve the old wedge into the profile (902-925)
          if (c.isv >= e.int_storm_max && c.nf0 > 0.0) {
            if (c.mu1 >= c.nwt0 * c.ths) {
              c.nwt1 = 0.0;
              c.mi1 = 0.0;
              c.sbsrf = (c.mu1 - c.nwt0 * c.ths) / dt;
            } else {
              mdelt = c.ths * c.nwt0 - c.mu0;
              c.nwt1 = c.solve_wt(mdelt, c.nwt0);
              c.mi1 = c.total_moist(c.nwt1);
            }
            c.nwt0 = c.nwt1;
            c.mi0 = c.mu0 = c.mu1 = c.mi1;
            c.nt0 = c.nt1 = 0.0;
            c.nf0 = c.nf1 = 0.0;
            c.ri0 = c.ri1 = 0.0;
            c.ru0 = c.ru1 = 0.0;
            c.isv = 0.0;
          }
        }
        if (c.nt0 == c.nf0) state = kStormUnsatEvol;
        if ((c.nt0 < c.nf0) && (c.nt0 != 0.0)) state = kPerchedEvol;
        if ((c.nt0 == 0.0) && (c.nf0 > 0.0)) {
          c.thri = c.init_moist_z(c.nf0);
          c.thre = c.ths;
          c.suction(c.nf0);
          qn = c.ks * c.f * c.nf0 / m_expm1(c.f * c.nf0) * (c.cosv + c.g / c.nf0);
          xxsrf = qn - c.ri0 * c.cosv;
          if (c.r1 >= xxsrf) state = kPerchedSurfSat;
          else state = kStormToInter;
        }
        if ((c.nf0 == c.nwt0) || (c.r <= kRadar && c.isv > 0 && c.nf0 == 0.0)) state = kStormEvol;
      }
      if (c.r1 <= 0.0) state = kStormToInter;
    }
    if (c.ks == 0.0) state = kPerchedSurfSat;
  }

  const double rflux = (c.r + (c.qpin - c.qpout)) * c.cosv;  // the cases' own R1 (no runon term)

  switch (state) {
    case kWtStaysAtSurf:
      if (c.r > 0.0) c.sbsrf = c.r * c.cosv;
      if (c.qpin > c.qpout) c.psrf = (c.qpin - c.qpout) * c.cosv;
      else c.sbsrf -= (c.qpout - c.qpin) * c.cosv;
      c.mi1 = c.mu1 = c.ri1 = c.ru1 = c.nf1 = c.nt1 = c.nwt1 = 0.0;
      c.qpout = 0.0;
      if (c.r <= kRadar) c.isv += dt;
      else if (c.isv > 0.0) c.isv = 0.0;
      break;

    case kWtGetsToSurf:
      if (c.qpin > c.qpout) {
        c.psrf = (c.qpin - c.qpout) * c.cosv + c.mu0 / dt - c.nwt0 * c.ths / dt;
        if (c.psrf < 0.0) { c.sbsrf = c.r * c.cosv + c.psrf; c.psrf = 0.0; }
        else if (c.r > 0.0) c.sbsrf = c.r * c.cosv;
      } else
        c.sbsrf = (c.r + (c.qpin - c.qpout)) * c.cosv + c.mu0 / dt - c.nwt0 * c.ths / dt;
      c.mu1 = c.ri1 = c.ru1 = c.nf1 = c.nt1 = c.nwt1 = c.mi1 = 0.0;
      c.qpout = 0.0;
      if (c.isv > 0.0) c.isv = 0.0;
      break;

    case kWtDropsFromSurf:
      c.r1 = rflux;
      if (c.r1 < 0.0) c.nwt1 = c.solve_wt(fabs(c.r1 * dt), 0.0);
      c.settle_profile();
      c.ri1 = 0.0; c.ru1 = 0.0;
      c.qpout = 0.0;
      c.isv += dt;
      if (c.isv >= e.int_storm_max) { c.nf1 = 0.0; c.nt1 = 0.0; }
      else { c.nf1 = c.nwt1; c.nt1 = c.nwt1; }
      break;

    case kExactInitial:
      c.mi1 = c.mi0;
      c.mu1 = c.mi1;
      c.nwt1 = c.nwt0;
      c.ri1 = 0.0; c.ru1 = 0.0;
      c.qpout = 0.0;
      if (c.r <= 0.0) c.isv += dt;
      if (c.nf0 == c.nwt0) {
        if (c.isv >= e.int_storm_max) { c.nf1 = 0.0; c.nt1 = 0.0; }
        else { c.nf1 = c.nwt1; c.nt1 = c.nwt1; }
      } else { c.nf1 = 0.0; c.nt1 = 0.0; }
      break;

    case kIntStormBelow:
      c.r1 = rflux;
      mdelt = c.ths * c.nwt0 - (c.mu0 + c.r1 * dt);
      c.nwt1 = c.solve_wt(mdelt, c.nwt0);
      c.ri1 = 0.0; c.ru1 = 0.0;
      c.settle_profile();
      c.isv += dt;
      if ((c.nf0 < c.nwt0) && (c.nf0 > 0.0)) {
        c.nf1 = c.nt1 = 0.0;
      } else if (c.isv < e.int_storm_max) {
        if (c.nf0 == 0.0) c.nf1 = c.nt1 = 0.0;
        else c.nf1 = c.nt1 = c.nwt1;
      } else
        c.nf1 = c.nt1 = 0.0;
      c.qpout = 0.0;
      break;

    case kStormToInter: {
      if (c.r <= kRadar) c.isv += dt; else c.isv = 0.0;
      c.r1 = rflux;
      c.mu1 = c.mu0 + c.r1 * dt;
      c.nwt1 = c.nwt0;
      c.mi1 = c.mi0;
      c.ri1 = c.ri0;
      if (c.isv >= e.int_storm_max && e.rdstr_option) {
        if (c.mu1 >= c.nwt0 * c.ths) {
          c.nwt1 = 0.0;
          c.sbsrf = (c.mu1 - c.nwt0 * c.ths) / dt;
          c.mi1 = 0.0;
        } else {
          mdelt = c.ths * c.nwt0 - c.mu1;
          c.nwt1 = c.solve_wt(mdelt, c.nwt0);
          c.mi1 = c.total_moist(c.nwt1);
        }
        c.mu1 = c.mi1;
        c.nt1 = 0.0; c.nf1 = 0.0; c.ri1 = 0.0; c.ru1 = 0.0;
        break;  // note: qpout keeps its incoming (old) value on this path, as in the reference
      }
      mdelt = c.lower_moist(c.nf0, c.nwt1);
      mdelt = c.mu1 - mdelt;  // water held by the wedge
      aa = c.wedge_cap(c.nf0);
      if (mdelt > aa && c.nt0 == c.nf0) mdelt = aa - 1.0E-4;

      if (mdelt <= aa) {  // unsaturated wedge
        if (c.r1 == 0.0) {
          c.ru1 = c.ru0;
          nstar = (m_log(c.ks / c.ru1)) / c.f;
        } else {
          c.ru1 = c.recharge_rate(mdelt, c.nf0);
          nstar = (m_log(c.ks / c.ru1)) / c.f;
        }
        (void)nstar;
        if ((100. * (c.mu1 - c.mi0) / c.mi0) <= 0.05) {
          c.nf1 = c.nt1 = 0.0;
          if ((c.mu1 - c.mi0) > 1.0E-4) {
            mdelt = c.ths * c.nwt0 - c.mu1;
            c.nwt1 = c.solve_wt(mdelt, c.nwt0);
          }
          c.settle_profile();
          c.ri1 = c.ru1 = 0.0;
        } else {
          c.thri = c.init_moist_z(c.nf0);
          c.thre = c.edge_moist_z(c.nf0);
          if (c.thre <= c.thri) {
            mdelt = c.ths * c.nwt0 - (c.mu0 + c.r1 * dt);
            c.nwt1 = c.solve_wt(mdelt, c.nwt0);
            c.nt1 = c.nf1 = 0.0;
            c.ru1 = c.ri1 = 0.0;
            c.settle_profile();
            c.qpout = 0.0;
          } else {
            c.suction(c.nf0);
            qn = c.ru1 * c.cosv + c.ks * m_exp(-c.f * c.nf0) * c.g / c.nf0;
            c.nf1 = c.nf0 + dt * (qn - c.ri1 * c.cosv) / (c.thre - c.thri);
            c.nt1 = c.nf1;
            if (c.nf1 < (c.nwt1 + c.psib)) {
              mdelt = c.lower_moist(c.nf1, c.nwt1);
              mdelt = c.mu1 - mdelt;
              c.ru1 = c.recharge_rate(mdelt, c.nf1);
              aa = c.wedge_cap(c.nf1);
              if (mdelt > aa) {
                c.nt1 = (m_log(c.ks / c.ru1)) / c.f;
                c.ru1 = c.ks * m_exp(-c.f * c.nt1);
              }
            } else if (c.nf1 >= (c.nwt1 + c.psib)) {
              if ((c.mu1 - c.mi0) > 1.0E-4) {
                mdelt = c.ths * c.nwt0 - c.mu1;
                c.nwt1 = c.solve_wt(mdelt, c.nwt0);
              }
              c.settle_profile();
              c.ri1 = c.ru1 = 0.0;
              if ((e.dest_nf_old == e.dest_nwt_old) && (c.isv < e.int_storm_max)) c.nf1 = c.nt1 = c.nwt1;
              else c.nf1 = c.nt1 = 0.0;
              c.ru1 = 0.0;
            }
          }
        }
      } else {  // perched or surface-saturated wedge
        c.thri = c.init_moist_z(c.nf0);
        c.thre = c.ths;
        c.suction(c.nf0);
        qn = c.ks * c.f * (c.nf0 - c.nt0) / (m_exp(c.f * c.nf0) - m_exp(c.f * c.nt0));
        qn *= (c.cosv + c.g / c.nf0);
        c.nf1 = c.nf0 + dt * (qn - c.ri1 * c.cosv) / (c.ths - c.thri);
        if (fabs(c.nf1 - (c.nwt1 + c.psib)) <= 1.0E-3) {
          c.nwt1 = c.nt0 - c.psib;
          if ((e.dest_nf_old == e.dest_nwt_old) && (c.isv < e.int_storm_max)) c.nf1 = c.nwt1;
          else c.nf1 = 0.0;
        } else if (c.nf1 > (c.nwt1 + c.psib)) {
          c.nwt1 = c.nt0 - c.psib;
          if ((e.dest_nf_old == e.dest_nwt_old) && (c.isv < e.int_storm_max)) c.nf1 = c.nwt1;
          else c.nf1 = 0;
        }
        if ((c.nf1 == 0.0) || (c.nf1 == c.nwt1)) {
          if (c.mu1 >= c.nwt0 * c.ths) {
            c.nwt1 = 0.0;
            c.nt1 = c.nf1 = 0.0;
            c.sbsrf = (c.mu1 - c.nwt0 * c.ths) / dt;
            c.mu1 = c.mi1 = 0.0;
            c.ru1 = c.ri1 = 0.0;
          } else {
            mdelt = c.ths * c.nwt0 - c.mu1;
            c.nwt1 = c.solve_wt(mdelt, c.nwt0);
            c.ri1 = c.ru1 = 0.0;
            c.settle_profile();
            if (c.nf1 == c.nwt1) c.nf1 = c.nt1 = c.nwt1;
            else c.nf1 = c.nt1 = 0.0;
          }
        } else if ((c.nf1 > 0.0) && (c.nf1 != c.nwt1)) {
          mdelt = c.wedge_cap(c.nt0);
          mdva = c.wedge_cap(c.nf1);
          aa = (mdelt + (c.nf1 - c.nt0) * c.ths - mdva);
          if (c.r1 < (qn - aa / dt)) {
            mperch = c.lower_moist(c.nf1, c.nwt1);
            mperch = c.mu1 - mperch;
            c.ru1 = c.recharge_rate(mperch, c.nf1);
            c.nt1 = c.nf1;
            if (mdelt >= mdva) c.mu1 -= mdelt - mdva + 1.0E-4;
          } else {
            bb = -dt * (qn - c.r1) / (c.ths - c.thr) - c.eps / c.f * m_exp(-c.f * c.nt0 / c.eps);
            aa = -m_exp(c.f * (bb - c.nt0) / c.eps);
            c.nt1 = c.nt0 + (c.eps / c.f * lambert_w(aa) - bb);
            c.ru1 = c.ks * m_exp(-c.f * c.nt1);
          }
        }
      }
      c.wedge_outflow(e.vel_mm);
      break;
    }

    case kStormEvol:
      c.r1 = rflux;
      mdelt = c.ths * c.nwt0 - (c.mu0 + c.r1 * dt);
      c.nwt1 = c.solve_wt(mdelt, c.nwt0);
      c.ri1 = c.ru1 = 0.0;
      c.settle_profile();
      c.qpout = 0.0;
      if (c.r <= kRadar) {
        c.isv += dt;
        if (c.nf0 == c.nwt0) {
          if (c.isv >= e.int_storm_max) c.nf1 = c.nt1 = 0.0;
          else c.nf1 = c.nt1 = c.nwt1;
        } else
          c.nf1 = c.nt1 = 0.0;
      } else {
        if (c.isv > 0.0) c.isv = 0.0;
        c.nf1 = c.nt1 = c.nwt1;
      }
      break;

    case kStormUnsatEvol: {
      if (c.r <= kRadar) c.isv += dt; else c.isv = 0.0;
      c.r1 = rflux;
      c.nwt1 = c.nwt0;
      c.mu1 = c.mu0 + c.r1 * dt;
      c.mi1 = c.mi0;
      c.ri1 = c.ri0 = 0.0;
      if (c.nf0 == 0.0 && c.nt0 == 0.0) {  // a wetting front is born
        double kunsat = ru0_in;
        if (ru0_in == 0.0) {
          const double thsurf = (c.ks != 0.0)
                                    ? ((0. >= (nwt0_in + c.psib)) ? c.ths : (c.thr + (c.ths - c.thr) * m_pow((c.psib / (0. - nwt0_in)), c.lam)))
                                    : 0.0;
          kunsat = c.ks * m_pow(((thsurf - c.thr) / (c.ths - c.thr)), c.eps);
        }
        c.thri = c.init_moist_z(0.0);
        if (c.r < kunsat && c.r1 / c.cosv < kunsat) {
          c.thre = m_pow((c.r1 / c.ks), (1 / c.eps)) * (c.ths - c.thr) + c.thr;
          qn = c.r1;
          c.nf1 = dt * qn / (c.thre - c.thri);
        } else if (c.r1 < c.ks) {
          c.thre = c.thr + (c.ths - c.thr) * m_pow((c.r1 / c.ks), (1 / c.eps));
          c.suction(0.0);
          if (fabs(c.thre - c.thri) < 1.0E-4) qn = c.r1;
          else qn = c.ks * m_exp(-c.f * c.r1 / (c.thre - c.thri)) * c.g * (c.thre - c.thri) / c.r1 + c.r1;
          c.nf1 = dt * qn / (c.ths - c.thri);
        } else {
          c.thre = c.ths;
          c.suction(0.0);
          qn = c.ks * c.g * (c.thre - c.thri) / c.ks + c.r1;
          c.nf1 = dt * qn / (c.thre - c.thri);
        }
        if (c.nf1 > 0.0 && c.nf1 < 32000.) {
          if (c.nf1 < 1.0) c.nf1 = 1.0;
          c.nt1 = c.nf1;
        } else
          c.nt1 = c.nf1 = c.nwt0 + 100.;

        if (c.nf1 < c.nwt0) {
          mdelt = c.lower_moist(c.nf1, c.nwt1);
          mdelt = c.mu1 - mdelt;
          aa = c.wedge_cap(c.nf1);
          if (mdelt >= aa) {
            if (fabs(mdelt - aa) <= 1.0E-6) {
              c.nt1 = c.nf1;
              c.nf1 += 1.0E-5;
              c.ru1 = c.ks * m_exp(-c.f * c.nt1);
            } else if (mdelt >= c.nf1 * c.ths) {
              c.thri = c.init_moist_z(c.nf1);
              c.nf1 += (mdelt - c.nf1 * c.ths) / (c.ths - c.thri);
              if (c.nf1 >= (c.nwt1 + c.psib)) {
                c.nf1 = c.nt1 = c.nwt1 = 0.0;
                c.mu1 = c.mi1 = 0.0;
                c.ru1 = c.ri1 = 0.0;
              } else {
                c.nt1 = 0.0;
                c.ru1 = c.ksat_mean(c.nf1);
                aa = c.lower_moist(c.nf1, c.nwt1);
                bb = c.nf1 * c.ths;
                c.mu1 = aa + bb;
              }
            } else if (mdelt < c.nf1 * c.ths) {
              c.ru1 = c.recharge_rate(mdelt, c.nf1);
              c.nt1 = (m_log(c.ks / c.ru1)) / c.f;
              c.nf1 += 1.0E-5;
            }
          } else
            c.ru1 = c.recharge_rate(mdelt, c.nf1);
          if ((c.ri1 - c.ru1) < 1.0E-2 && (c.ri1 - c.ru1) > 0.0) c.ru1 = c.ri1;
        }
      } else if (c.nf0 > 0.0 && c.nt0 > 0.0) {  // fronts already below the surface
        mdelt = c.lower_moist(c.nf0, c.nwt1);
        mdelt = c.mu1 - mdelt;
        if (mdelt >= c.nf0 * c.ths) {
          c.thri = c.init_moist_z(c.nf0);
          c.thre = c.ths;
          c.suction(c.nf0);
          qn = c.ks * c.f * c.nf0 / m_expm1(c.f * c.nf0);
          if ((qn * (c.cosv + c.g / c.nf0) - c.ri0 * c.cosv) <= c.r1) {
            qn *= (c.cosv + c.g / c.nf0);
            if (c.r1 >= (qn - c.ri0 * c.cosv)) {
              c.hsrf += (c.r1 - (qn - c.ri0 * c.cosv));
              c.r1 -= c.hsrf;
            }
            c.nt1 = 0.0;
            c.mu1 = c.mu0 + c.r1 * dt;
            c.nf1 = c.nf0 + dt * (qn - c.ri1 * c.cosv) / (c.ths - c.thri);
            c.ru1 = c.ksat_mean(c.nf1);
            if ((fabs(c.nf1 - (c.nwt1 + c.psib)) <= 1.0E-3) || (c.nf1 > (c.nwt1 + c.psib))) {
              c.nwt1 = c.mi1 = 0.0;
              c.ru1 = c.ri1 = 0.0;
              mperch = c.mu1 - c.nwt0 * c.ths;
              if (mperch > 0.0) c.sbsrf += mperch / dt;
              c.mu1 = 0.0;
              c.nt1 = c.nf1 = 0.0;
            }
          } else {
            c.thre = c.edge_moist_z(c.nf0);
            c.suction(c.nf0);
            qn = c.ru0 * c.cosv + c.ks * m_exp(-c.f * c.nf0) * c.g / c.nf0;
            c.nf1 = c.nf0 + dt * (qn - c.ri1 * c.cosv) / (c.ths - c.thri);
            c.nt1 = c.nf1;
            if (c.nf1 < (c.nwt1 + c.psib)) {
              mdelt = c.lower_moist(c.nf1, c.nwt1);
              mdelt = c.mu1 - mdelt;
              c.ru1 = c.recharge_rate(mdelt, c.nf1);
            } else if (c.mu1 > c.ths * c.nwt1) {
              c.sbsrf += (c.mu1 - c.ths * c.nwt1);
              c.nf1 = c.nt1 = c.nwt1 = 0.0;
              c.mu1 = c.mi1 = 0.0;
              c.ru1 = c.ri1 = 0.0;
            }
          }
        } else {
          c.ru1 = c.recharge_rate(mdelt, c.nf0);
          nstar = (m_log(c.ks / c.ru1)) / c.f;
          if (nstar <= c.nf0) {
            c.nt1 = nstar;
            c.nf1 = c.nf0 + 1.0E-5;
          } else {
            c.thri = c.init_moist_z(c.nf0);
            c.thre = c.edge_moist_z(c.nf0);
            if (c.thre <= c.thri) {
              mdelt = c.ths * c.nwt0 - (c.mu0 + c.r1 * dt);
              c.nwt1 = c.solve_wt(mdelt, c.nwt0);
              c.ru1 = c.ri1 = 0.0;
              c.settle_profile();
              c.qpout = 0.0;
              if (c.nwt0 < 2.0 * fabs(c.psib)) c.nt1 = c.nf1 = c.nwt1;
              else c.nt1 = c.nf1 = 0.0;
            } else {
              c.suction(c.nf0);
              qn = c.ru1 * c.cosv + c.ks * m_exp(-c.f * c.nf0) * c.g / c.nf0;
              c.nf1 = c.nf0 + dt * (qn - c.ri1 * c.cosv) / (c.thre - c.thri);
              c.nt1 = c.nf1;
              if (c.nf1 < (c.nwt1 + c.psib)) {
                mdelt = c.lower_moist(c.nf1, c.nwt1);
                mdelt = c.mu1 - mdelt;
                c.ru1 = c.recharge_rate(mdelt, c.nf1);
              }
            }
          }
        }
      }
      if ((c.nf1 > (c.nwt1 + c.psib)) && (c.nf1 != c.nwt1)) {  // front overran the fringe (1648)
        mdelt = c.ths * c.nwt1 - c.mu1;
        c.nwt1 = c.solve_wt(mdelt, c.nwt0);
        c.ri1 = c.ru1 = 0.0;
        c.mu1 = c.mi1 = c.total_moist(c.nwt1);
        c.nf1 = c.nt1 = c.nwt1;
      }
      c.wedge_outflow(e.vel_mm);
      break;
    }

    case kPerchedEvol: {
      if (c.r <= kRadar) c.isv += dt; else c.isv = 0.0;
      c.r1 = rflux;
      c.nwt1 = c.nwt0;
      c.mu1 = c.mu0 + c.r1 * dt;
      c.mi1 = c.mi0;
      c.ri1 = c.ri0;
      c.thri = c.init_moist_z(c.nf0);
      c.thre = c.ths;
      c.suction(c.nf0);
      qn = c.ks * c.f * (c.nf0 - c.nt0) / (m_exp(c.f * c.nf0) - m_exp(c.f * c.nt0));
      qn *= (c.cosv + c.g / c.nf0);
      c.nf1 = c.nf0 + dt * (qn - c.ri1 * c.cosv) / (c.ths - c.thri);
      if (fabs(c.nf1 - (c.nwt1 + c.psib)) <= 1.0E-3) {
        c.nwt1 = c.nt0 - c.psib;
        c.nf1 = c.nwt1;
      } else if (c.nf1 > (c.nwt1 + c.psib)) {
        c.nwt1 = c.nt0 - c.psib;
        c.nf1 = c.nwt1;
        c.r1 += qn - ((c.nwt0 + c.psib - c.nf0) * (c.ths - c.thri) / dt + c.ri1 * c.cosv);
      }
      if (c.nf1 == c.nwt1) {
        if (c.mu1 >= c.nwt0 * c.ths) {
          c.nwt1 = c.nt1 = c.nf1 = 0.0;
          c.sbsrf = (c.mu1 - c.nwt0 * c.ths) / dt;
          c.mu1 = c.mi1 = 0.0;
          c.ru1 = c.ri1 = 0.0;
        } else {
          mdelt = c.ths * c.nwt0 - c.mu1;
          c.nwt1 = c.solve_wt(mdelt, c.nwt0);
          c.ri1 = c.ru1 = 0.0;
          c.settle_profile();
          c.nt1 = c.nf1 = c.nwt1;
        }
      } else {
        mdelt = c.wedge_cap(c.nt0);
        mdva = c.wedge_cap(c.nf1);
        aa = (mdelt + (c.nf1 - c.nt0) * c.ths - mdva);
        if ((c.r1 + c.ri1 * c.cosv) < (qn - aa / dt)) {
          mperch = c.lower_moist(c.nf1, c.nwt1);
          mperch = c.mu1 - mperch;
          c.ru1 = c.recharge_rate(mperch, c.nf1);
          c.nt1 = c.nf1;
        } else {
          xxsrf = (c.r1 - qn) * dt + mdelt;
          if (xxsrf >= c.nt0 * c.ths) {
            c.nt1 = 0.0;
            c.hsrf += (xxsrf - c.nt0 * c.ths) / dt;
            c.mu1 -= c.hsrf * dt;
            c.ru1 = c.ksat_mean(c.nf1);
          } else {
            bb = -dt * (qn - c.r1) / (c.ths - c.thr) - c.eps / c.f * m_exp(-c.f * c.nt0 / c.eps);
            aa = -m_exp(c.f * (bb - c.nt0) / c.eps);
            c.nt1 = c.nt0 + (c.eps / c.f * lambert_w(aa) - bb);
            c.ru1 = c.ks * m_exp(-c.f * c.nt1);
          }
          if (c.nwt1 != 0.0) {
            aa = c.lower_moist(c.nf1, c.nwt1);
            bb = (c.nf1 - c.nt1) * c.ths + c.eps / c.f * (c.ths - c.thr) * (1.0 - m_exp(-c.f * c.nt1 / c.eps)) +
                 c.thr * c.nt1;
            if (c.mu1 > (aa + bb)) {
              c.thri = c.init_moist_z(c.nf1);
              c.nf1 += (c.mu1 - (aa + bb)) / (c.ths - c.thri);
              if (c.nf1 >= (c.nwt1 + c.psib)) {
                if (c.nt1 > 0.0) {
                  c.nwt1 = c.nt1 - c.psib;
                  c.nf1 = c.nt1 = c.nwt1;
                  c.mu1 = c.mi1 = c.total_moist(c.nwt1);
                  c.ri1 = c.ru1 = 0.0;
                } else if (c.nt1 == 0.0) {
                  c.nwt1 = c.nf1 = c.nt1 = 0.0;
                  c.mu1 = c.mi1 = 0.0;
                  c.ri1 = c.ru1 = 0.0;
                }
              } else {
                c.mu1 = aa + bb;
                if (c.nt1 == 0.0) c.ru1 = c.ksat_mean(c.nf1);
              }
            }
          }
        }
      }
      c.wedge_outflow(e.vel_mm);
      break;
    }

    case kPerchedSurfSat: {
      if (c.r <= kRadar) c.isv += dt; else c.isv = 0.0;
      if (c.ks == 0.0) {
        c.mu1 = c.mi1 = 0.0;
        c.ri1 = c.ru1 = 0.0;
        c.nf1 = c.nt1 = 0.0;
        c.nwt1 = c.nwt0;
        c.qpout = 0.0;
        c.hsrf += c.r * c.cosv;
        break;
      }
      c.nwt1 = c.nwt0;
      if ((c.nt0 == 0.0) && (c.nf0 > 0.0) && (c.nf0 != c.nwt0)) {
        c.thri = c.init_moist_z(c.nf0);
        c.thre = c.ths;
        c.suction(c.nf0);
        qn = c.ks * c.f * c.nf0 / m_expm1(c.f * c.nf0);
        qn *= (c.cosv + c.g / c.nf0);
        c.r1 = rflux;
        if (c.r1 >= qn) {
          c.hsrf += (c.r1 - qn);
          c.r1 -= c.hsrf;
        }
        c.nt1 = 0.0;
        c.ri1 = c.ri0;
        c.mi1 = c.mi0;
        c.mu1 = c.mu0 + c.r1 * dt;
        c.nf1 = c.nf0 + dt * (qn - c.ri1 * c.cosv) / (c.ths - c.thri);
      }
      if (c.nf1 >= 1.0) {
        c.ru1 = c.ksat_mean(c.nf1);
        if (c.ru1 > c.ks) c.ru1 = c.ks;
      } else
        c.ru1 = c.ks;
      if ((fabs(c.nf1 - (c.nwt1 + c.psib)) <= 1.0E-3) || (c.nf1 > (c.nwt1 + c.psib))) {
        mperch = c.mu1 - c.nwt0 * c.ths;
        if (mperch >= 0.0) {
          c.sbsrf += mperch / dt;
          c.nf1 = c.nt1 = c.nwt1 = 0.0;
          c.mi1 = c.mu1 = 0.0;
          c.ri1 = c.ru1 = 0.0;
        } else {
          c.nwt1 = c.solve_wt(fabs(mperch), 0.0);
          c.mu1 = c.mi1 = c.total_moist(c.nwt1);
          c.ri1 = c.ru1 = 0.0;
          c.nf1 = c.nt1 = c.nwt1;
        }
      }
      if (c.nwt1 != 0.0 && c.nf1 >= 1.0 && c.nf1 != c.nwt1) {
        aa = c.lower_moist(c.nf1, c.nwt1);
        bb = c.nf1 * c.ths;
        if (c.mu1 >= (aa + bb)) {
          c.thri = c.init_moist_z(c.nf1);
          c.nf1 += (c.mu1 - (aa + bb)) / (c.ths - c.thri);
          if (c.nf1 >= (c.nwt1 + c.psib)) {
            if (c.mu1 >= (c.nwt0 * c.ths)) c.sbsrf += (c.mu1 - c.nwt0 * c.ths);
            c.nwt1 = 0.0;
            c.ri1 = c.ru1 = 0.0;
            c.mi1 = c.mu1 = 0.0;
            c.nf1 = c.nt1 = 0.0;
          } else {
            c.mu1 = aa + bb;
            c.ru1 = c.ksat_mean(c.nf1);
          }
        }
      }
      c.qpout = 0.0;
      if ((c.nf1 - c.nt1) > 10.0) {
        c.qpout = c.sat_lateral(c.nt1, c.nf1, c.ru1, e.vel_mm);
        c.qpout *= c.sinv;
      }
      break;
    }
    default:
      break;
  }
  return state;
}

// ------------------------------------------------------------------ saturated zone, one column
// On entry: soil, Old state (both copies equal), cosv, gw = sorted sum of edge fluxes (mm^3/h).
// On exit: New state and satsrf (slope plane, not yet / cos).
TRIBS_HD int sat_advance(Column &c, double gw, double area, double dtgw) {
  double mdelt, dm1, dm2, dm3 = 0, nstar;
  c.satsrf = 0.0;
  c.nwt1 = c.nwt0 + dtgw * (gw * 1.0E-6) * c.cosv / (area * c.ths);
  if (c.nwt1 < 0.0) {
    c.satsrf = fabs(c.nwt1 * c.ths) / dtgw;
    c.nwt1 = 0.0;
  }
  int st = -1000;
  if (c.nwt1 == c.nwt0) st = kGwInitial;
  else if (c.nwt1 - c.nwt0 > 1.0E-3) st = kGwIntStormLike;
  else if (c.nwt0 - c.nwt1 > 1.0E-3) st = kGwPositiveBal;
  if (c.mu0 > c.nwt1 * c.ths) st = kGwExfiltrate;
  if (st == -1000) st = kGwInitial;

  switch (st) {
    case kGwExfiltrate:
      c.satsrf += (c.mu0 - c.nwt1 * c.ths) / dtgw;
      c.nwt1 = 0;
      c.mu1 = c.mi1 = 0.0;
      c.ri1 = c.ru1 = 0.0;
      c.nf1 = c.nt1 = 0.0;
      break;
    case kGwInitial:
      c.nwt1 = c.nwt0; c.mu1 = c.mu0; c.mi1 = c.mi0; c.nf1 = c.nf0;
      c.nt1 = c.nt0; c.ru1 = c.ru0; c.ri1 = c.ri0;
      break;
    case kGwIntStormLike:
      mdelt = (c.nwt1 - c.nwt0) * c.ths;
      if (c.nwt0 > 0.0) c.nwt1 = c.solve_wt((c.ths * c.nwt0 - (c.mi0 - mdelt)), c.nwt1);
      else c.nwt1 = c.solve_wt((c.ths * c.nwt0 - (c.mi0 - mdelt)), 0.0);
      c.mi1 = c.total_moist(c.nwt1);
      c.ri1 = 0.0;
      if ((c.nf0 == 0.0) || (c.nf0 == c.nwt0)) {
        c.ru1 = 0.0;
        c.mu1 = c.mi1;
        if (c.nf0 == 0.0) c.nf1 = c.nt1 = 0.0;
        else c.nf1 = c.nt1 = c.nwt1;
      } else {
        dm1 = c.upper_moist(c.nf0, c.nwt1);
        dm2 = c.mu0 - c.mi0;
        mdelt = dm1 + dm2;
        dm3 = c.wedge_cap(c.nf0);
        c.mu1 = c.mi1 + dm2;
        if (mdelt < dm3) {
          c.nt1 = c.nf1 = c.nf0;
          c.ru1 = c.recharge_rate(mdelt, c.nf0);
        } else {
          mdelt = c.upper_moist(c.nf0, c.nwt0);
          mdelt -= dm1;
          if ((c.nt0 > 0.0) && (c.nt0 < c.nf0)) {
            c.nf1 = c.solve_front(mdelt, c.nwt1, c.nf0, 0);
            if (c.nf1 <= c.nt0) {
              if (!((c.nt0 - c.nf1) >= 10.0)) c.nf1 = c.nt0 + 0.001;
            }
            c.nt1 = c.nt0;
            c.ru1 = c.ru0;
          } else if (c.nt0 == 0.0 && c.nf0 > 0.0) {
            c.nf1 = c.solve_front(mdelt, c.nwt1, c.nf0, 0);
            c.nt1 = c.nt0;
            c.ru1 = c.ksat_mean(c.nf1);
          }
        }
      }
      break;
    case kGwPositiveBal:
      if (((c.nwt1 + c.psib) > c.nf0) || (c.nf0 == c.nwt0)) {
        c.mi1 = c.total_moist(c.nwt1);
        dm1 = c.lower_moist(c.nwt1, c.nwt0);
        dm2 = c.mi0 - dm1;
        mdelt = dm1 - (c.mi1 - dm2);
        c.nwt1 = c.solve_wt((c.ths * c.nwt1 - (c.mi1 + mdelt)), c.nwt1);
        c.mi1 = c.total_moist(c.nwt1);
        if ((c.nf0 == 0.0) || (c.nf0 == c.nwt0)) {
          c.mu1 = c.mi1;
          c.ru1 = c.ri1 = 0.0;
          if (c.nf0 == c.nwt0) c.nf1 = c.nt1 = c.nwt1;
          else c.nf1 = c.nt1 = 0.0;
        } else if ((c.nf0 > 0.0) && (c.nf0 != c.nwt0)) {
          if ((c.nwt1 + c.psib) > c.nf0) {
            mdelt = c.upper_moist(c.nf0, c.nwt1);
            mdelt += (c.mu0 - c.mi0);
            if (c.nt0 == c.nf0) {
              if (mdelt >= c.nf0 * c.ths) {
                mdelt = -dtgw * (gw * 1.0E-6) * c.cosv / area;
                c.nwt1 = c.solve_wt((c.ths * c.nwt0 - (c.mu0 + mdelt)), (c.nt0 - c.psib));
                c.mu1 = c.mi1 = c.total_moist(c.nwt1);
                c.ri1 = c.ru1 = 0.0;
                c.nf1 = c.nt1 = c.nwt1;
              } else {
                c.ru1 = c.recharge_rate(mdelt, c.nf0);
                nstar = (m_log(c.ks / c.ru1)) / c.f;
                if (nstar <= c.nf0) { c.nt1 = nstar; c.nf1 = c.nf0 + 1.0E-5; }
                else c.nf1 = c.nt1 = c.nf0;
                c.mu1 = c.mi1 + (c.mu0 - c.mi0);
                c.ri1 = 0.0;
              }
            } else if (c.nt0 < c.nf0) {
              if (mdelt >= c.nf0 * c.ths) {
                mdelt -= c.nf0 * c.ths;
              } else {
                dm1 = c.upper_moist(c.nf0, c.nwt1);
                dm2 = c.upper_moist(c.nf0, c.nwt0);
                mdelt = dm1 - dm2;
              }
              // the reference works on with its Old water table overwritten (3261-3262, 3282-3283)
              dm3 = c.nwt0;
              c.nwt0 = c.nwt1;
              dm1 = c.ths * (c.nwt1 + c.psib - c.nf1);
              dm2 = c.z1z2_moist(c.nf1, c.nwt1 + c.psib, c.nwt1);
              if ((dm1 - dm2 - mdelt) > 1.0E-6) c.nf1 = c.solve_front(mdelt, c.nwt1, c.nf0, 1);
              else c.nf1 = c.nwt1 + c.psib + 1.0E-6;
              c.nt1 = c.nt0;
              if (c.nf1 >= (c.nwt1 + c.psib)) {
                if (c.nt0 > 0.0) {
                  mdelt = -dtgw * (gw * 1.0E-6) * c.cosv / area;
                  c.nwt0 = dm3;
                  c.nwt1 = c.solve_wt((c.ths * c.nwt0 - (c.mu0 + mdelt)), (c.nt0 - c.psib));
                  c.mu1 = c.mi1 = c.total_moist(c.nwt1);
                  c.ri1 = c.ru1 = 0.0;
                  c.nf1 = c.nt1 = c.nwt1;
                } else if (c.nt1 == 0.0) {
                  c.nwt1 = c.nf1 = c.nt1 = 0.0;
                  c.mu1 = c.mi1 = 0.0;
                  c.ri1 = c.ru1 = 0.0;
                }
              } else {
                c.mu1 = c.mi1 + (c.mu0 - c.mi0);
                c.ri1 = 0.0;
                if (c.nt1 == 0.0) c.ru1 = c.ksat_mean(c.nf1);
                else c.ru1 = c.ru0;
              }
            }
          } else {
            mdelt = -dtgw * (gw * 1.0E-6) * c.cosv / area;
            c.nwt1 = c.solve_wt((c.ths * c.nwt0 - (c.mu0 + mdelt)), c.nwt0);
            c.mu1 = c.mi1 = c.total_moist(c.nwt1);
            c.ri1 = c.ru1 = 0.0;
            c.nf1 = c.nt1 = c.nwt1;
          }
        }
      } else if (((c.nwt1 + c.psib) <= c.nf0) && (c.nf0 != c.nwt0)) {
        c.mi1 = c.total_moist(c.nwt1);
        if (c.nt0 == c.nf0)
          c.nwt1 = c.solve_wt((c.ths * c.nwt0 - (c.mu0 + c.ths * (c.nwt0 - c.nwt1))), c.nwt1);
        else if (c.nt0 < c.nf0)
          c.nwt1 = c.solve_wt((c.ths * c.nwt0 - (c.mu0 + c.ths * (c.nwt0 - c.nwt1))), (c.nt0 - c.psib));
        c.mu1 = c.mi1 = c.total_moist(c.nwt1);
        c.ru1 = c.ri1 = 0.0;
        c.nt1 = c.nwt1;
        c.nf1 = c.nwt1;
      }
      break;
  }
  return st;
}

// ------------------------------------------------------------------ running means packed in AvSoilMoisture
// integer part = 1e4 * mean surface saturation, fraction = 0.1 * mean root saturation
// (tHydroModel.cpp:2045-2057, 3489-3501)
TRIBS_HD double av_soil_moist_update(double av, double surf_sc, double root_sc, double elapsed) {
  double bb = floor(av) * 1.0E-4;
  double md = (av - floor(av)) * 1.0E+1;
  double out = floor((bb * elapsed + surf_sc) / (elapsed + 1) * 1.0E+4);
  out += (md * elapsed + root_sc) / (elapsed + 1.0) * 1.0E-1;
  return out;
}

}  // namespace tribs
