/* TEST INFRASTRUCTURE — NOT PRODUCT CODE. See tribs_oracle_et.h.
 *
 * The reference keeps the working set of one cell in tEvapoTrans / tIntercept members; here it
 * is the struct `EtCell`, rebuilt for every cell from the arrays. Operation order follows the
 * reference expression by expression (compile with -ffp-contract=off) so results are bit-equal.
 */
#include "tribs_oracle_et.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define F(name) etf[ORCET_##name][i]
#define SF(name) snf[ORCSN_##name][i]

typedef struct {
  /* met record of the cell (tEvapoTrans::newHydroMetData, tEvapoTrans.cpp:3962-4010) */
  double airTemp, dewTemp, rHumidity, vPress, atmPress, windSpeed, skyCover, inShortR;
  /* derived ("C") copies the met functions leave behind */
  double dewTempC, vPressC, rHumidityC, atmPressC, skyCoverC, windSpeedC;
  int dewHumFlag, skyFlag, tsOption;
  /* land / soil coefficients (setCoeffs, tEvapoTrans.cpp:915-952) */
  double coeffAl, coeffH, coeffKt, coeffRs, coeffV, coeffSE, coeffST, coeffKs, coeffCs;
  double betaS, betaT;
  /* geometry */
  double elevation, slope, aspect, shelterFactor, horAngle0000;
  /* sun (global) */
  double alphaD, sinAlpha, sunaz, Io;
  int hour;
  /* radiation */
  double Ic, Is, Id, Ids, Ics, SunHour, inLongR, outLongR;
  /* energy balance */
  double Tso, Tlo, Gso, Rah, Rstm, hFlux, lFlux, gFlux, netRadC, potEvap, actEvap;
  double rain;
  /* options */
  int etOpt, gOpt, shOpt;
  double timeStep, Tlinke;
  double *const *snf;   /* snow fields of the nodes, or NULL without OPTSNOW */
  double coeffLAI;
} EtCell;

/* ---- tEvapoTrans.cpp:1224-1413 */
static double latent_heat(const EtCell *c) { return ((597.3 - c->airTemp * 0.57) * (4186.8)); }
static double sat_vp(const EtCell *c) { return (6.112 * exp((17.67 * c->airTemp) / (c->airTemp + 243.5))); }
static double sat_vp_at(double t) { return (6.112 * exp((17.67 * t) / (t + 243.5))); }

static double vapor_press(EtCell *c) {
  double dewTempK, rv = 461.5, eo = 6.112, to = 273.15;
  if (c->dewHumFlag == 0) {
    c->vPressC = sat_vp(c) * c->rHumidity / 100.0;
    dewTempK = 1.0 / ((1.0 / to) - (log(c->vPressC / eo) * rv / latent_heat(c)));
    c->dewTempC = dewTempK - 273.15;
    c->rHumidityC = c->rHumidity;
    c->dewTemp = c->dewTempC;
    c->vPress = c->vPressC;
  } else if (c->dewHumFlag == 1) {
    dewTempK = c->dewTemp + 273.15;
    c->dewTempC = c->dewTemp;
    c->vPressC = eo * exp((latent_heat(c) / rv) * ((1.0 / to) - (1.0 / dewTempK)));
    c->rHumidityC = 100.0 * c->vPressC / sat_vp(c);
    c->rHumidity = c->rHumidityC;
    c->vPress = c->vPressC;
  } else if (c->dewHumFlag == 2) {
    c->vPressC = c->vPress;
    c->rHumidityC = 100.0 * c->vPressC / sat_vp(c);
    dewTempK = 1.0 / ((1.0 / to) - (log(c->vPressC / eo) * rv / latent_heat(c)));
    c->dewTempC = dewTempK - 273.15;
    c->rHumidity = c->rHumidityC;
    c->dewTemp = c->dewTempC;
  }
  return c->vPressC;
}

static double total_press(EtCell *c) {
  double atmospress;
  if (fabs(c->atmPress - 9999.99) < 1.0E-3) atmospress = 1000.0;
  else atmospress = c->atmPress;
  c->atmPressC = atmospress;
  return atmospress + vapor_press(c);
}

static double claus_clap(EtCell *c) {
  double latHeat = latent_heat(c), esat = sat_vp(c), rv = 461.5, airTempK = c->airTemp + 273.15;
  return 100.0 * (latHeat * esat) / (rv * pow(airTempK, 2.0));
}

static double psychometric(EtCell *c) {
  double eps = 0.622, cp = 1013.0;
  return ((cp * total_press(c)) / (eps * latent_heat(c)));
}

static double density_moist(EtCell *c) {
  double Rconst = 287.6, airTempK = c->airTemp + 273.15;
  return (100.0 * total_press(c) / (Rconst * airTempK)) * (1.0 - 0.378 * vapor_press(c) / total_press(c));
}

/* ---- tEvapoTrans.cpp:1442-1512 */
static double aero_resist(EtCell *c) {
  double vonKarm = 0.41, vegHeight, vegBare = 1.0, vegFrac = c->coeffV;
  double zm, zom, zov, d, rav, ras;
  vegHeight = (c->coeffH == 0) ? 0.1 : c->coeffH;
  if (c->windSpeed == 0.0 || fabs(c->windSpeed - 9999.99) < 1.0E-3) c->windSpeedC = 0.1;
  else c->windSpeedC = c->windSpeed;
  zm = 2.0 + vegHeight; zom = 0.123 * vegHeight; zov = 0.0123 * vegHeight; d = 0.67 * vegHeight;
  rav = log((zm - d) / zom) * log((zm - d) / zov) / (c->windSpeedC * pow(vonKarm, 2.0));
  zm = 2.0 + vegBare; zom = 0.123 * vegBare; zov = 0.0123 * vegBare; d = 0.67 * vegBare;
  ras = log((zm - d) / zom) * log((zm - d) / zov) / (c->windSpeedC * pow(vonKarm, 2.0));
  return (1 - vegFrac) * ras + vegFrac * rav;
}

static double stom_resist(const EtCell *c) {
  static const double rsRatio[24] = {3.837, 3.589, 3.21, 2.43, 1.617, 1.196, 1.067, 1.014,
                                     0.995, 0.976, 0.976, 1.0, 1.053, 1.167, 1.354, 1.637,
                                     2.043, 2.66, 3.215, 3.507, 3.689, 3.818, 3.923, 4.024};
  double rs = c->coeffRs * rsRatio[c->hour];
  if (c->alphaD < 0.0) rs *= 1000.0;
  return rs;
}

/* ---- tEvapoTrans.cpp:2048-2068 */
static double comp_sky_cover(const EtCell *c) {
  double sc;
  if (c->rain > 0) sc = 10.0;
  else sc = round(10.0 * (3.2 * c->rHumidityC / 100.0 - 2.4) / 0.8);
  if (sc < 0) sc = 0;
  else if (sc > 10) sc = 10.0;
  return sc;
}

/* ---- tEvapoTrans.cpp:2004-2037 */
static double in_long_wave(EtCell *c) {
  double sigma = 5.67E-8, kCloud, airTempK, Ea, scover, N, v0;
  v0 = (c->shOpt < 3) ? c->shelterFactor : 1;
  if (c->skyFlag == 1) { c->skyCover = comp_sky_cover(c); scover = c->skyCover; }
  else scover = c->skyCover;
  c->skyCoverC = scover;
  N = scover / 10.0;
  kCloud = 1.0 + 0.17 * pow(N, 2.0);
  airTempK = c->airTemp + 273.15;
  Ea = 0.74 + 0.0049 * vapor_press(c);
  return v0 * kCloud * Ea * sigma * pow(airTempK, 4.0);
}

/* ---- tEvapoTrans.cpp:1939-1993 */
static void direct_diffuse(EtCell *c, double elev) {
  double h0, m, pp0, Dh0ref, h0ref, drm, TnTLK, Fdh0, A1p, A1, A2, A3;
  double pi = 4.0 * atan(1.0), T = c->Tlinke;
  h0 = c->alphaD;
  pp0 = exp(-elev / 8434.5);
  Dh0ref = 0.061359 * (0.1594 + 1.123 * h0 + 0.065656 * h0 * h0) / (1 + 28.9344 * h0 + 277.3971 * h0 * h0);
  h0ref = h0 + Dh0ref;
  m = pp0 / (sin(h0ref * pi / 180.) + 0.50572 * pow((h0ref + 6.07995), -1.6364));
  if (m <= 20.0) drm = 1 / (6.6296 + 1.7513 * m - 0.1202 * m * m + 0.0065 * pow(m, 3.0) - 0.00013 * pow(m, 4.0));
  else drm = 1 / (10.4 + 0.718 * m);
  c->Ic = c->Io * exp(-0.8662 * T * m * drm);
  TnTLK = -0.015843 + 0.030543 * T + 0.0003797 * pow(T, 2.0);
  A1p = 0.26463 - 0.061581 * T + 0.0031408 * pow(T, 2.0);
  if (A1p * TnTLK < 0.0022) A1 = 0.0022 / TnTLK;
  else A1 = A1p;
  A2 = 2.04020 + 0.018945 * T - 0.011161 * pow(T, 2.0);
  A3 = -1.3025 + 0.039231 * T + 0.0085079 * pow(T, 2.0);
  Fdh0 = A1 + A2 * c->sinAlpha + A3 * pow(c->sinAlpha, 2.0);
  c->Id = c->Io * TnTLK * Fdh0;
}

/* ---- tEvapoTrans::aboveHorizon, tEvapoTrans.cpp:2080-2270. The sector test `(x < 12.25) || (x >= -12.25)` is
 * always true, so the first of the 16 cases (azimuth 0) runs, sets `dummy`, and the loop ends: the
 * horizon angle at azimuth 0 alone decides whether the cell sees the sun. */
static double above_horizon(const EtCell *c) { return (c->horAngle0000 * (180 / 3.1416) < c->alphaD) ? 1.0 : 0.0; }

/* ---- tEvapoTrans.cpp:1751-1930 (observed stations) */
static void in_short_wave(EtCell *c, double *const *etf, int i) {
  double N = 0, Iv = 0, Isw = 0, Ir = 0, v, cosi, vegAbs = 0.0, soilAbs = 0.0;
  c->Ic = c->Is = c->Id = c->Ids = c->Ics = 0.0;
  F(ShortRadIn_dir) = 0.0; F(ShortRadIn_dif) = 0.0; F(ShortRadSlope) = 0.0;
  F(ShortAbsbVeg) = 0.0; F(ShortAbsbSoi) = 0.0;
  c->SunHour = 0.0;
  if (c->alphaD > 0.0) {
    direct_diffuse(c, c->elevation);
    N = 0;
    if (c->tsOption > 1) {
      double RadGlobClr = (c->inShortR / (1.0 - 0.65 * pow(N, 2.0)));
      c->Ic = c->Ic / (c->Ic * c->sinAlpha + c->Id) * RadGlobClr;
      c->Id = RadGlobClr - c->Ic * c->sinAlpha;
    }
    if (c->shOpt < 4) {
      cosi = cos(c->slope) * c->sinAlpha + sin(c->slope) * cos(asin(c->sinAlpha)) * cos(c->sunaz - c->aspect);
      if (cosi >= 0) { c->Ics = c->Ic * cosi; c->SunHour = 1.0; }
      else { c->Ics = 0.0; c->SunHour = 0.0; }
    } else { c->Ics = c->Ic; c->SunHour = 1.0; }
    if ((c->shOpt == 2) || (c->shOpt == 1)) { c->Ics *= above_horizon(c); c->SunHour *= above_horizon(c); }
    if ((c->shOpt > 0) && (c->shOpt < 3)) v = c->shelterFactor;
    else if ((c->shOpt == 0) || (c->shOpt == 3)) v = 0.5 * (1.0 + cos(c->slope));
    else v = 1.0;
    c->Ids = c->Id * v;
    c->Is = (1.0 - 0.65 * pow(N, 2.0)) * (c->Ics + c->Ids);
    if (c->shOpt == 0) { Ir = c->coeffAl * c->Is * (1 - cos(c->slope)) * 0.5; c->Is += Ir; }
    else if ((c->shOpt > 1) && (c->shOpt < 4)) { Ir = c->coeffAl * c->Is * (0.5 * (1 + cos(c->slope)) - c->shelterFactor); c->Is += Ir; }
    else Ir = 0.0;
    if (c->etOpt == 1) {
      F(ShortRadSlope) = c->Is;
      soilAbs = (c->Is * c->coeffKt * c->coeffV + c->Is * (1.0 - c->coeffV)) * (1.0 - c->coeffAl);
      vegAbs = c->Is * c->coeffV * (1.0 - c->coeffKt) * (1.0 - c->coeffAl);
      F(ShortAbsbSoi) = soilAbs; F(ShortAbsbVeg) = vegAbs;
      Iv = soilAbs;
    } else Iv = c->Is * (1.0 - c->coeffAl);
    Isw = Iv;
  } else {
    c->Ic = c->Is = N = Iv = Isw = c->Id = c->Ids = c->Ics = Ir = 0.0;
  }
  if (c->tsOption > 1) F(ShortRadIn) = c->inShortR;
  else F(ShortRadIn) = Isw;
  F(ShortRadIn_dir) = c->Ics * (1.0 - 0.65 * pow(N, 2.0));
  F(ShortRadIn_dif) = (c->Ids + Ir) * (1.0 - 0.65 * pow(N, 2.0));
}

/* ---- tEvapoTrans.cpp:2663-2716 */
static double force_restore(EtCell *c, double Ts, int option) {
  double w1, d1, k, cs, ddel, pi, dt, alpha = 0, nd, G, Tg, Tl, dg, dTgdt, dGdTg, dTK, tempo = 0;
  dTK = 0.001;
  pi = 4.0 * atan(1.0);
  w1 = 2. * pi / 86400.;
  dt = c->timeStep * 60.0;
  cs = c->coeffCs;
  k = c->coeffKs / c->coeffCs;
  d1 = pow((2.0 * k / w1), 0.5);
  ddel = 0.1;
  nd = ddel / d1;
  if (nd >= 0.0 && nd <= 5.0)
    alpha = 1 + 0.943 * nd + 0.223 * pow(nd, 2.0) + 0.0168 * pow(nd, 3.0) - 0.00527 * pow(nd, 4.0);
  Tg = Ts; Tl = c->Tlo;
  dg = (0.5 * cs * d1) / (1.0 + (w1 * dt) / (2.0 * pow(365., 0.5)));
  if (option == 1) {
    dTgdt = (Tg - c->Tso) / dt;
    G = dg * (alpha * dTgdt + w1 * (Tg - Tl));
    c->Gso = G;
    tempo = G;
  } else if (option == 2) {
    dTgdt = (Tg + dTK - c->Tso) / dt;
    dGdTg = (dg * (alpha * dTgdt + w1 * (Tg - Tl)) - c->Gso) / dTK;
    tempo = dGdTg;
  }
  return tempo;
}

/* ---- tEvapoTrans.cpp:2486-2632 */
static void func_and_deriv(EtCell *c, double *const *etf, int i, double Tg, double *fv, double *dv, double Rabsb_soi) {
  double Rn, Lsoi, Hsoi, lEsoi, G = 0, dQ, num, Ep = 0, Eps, LE = 0, H = 0, ccTs = 0;
  double alpha = 1.26, sigma = 5.67E-8, Es = 0.98, Cp = 1013.0, Ch = 0.0025, rv = 461.5;
  double rho = density_moist(c);
  double P = total_press(c);
  double satVap = sat_vp(c);
  double es = vapor_press(c);
  double lam = latent_heat(c);
  double cc = claus_clap(c);
  double psy = 100.0 * psychometric(c);
  double denom = ((cc / psy) + 1.0);
  double denomrs = ((cc / psy) + 1.0 + c->Rstm / c->Rah);
  double pi = 4.0 * atan(1.0);
  double DTime = c->timeStep * 60.0;
  double dRndTg = 0, dLdTg, dHdTg = 0, dlEdTg = 0, dGdTg = 0, soiFct = 0, vegFct = 0, esat, qhSatTs, qhTa, v1;
  v1 = (c->shOpt < 3) ? c->shelterFactor : 1;
  F(SurfTemp) = Tg - 273.15;
  c->inLongR = F(LongRadIn);
  c->outLongR = v1 * Es * sigma * pow(Tg, 4.0);
  Lsoi = c->outLongR - c->inLongR;
  if (c->gOpt == 1) G = pow((4.0 * c->coeffKs * c->coeffCs / (pi * DTime)), 0.5) * (Tg - c->Tso);
  else if (c->gOpt == 2) G = force_restore(c, Tg, 1);
  Rn = Rabsb_soi - Lsoi;
  if (c->etOpt == 1 || c->etOpt == 3) H = rho * Cp * (Tg - (c->airTemp + 273.15)) / c->Rah;
  else if (c->etOpt == 2) H = rho * Cp * (Tg - (c->airTemp + 273.15)) * Ch * c->windSpeedC;
  Hsoi = H;
  if (c->etOpt == 1) {
    dQ = (0.622 / P) * (satVap - es) * rho * lam / c->Rah;
    soiFct = (1.0 - c->coeffV) * c->betaS;
    vegFct = c->coeffV * c->betaT;
    num = (cc / psy) * (Rn - G) + dQ;
    Eps = num / denom / lam;
    Ep = num / denomrs / lam;
    LE = soiFct * lam * Eps + vegFct * lam * Ep;
    Ep *= (denomrs / denom);
  } else if (c->etOpt == 2) {
    esat = sat_vp_at(Tg - 273.15);
    ccTs = (lam * esat) / (rv * pow(Tg, 2.0));
    qhSatTs = (0.622 / P) * esat;
    qhTa = (0.622 / P) * es;
    if (qhSatTs > qhTa) Ep = rho * Ch * (qhSatTs - qhTa) * c->windSpeedC; else Ep = 0.0;
    LE = Ep * lam * c->betaS;
  } else if (c->etOpt == 3) {
    if (Rn > G) Ep = (alpha / lam) * (Rn - G) * ((cc / psy) / denom); else Ep = 0.0;
    LE = Ep * lam * c->betaS;
  }
  lEsoi = LE;
  *fv = -Rabsb_soi + Lsoi + Hsoi + lEsoi + G;
  dLdTg = 4.0 * Es * sigma * pow(Tg, 3.0);
  dRndTg = -dLdTg;
  if (c->gOpt == 1) dGdTg = pow((4.0 * c->coeffKs * c->coeffCs / (pi * DTime)), 0.5);
  else if (c->gOpt == 2) dGdTg = force_restore(c, Tg, 2);
  if (c->etOpt == 1) {
    dHdTg = rho * Cp / c->Rah;
    dlEdTg = soiFct * (cc / psy) * (dRndTg - dGdTg) / denom + vegFct * (cc / psy) * (dRndTg - dGdTg) / denomrs;
  } else if (c->etOpt == 2) {
    dHdTg = rho * Cp * Ch * c->windSpeedC;
    dlEdTg = lam * rho * Ch * c->windSpeedC * (0.622 / P) * ccTs * c->betaS;
  } else if (c->etOpt == 3) {
    dHdTg = rho * Cp / c->Rah;
    dlEdTg = c->betaS * (alpha) * ((cc / psy) / denom) * (dRndTg - dGdTg);
  }
  *dv = dLdTg + dHdTg + dlEdTg + dGdTg;
  c->hFlux = Hsoi; c->lFlux = lEsoi; c->netRadC = Rn; c->gFlux = G; c->potEvap = Ep;
}

/* ---- tEvapoTrans.cpp:2391-2474 */
static double rtsafe_energy(EtCell *c, double *const *etf, int i, double x1, double x2, double xacc,
                            double xguess, double Rabsb) {
  int j;
  double df = 0, dx = 0, dxold, f, fh = 0, fl = 0, temp, xh, xl, rts;
  func_and_deriv(c, etf, i, x1, &fl, &df, Rabsb);
  if (fl == 0.0) return x1;
  func_and_deriv(c, etf, i, x2, &fh, &df, Rabsb);
  if (fh == 0.0) return x2;
  if (fl < 0.0) { xl = x1; xh = x2; } else { xh = x1; xl = x2; }
  rts = xguess;
  dxold = fabs(x2 - x1);
  dx = dxold;
  func_and_deriv(c, etf, i, rts, &f, &df, Rabsb);
  for (j = 1; j <= 100; j++) {
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
    func_and_deriv(c, etf, i, rts, &f, &df, Rabsb);
    if (f < 0.0) xl = rts; else xh = rts;
  }
  return xguess;
}

/* ---- tEvapoTrans.cpp:2281-2357 */
static double energy_balance(EtCell *c, double *const *etf, int i) {
  double f, df, Tg, soilAbs;
  c->SunHour = 0;
  c->inLongR = in_long_wave(c);
  F(LongRadIn) = c->inLongR;
  in_short_wave(c, etf, i);
  soilAbs = F(ShortAbsbSoi);
  c->Rah = aero_resist(c);
  c->Rstm = stom_resist(c);
  Tg = c->Tso;
  Tg = rtsafe_energy(c, etf, i, 223.15, 373.15, 1.0e-5, Tg, soilAbs);
  func_and_deriv(c, etf, i, Tg, &f, &df, soilAbs);
  c->Tso = Tg;
  c->Tlo += (c->gFlux * c->timeStep * 60.0) / (c->coeffCs * 33.862683 * pow((c->coeffKs / c->coeffCs / 3.6361E-05), 0.5));
  F(SurfTemp) = Tg - 273.15;
  F(SoilTemp) = c->Tlo - 273.15;
  F(NetRad) = c->netRadC; F(GFlux) = c->gFlux; F(HFlux) = c->hFlux; F(LFlux) = c->lFlux;
  F(LongRadOut) = c->outLongR;
  if (c->snf) {   /* energyBalance clears the snow fluxes of the node, tEvapoTrans.cpp:2343-2351 */
    double *const *snf = c->snf;
    SF(SnLHF) = 0.0; SF(SnSHF) = 0.0; SF(SnSub) = 0.0; SF(SnEvap) = 0.0; SF(SnPHF) = 0.0; SF(SnGHF) = 0.0;
    SF(SnRLin) = 0.0; SF(SnRLout) = 0.0; SF(SnRSin) = 0.0;
  }
  return c->potEvap;
}

static void load_coeffs(const orc_et_basin *b, EtCell *c, int i) {
  const double *lp = b->land_table + (size_t)b->land_class[i] * 15;
  c->etOpt = b->et_option; c->gOpt = b->gflux_option; c->shOpt = b->shelter_option;
  c->timeStep = b->metstep_min; c->Tlinke = b->tlinke;
  c->coeffAl = c->coeffH = c->coeffKt = c->coeffRs = 0.0;
  c->coeffV = 0.0;
  if (b->et_option == 1) { c->coeffAl = lp[7]; c->coeffH = lp[8]; c->coeffKt = lp[9]; c->coeffRs = lp[10]; c->coeffV = lp[11]; }
  else if (b->et_option == 2 || b->et_option == 3) { c->coeffAl = lp[7]; c->coeffV = lp[11]; }
  else if (b->et_option == 4) c->coeffV = lp[11];
  c->coeffSE = lp[13]; c->coeffST = lp[14];
  if (c->coeffV >= 1.0) c->coeffV = 0.99;
  c->coeffKs = b->VolHeatCond[i]; c->coeffCs = b->SoilHeatCap[i];
}

static void load_met(const orc_et_basin *b, const orc_et_step *s, const orc_met *m, EtCell *c, int i) {
  int r = m->cell_record[i];
  c->airTemp = m->air_temp[r];
  c->airTemp += b->temp_lapse * (c->elevation - m->elev[r]);
  c->dewTemp = m->dew_temp[r]; c->rHumidity = m->rel_humid[r]; c->vPress = m->vap_press[r];
  c->atmPress = m->atm_press[r]; c->windSpeed = m->wind_speed[r]; c->skyCover = m->sky_cover[r];
  c->inShortR = m->rad_global[r];
  c->dewHumFlag = s->dew_hum_flag; c->tsOption = s->ts_option;
  c->alphaD = s->alphaD; c->sinAlpha = s->sinAlpha; c->sunaz = s->sunaz; c->Io = s->Io; c->hour = s->hour;
}

/* tEvapoTrans::betaFunc / betaFuncT, tEvapoTrans.cpp:2911-3010 */
static void beta_funcs(const orc_et_basin *b, const orc_et_shared *sh, EtCell *c, int i) {
  double Thr = b->ThetaR[i], Th, ratio, beta;
  Th = sh->SoilMoisture[i];
  ratio = (Th - Thr) / (c->coeffSE - Thr);
  if (ratio > 1.0) beta = 1.0; else if (ratio < 0.0) beta = 0.0; else beta = ratio;
  if (Th <= b->soil_cutoff[i]) beta = 0.0;
  c->betaS = beta;
  Th = sh->RootMoisture[i];
  ratio = (Th - Thr) / (c->coeffST - Thr);
  if (ratio < 1.0) beta = ratio; else beta = 1.0;
  if (Th <= b->root_cutoff[i]) beta = 0.0;
  c->betaT = beta;
}

/* met assembly of one cell: the part of callEvapoPotential / callSnowPack before the evaporation method */
static void met_assembly(const orc_et_basin *b, const orc_et_step *s, const orc_met *m,
                         double *const *etf, const orc_et_shared *sh, int i, EtCell *c) {
  double vPress;
  c->snf = NULL;
  c->coeffLAI = b->land_table[(size_t)b->land_class[i] * 15 + 12];
  c->rain = sh->Rain[i];
  c->elevation = b->z[i];
  c->slope = fabs(atan(b->flow_slope[i]));
  c->aspect = b->aspect[i];
  c->horAngle0000 = 0.0;
  if ((b->shelter_option > 0) && (b->shelter_option < 4)) {   /* tShelter's sky-view factor lives in the node */
    c->shelterFactor = F(SheltFact);
    c->horAngle0000 = b->hor_angle_0000[i];
  } else if (b->shelter_option == 0) { c->shelterFactor = 0.5 * (1 + cos(c->slope)); F(SheltFact) = c->shelterFactor; }
  else c->shelterFactor = 1;
  load_coeffs(b, c, i);
  load_met(b, s, m, c, i);
  if (s->first_call) { c->Tso = m->air_temp[m->cell_record[i]] + 273.15; c->Tlo = c->Tso; }
  else { c->Tso = 0.0; c->Tlo = 0.0; }
  c->skyFlag = (i >= s->sky_flag_from) || (fabs(c->skyCover - 9999.99) < 1.0E-3);
  vPress = vapor_press(c);
  F(AirTemp) = c->airTemp; F(DewTemp) = c->dewTemp; F(RelHumid) = c->rHumidity; F(VapPressure) = vPress;
  if (c->skyFlag == 1) c->skyCover = comp_sky_cover(c);
  F(SkyCover) = c->skyCover; F(WindSpeed) = c->windSpeed; F(AirPressure) = c->atmPress;
  F(ShortRadIn) = c->inShortR;
  if (s->first_call) { F(SoilTemp) = c->Tlo - 273.15; F(SurfTemp) = c->Tso - 273.15; }
  if (b->i_option == 0) sh->NetPrecipitation[i] = c->rain;
  beta_funcs(b, sh, c, i);
  c->Tso = F(SurfTemp) + 273.15;
  c->Tlo = F(SoilTemp) + 273.15;
  c->Gso = 0.0;
}

/* EvapPenmanMonteith / Deardorff / PriestlyTaylor (tEvapoTrans.cpp:2790-2870) */
static void evap_method(const orc_et_basin *b, double *const *etf, int i, EtCell *c) {
  c->potEvap = 3600.0 * energy_balance(c, etf, i);
  if (b->et_option == 1) c->actEvap = 3600.0 * (c->lFlux / (latent_heat(c)));
  else c->actEvap = c->betaS * c->potEvap;
}

/* setToNode, tEvapoTrans.cpp:3039-3076 */
static void set_to_node(const orc_et_step *s, double *const *etf, int i, EtCell *c) {
  double tmp = 0;
  if (c->dewHumFlag == 0) { F(DewTemp) = c->dewTempC; F(VapPressure) = c->vPressC; }
  else if (c->dewHumFlag == 1) { F(VapPressure) = c->vPressC; F(RelHumid) = c->rHumidityC; }
  else if (c->dewHumFlag == 2) { F(DewTemp) = c->dewTempC; F(RelHumid) = c->rHumidityC; }
  F(PotEvap) = c->potEvap; F(ActEvap) = c->actEvap; F(AirTemp) = c->airTemp;
  F(AirPressure) = c->atmPressC; F(SkyCover) = c->skyCoverC; F(WindSpeed) = c->windSpeedC;
  F(CumHrsSun) += c->SunHour;
  if ((fabs(c->hFlux) + fabs(c->lFlux)) > 1.0) tmp = fabs(c->lFlux) / (fabs(c->hFlux) + fabs(c->lFlux));
  if (fabs(s->te_met - 1.0) <= 1.0E-6) F(AvEvapFract) = tmp;
  else if (s->te_met > 1.0) F(AvEvapFract) = (F(AvEvapFract) * (s->te_met - 1.0) + tmp) / s->te_met;
  F(SheltFact) = c->shelterFactor;
  F(CumRSin) += c->inShortR * 3600.0;
}

int orc_et_potential(const orc_et_basin *b, const orc_et_step *s, const orc_met *m,
                     double *const *etf, const orc_et_shared *sh) {
  if (b->shelter_option >= 1 && b->shelter_option <= 3 && !b->hor_angle_0000) return 1;
  if (b->et_option < 1 || b->et_option > 3) return 1;
  for (int i = 0; i < b->n; i++) {
    EtCell cell, *c = &cell;
    met_assembly(b, s, m, etf, sh, i, c);
    evap_method(b, etf, i, c);
    set_to_node(s, etf, i, c);
  }
  return 0;
}

/* ---- tIntercept.cpp:214-253 */
static void intercept_gray(const orc_et_basin *b, double *const *etf, const orc_et_shared *sh, int i) {
  const double *lp = b->land_table + (size_t)b->land_class[i] * 15;
  double coeffA = lp[1], coeffB = lp[2], coeffV = lp[11];
  double cumIntercept, intercept, rainfall, minRainAmount = 1.0;
  rainfall = sh->Rain[i];
  if (rainfall <= minRainAmount) F(StormLength) += b->etistep_h; else F(StormLength) = 0.0;
  cumIntercept = F(CumIntercept);
  if (cumIntercept <= coeffA) intercept = rainfall; else intercept = rainfall * coeffB;
  F(InterceptLoss) = intercept;
  if (F(StormLength) < b->max_interstorm) F(CumIntercept) = cumIntercept + intercept * b->etistep_h;
  else F(CumIntercept) = 0.0;
  sh->NetPrecipitation[i] = coeffV * (rainfall - intercept) + (1 - coeffV) * rainfall;
}

/* ---- tIntercept.cpp:283-459 */
static double rutter_fn(double C, double R, double Ep, double P, double S, double K, double bb) {
  if (C < S) return ((1 - P) * R - (C / S) * Ep - K * exp(bb * (C - S)));
  return ((1 - P) * R - Ep - K * exp(bb * (C - S)));
}

static void intercept_rutter(const orc_et_basin *b, double *const *etf, const orc_et_shared *sh, int i, double evaporation) {
  const double *lp = b->land_table + (size_t)b->land_class[i] * 15;
  double coeffP = lp[3], coeffS = lp[4], coeffK = lp[5], coeffb = lp[6], coeffV = lp[11];
  double rainfall, prevStorage, currentStorage, throughfall, netPrecipitation, interception, cumIntercept, ctos;
  double minRainAmount = 1.0, dt = b->etistep_h;
  rainfall = sh->Rain[i];
  prevStorage = F(CanStorage);
  ctos = F(CanStorage) / coeffS;
  if (ctos > 1) ctos = 0;
  if (rainfall <= minRainAmount) F(StormLength) += dt; else F(StormLength) = 0.0;
  if (prevStorage > evaporation * ctos * dt || rainfall > 0) {
    cumIntercept = F(CumIntercept);
    {   /* storageRungeKutta */
      double numiter = floor(rainfall * 3.0), h, t0, tlast = dt, c0 = prevStorage, ta, tb, tc, td;
      if (numiter < 50.0) numiter = 50.0;
      h = dt / numiter;
      for (t0 = 0.0; t0 < tlast; t0 += h) {
        ta = h * rutter_fn(c0, rainfall, evaporation, coeffP, coeffS, coeffK, coeffb);
        tb = h * rutter_fn(c0 + ta / 2.0, rainfall, evaporation, coeffP, coeffS, coeffK, coeffb);
        tc = h * rutter_fn(c0 + tb / 2.0, rainfall, evaporation, coeffP, coeffS, coeffK, coeffb);
        td = h * rutter_fn(c0 + tc, rainfall, evaporation, coeffP, coeffS, coeffK, coeffb);
        c0 += (ta / 6.0 + tb / 3.0 + tc / 3.0 + td / 6.0);
      }
      currentStorage = c0;
    }
    double rain_on_canopy_vol = (1 - coeffP) * rainfall * dt;
    double total_available_water = prevStorage + rain_on_canopy_vol;
    if (currentStorage > total_available_water) currentStorage = total_available_water;
    double avg_storage = (prevStorage + currentStorage) / 2.0;
    double evapWetCanopy_rate;
    if (avg_storage < coeffS) evapWetCanopy_rate = (avg_storage / coeffS) * evaporation;
    else evapWetCanopy_rate = evaporation;
    if (evapWetCanopy_rate < 0) evapWetCanopy_rate = 0.0;
    double storage_change_vol = currentStorage - prevStorage;
    double total_loss_vol = rain_on_canopy_vol - storage_change_vol;
    if (total_loss_vol < 0.0) total_loss_vol = 0.0;
    double evapWetCanopy_vol = evapWetCanopy_rate * dt;
    if (evapWetCanopy_vol > total_loss_vol) evapWetCanopy_vol = total_loss_vol;
    if (evapWetCanopy_vol < 0.0) evapWetCanopy_vol = 0.0;
    double drainage_vol = total_loss_vol - evapWetCanopy_vol;
    if (drainage_vol < 0.0) drainage_vol = 0.0;
    if (currentStorage * 100 / coeffS <= 3) {
      double potential_evap_vol = evaporation * dt;
      if ((evapWetCanopy_vol + currentStorage) <= potential_evap_vol) evapWetCanopy_vol += currentStorage;
      else drainage_vol += currentStorage;
      currentStorage = 0.0;
    }
    throughfall = coeffP * rainfall;
    netPrecipitation = (drainage_vol / dt) + throughfall;
    if (netPrecipitation <= rainfall) interception = rainfall - netPrecipitation; else interception = 0.0;
    netPrecipitation = coeffV * netPrecipitation + (1 - coeffV) * rainfall;
    if (currentStorage < 0.0) currentStorage = 0.0;
    sh->NetPrecipitation[i] = netPrecipitation;
    F(InterceptLoss) = interception;
    F(CanStorage) = currentStorage;
    F(EvapWetCanopy) = evapWetCanopy_vol / dt;
    if (F(StormLength) < b->max_interstorm) F(CumIntercept) = cumIntercept + interception * dt;
    else F(CumIntercept) = 0.0;
  } else {
    sh->NetPrecipitation[i] = 0.0;
    F(InterceptLoss) = 0.0; F(CanStorage) = 0.0; F(CumIntercept) = 0.0; F(EvapWetCanopy) = 0.0;
  }
}

static int there_is_canopy(const orc_et_basin *b, int i) {
  if (b->i_option == 1) return 1;
  if (b->i_option == 2) return b->land_table[(size_t)b->land_class[i] * 15 + 3] < 0.999;
  return 0;
}

/* ---- tEvapoTrans.cpp:980-1197 */
static void components_cell(const orc_et_basin *b, const orc_et_step *s, const orc_met *m,
                            double *const *etf, const orc_et_shared *sh, int flag, int i) {
  {
    EtCell cell, *c = &cell;
    double potEvaporation, actEvaporation, cc, psy, ra, rs, transFactor;
    double evapWetCanopy = 0.0, evapDryCanopy, evapSoil, evapoTranspiration;
    if (!b->et_option) {
      if (flag && there_is_canopy(b, i)) {
        if (b->i_option == 1) intercept_gray(b, etf, sh, i); else intercept_rutter(b, etf, sh, i, 0.0);
      }
      return;
    }
    c->elevation = b->z[i];
    load_coeffs(b, c, i);
    beta_funcs(b, sh, c, i);
    load_met(b, s, m, c, i);
    potEvaporation = F(PotEvap);
    actEvaporation = F(ActEvap);
    if (potEvaporation < 0) { potEvaporation = 0.0; F(PotEvap) = 0.0; }
    if (actEvaporation < 0.0) { actEvaporation = 0.0; F(ActEvap) = 0.0; }
    cc = claus_clap(c);
    psy = psychometric(c);
    ra = aero_resist(c);
    rs = stom_resist(c);
    transFactor = (cc + psy) / (cc + psy * (1 + rs / ra));
    if (flag && (c->coeffV > 0) && there_is_canopy(b, i)) {
      if (b->i_option == 1) {
        double CanStorage = F(CumIntercept);
        if (CanStorage >= potEvaporation * b->etistep_h) evapWetCanopy = potEvaporation;
        else evapWetCanopy = CanStorage / b->etistep_h;
        if (evapWetCanopy > potEvaporation) evapWetCanopy = potEvaporation;
        intercept_gray(b, etf, sh, i);
      } else if (b->i_option == 2) {
        intercept_rutter(b, etf, sh, i, potEvaporation);
        evapWetCanopy = F(EvapWetCanopy);
      }
    } else {
      sh->NetPrecipitation[i] = sh->Rain[i];
      evapWetCanopy = 0.0;
    }
    evapDryCanopy = c->betaT * ((potEvaporation - evapWetCanopy) * transFactor);
    double potEvaporationRemaining = potEvaporation - evapWetCanopy - evapDryCanopy;
    if (potEvaporationRemaining < 0.0) potEvaporationRemaining = 0.0;
    evapWetCanopy *= c->coeffV;
    evapDryCanopy *= c->coeffV;
    if (b->bedrock[i] <= 0) evapSoil = 0;
    else evapSoil = (1 - c->coeffV) * potEvaporationRemaining * c->betaS;
    evapoTranspiration = evapWetCanopy + evapDryCanopy + evapSoil;
    if (b->i_option == 1) {
      double CanStorage = F(CumIntercept);
      F(CumIntercept) = CanStorage - evapWetCanopy / c->coeffV * b->etistep_h;
    }
    F(EvapWetCanopy) = evapWetCanopy;
    sh->EvapDryCanopy[i] = evapDryCanopy;
    sh->EvapSoil[i] = evapSoil;
    F(EvapoTrans) = evapoTranspiration;
    F(CumTotEvap) += evapoTranspiration;
    F(CumBarEvap) += evapSoil;
    if (fabs(s->te_eti - 1.0) < 1.0E-6) F(AvET) = evapoTranspiration;
    else if (s->te_eti > 1.0) F(AvET) = (F(AvET) * (s->te_eti - 1.0) + evapoTranspiration) / s->te_eti;
  }
}

int orc_et_components(const orc_et_basin *b, const orc_et_step *s, const orc_met *m,
                      double *const *etf, const orc_et_shared *sh, int flag) {
  for (int i = 0; i < b->n; i++) components_cell(b, s, m, etf, sh, flag, i);
  return 0;
}

int orc_et_n_fields(void) { return ORC_N_ET_FIELDS; }

/* ==================================================================== snow pack (tSnowPack.cpp) */
typedef struct {   /* the tSnowPack members one cell works with */
  double liqWE, iceWE, snWE, snSub, snEvap, liqRoute, Utot, Usn, Uwat, Utotold, liqWatCont;
  double liqTempC, iceTempC, snTempC, crustAge, densityAge, ETAge;
  double H, L, G, Prec, Rn, dUint, RLin, RLout, RSin, Uerr, snPrec, liqPrec, snUnload, snCanWE;
  double Qcs, I, Iold, Imax, LAI, Lm, airTempK, phfOnOff, albedo, snDepthm, snOnOff;
  double peakSnWE, peakSnWEtemp, persMax, persMaxtemp, inittime, inittimeTemp, peaktime;
  double timeSteph, timeSteps, timeStepm, minSnTemp, snliqfrac;
  int hillAlbedoOption, hourly;
} Snow;

static const double rholiqkg = 1000.0, rhoicekg = 920.0, rhoAir = 1.3, cpicekJ = 2.1, cpwaterkJ = 4.190, cpairkJ = 1.006;
static const double latFreezekJ = 334.0, latVapkJ = 2470.0, latSubkJ = 334.0 + 2470.0;
static const double naughttokilo = 1e-3, kilotonaught = 1e3, naughttocm = 100.0, cmtonaught = 0.01, ctom = 10.0, mtoc = 0.1;
#define REF_PI 3.1415926

/* tSnowPack.cpp:1244-1290 */
static double snow_frac(const EtCell *c) {
  double Ta = c->airTemp, RH = c->rHumidity, Tw, f1;
  Tw = Ta * atan(0.151977 * pow(RH + 8.313659, 0.5)) + atan(Ta + RH) - atan(RH - 1.676331) +
       0.00391838 * pow(RH, 1.5) * atan(0.023101 * RH) - 4.686035;
  f1 = 1 + 6.99 * pow(10, -5) * exp(2 * (Tw + 3.97));
  if (Tw > 5) return 0;
  return 1 / f1;
}

/* tSnowPack.cpp:1373-1480 */
static double res_fact(const EtCell *c, const Snow *w) {
  const double vonKarm = 0.41, g = 9.81, Ri_cr = 0.1;
  double windSpeedC, windSpeedS, ra, vegHeight, vegFrac, vegBare, zm, zom, zov, d, rav, ras;
  windSpeedC = (c->windSpeed == 0.0 || fabs(c->windSpeed - 9999.99) < 1e-3) ? 0.01 : c->windSpeed;
  vegHeight = (c->coeffH <= 0) ? 0.1 : c->coeffH;
  vegHeight = (vegHeight > w->snDepthm) ? vegHeight - w->snDepthm : 0.1;
  vegBare = 0.1;
  vegFrac = c->coeffV;
  if (w->snDepthm < c->coeffH) windSpeedS = windSpeedC * exp(-0.5 * c->coeffLAI * (1.0 - (w->snDepthm / c->coeffH)));
  else windSpeedS = windSpeedC;
  zm = 2.0 + vegHeight; zom = 0.123 * vegHeight; zov = 0.0123 * vegHeight; d = 0.67 * vegHeight;
  rav = log((zm - d) / zom) * log((zm - d) / zov) / (windSpeedC * pow(vonKarm, 2));
  zm = 2.0 + vegBare; zom = 0.123 * vegBare; zov = 0.0123 * vegBare; d = 0.67 * vegBare;
  ras = log((zm - d) / zom) * log((zm - d) / zov) / (windSpeedS * pow(vonKarm, 2));
  ra = (1 - vegFrac) * ras + vegFrac * rav;
  double Ts = (w->liqWE > 1e-5) ? 273.15 : w->snTempC + 273.15;
  double Ta = c->airTemp + 273.15, T_avg = 0.5 * (Ta + Ts);
  double zm_eff = (0.1 > 2.0 - w->snDepthm) ? 0.1 : 2.0 - w->snDepthm;
  double z0 = 0.123 * vegHeight;
  double RiB = (g * zm_eff * (Ta - Ts)) / (T_avg * windSpeedS * windSpeedS);
  double Ri_u = 1.0 / (log(zm_eff / z0) + 5.0), C_stab;
  if (RiB < 0.0) C_stab = pow(1.0 - 16.0 * RiB, 0.5);
  else if (RiB <= Ri_u) C_stab = 1.0 / pow(1.0 - (RiB / Ri_cr), 2.0);
  else C_stab = 1.0 / pow(1.0 - (Ri_u / Ri_cr), 2.0);
  if (C_stab > 15.0) C_stab = 15.0;
  ra *= C_stab;
  return 1.0 / ra;
}

/* tSnowPack.cpp:1177-1236 */
static double latent_hf(const EtCell *c, const Snow *w, double Kaero) {
  double t = (w->liqWE > 1e-5) ? 0.0 : w->snTempC;
  if (t == 0.0) return (latVapkJ * 0.622 * rhoAir * Kaero * (c->vPress - 6.111) / c->atmPress);
  return (latSubkJ * 0.622 * rhoAir * Kaero * (c->vPress - 6.112 * exp((17.67 * t) / (t + 243.5))) / c->atmPress);
}
static double sensible_hf(const EtCell *c, const Snow *w, double Kaero) {
  double t = (w->liqWE > 1e-5) ? 0.0 : w->snTempC;
  return (rhoAir * cpairkJ * Kaero * ((c->airTemp + 273.15) - (t + 273.15)));
}

/* tSnowPack.cpp:1298-1328 */
static double precip_hf(const EtCell *c, Snow *w) {
  double phf = 0;
  w->snPrec = (snow_frac(c) * (c->rain + ctom * w->snUnload)) * mtoc;
  w->liqPrec = ((1 - snow_frac(c)) * (c->rain + ctom * w->snUnload)) * mtoc;
  if (c->airTemp > 0)
    phf = (cmtonaught * w->snPrec * 0 * rholiqkg * cpicekJ + cmtonaught * w->liqPrec * (latFreezekJ + c->airTemp * rholiqkg * cpwaterkJ)) / 3600;
  else if (c->airTemp <= 0)
    phf = (cmtonaught * w->snPrec * c->airTemp * rholiqkg * cpicekJ + cmtonaught * w->liqPrec * latFreezekJ * rholiqkg) / 3600;
  return phf;
}

/* tSnowPack.cpp:1336-1365 */
static double aging_albedo(const Snow *w) {
  double alb;
  if (w->liqWE < 1.0E-5) alb = 0.85 * pow(0.96, pow(w->crustAge / 24.0, 0.58));
  else alb = 0.85 * pow(0.87, pow(w->crustAge / 24.0, 0.46));
  if (alb < 0.45) alb = 0.45;
  return alb;
}

/* tSnowPack.cpp:944-990 */
static void update_ripe(const EtCell *c, Snow *w, double precip) {
  w->snEvap = (1 - c->coeffV) * latent_hf(c, w, res_fact(c, w)) * w->timeSteps / (rholiqkg * latVapkJ);
  w->liqWE += cmtonaught * ((mtoc * precip) * (1 - snow_frac(c))) * w->timeSteps / 3600;
  if (w->liqWE + w->snEvap <= 0) { w->snEvap = w->liqWE; w->liqWE = 0; }
  else w->liqWE += w->snEvap;
  w->iceWE += cmtonaught * (mtoc * (precip * snow_frac(c))) * w->timeSteps / 3600;
  w->snSub = 0.0;
}
static void update_solid(const EtCell *c, Snow *w, double precip) {
  w->liqWE += cmtonaught * (mtoc * (precip * (1 - snow_frac(c)))) * w->timeSteps / 3600;
  w->snEvap = 0.0;
  w->snSub = (1 - c->coeffV) * latent_hf(c, w, res_fact(c, w)) * w->timeSteps / (rholiqkg * latSubkJ);
  w->iceWE += cmtonaught * (mtoc * (precip * snow_frac(c))) * w->timeSteps / 3600;
  if (w->iceWE + w->snSub <= 0) { w->snSub = w->iceWE; w->iceWE = 0; }
  else w->iceWE += w->snSub;
}

/* tSnowPack.cpp:1574-1748 (shelter options 0 and >= 4) */
static double in_short_wave_sn(EtCell *c, Snow *w, double *const *etf, int i, int snow_option) {
  double Is = 0, N = 0, Iv = 0, Isw = 0, Ir = 0, v, cosi, hillalbedo = 0;
  c->Ic = c->Id = c->Ids = c->Ics = 0.0;
  c->SunHour = 0.0;
  if (c->alphaD > 0.0) {
    direct_diffuse(c, c->elevation);
    N = 0;
    if (c->tsOption > 1) {
      double RadGlobClr = (c->inShortR / (1.0 - 0.65 * pow(N, 2.0)));
      c->Ic = c->Ic / (c->Ic * c->sinAlpha + c->Id) * RadGlobClr;
      c->Id = RadGlobClr - c->Ic * c->sinAlpha;
    }
    if (c->shOpt < 4) {
      cosi = 1.0 * (cos(c->slope) * c->sinAlpha + sin(c->slope) * cos(asin(c->sinAlpha)) * cos(c->sunaz - c->aspect));
      if (cosi >= 0.0) { c->Ics = c->Ic * cosi; c->SunHour = 1.0; }
      else { c->Ics = 0.0; c->SunHour = 0.0; }
    } else { c->Ics = 1.0 * c->Ic; c->SunHour = 1.0; }
    if ((c->shOpt == 2) || (c->shOpt == 1)) { c->Ics *= above_horizon(c); c->SunHour *= above_horizon(c); }
    if ((c->shOpt > 0) && (c->shOpt < 3)) v = c->shelterFactor;
    else if (c->shOpt == 0 || c->shOpt == 3) v = 0.5 * (1 + cos(c->slope));
    else v = 1.0;
    c->Ids = c->Id * v;
    Is = (1.0 - 0.65 * pow(N, 2)) * (c->Ics + c->Ids);
    if (w->hillAlbedoOption == 0) hillalbedo = w->albedo;
    else if (w->hillAlbedoOption == 1) hillalbedo = c->coeffAl;
    else if (w->hillAlbedoOption == 2) {
      if (w->snCanWE == 0) hillalbedo = c->coeffV * c->coeffAl + (1 - c->coeffV) * w->albedo;
      else hillalbedo = w->albedo;
    }
    if (c->shOpt == 0) { Ir = hillalbedo * Is * (1 - cos(c->slope)) * 0.5; Is += Ir; }
    else if ((c->shOpt > 1) && (c->shOpt < 4)) { Ir = hillalbedo * Is * (0.5 * (1 + cos(c->slope)) - c->shelterFactor); Is += Ir; }
    else Ir = 0.0;
    if ((c->etOpt == 1) || snow_option) Iv = Is * exp((c->coeffKt - 1) * c->coeffLAI) * c->coeffV + Is * (1.0 - c->coeffV);
    else Iv = Is;
    Isw = Iv * (1.0 - w->albedo);
  } else {
    c->Ic = Is = N = Iv = Isw = c->Id = c->Ids = c->Ics = Ir = 0.0;
  }
  c->Is = Is;
  if (c->tsOption > 1) F(ShortRadIn) = c->inShortR;
  else F(ShortRadIn) = Isw;
  F(ShortRadIn_dir) = c->Ics * (1.0 - 0.65 * pow(N, 2.0));
  F(ShortRadIn_dif) = (c->Ids + Ir) * (1.0 - 0.65 * pow(N, 2.0));
  return Isw;
}

/* tSnowPack.cpp:1940-2001 */
static void snow_eb(EtCell *c, Snow *w, double *const *etf, double *const *snf, int i) {
  double sigma = 5.67e-8, v1, snTempK, resFact;
  v1 = (c->shOpt > 0 && c->shOpt < 3) ? c->shelterFactor : 1;   /* tSnowPack.cpp:1945-1948 */
  snTempK = w->snTempC + 273.15;
  resFact = res_fact(c, w);
  w->H = sensible_hf(c, w, resFact);
  w->L = latent_hf(c, w, resFact);
  w->Prec = w->phfOnOff * precip_hf(c, w);
  w->RSin = naughttokilo * in_short_wave_sn(c, w, etf, i, 1);
  w->RLin = naughttokilo * in_long_wave(c);
  w->RLout = -naughttokilo * v1 * 0.9 * sigma * pow(snTempK, 4.0);
  c->inShortR = kilotonaught * w->RSin;
  c->inLongR = kilotonaught * w->RLin;
  c->outLongR = kilotonaught * w->RLout;
  F(HFlux) = 0.0; F(LFlux) = 0.0; F(LongRadIn) = 0.0; F(LongRadOut) = 0.0; F(ShortAbsbVeg) = 0.0; F(ShortAbsbSoi) = 0.0;
  SF(SnLHF) = w->L * kilotonaught; SF(SnSHF) = w->H * kilotonaught; SF(SnPHF) = w->Prec * kilotonaught;
  SF(SnGHF) = w->G * kilotonaught; SF(SnRLin) = w->RLin * kilotonaught; SF(SnRLout) = w->RLout * kilotonaught;
  SF(SnRSin) = w->RSin * kilotonaught;
  w->Rn = w->RSin + w->RLin + w->RLout;
  w->G = 0;
  w->dUint = (w->H + w->Rn + w->L + w->Prec + w->G) * w->timeSteps;
  w->Utot += w->dUint;
}

/* tSnowPack.cpp:1492-1568 */
static void compute_sub(const EtCell *c, Snow *w) {
  const double kc = 0.010, iceRad = 500e-6, Mwater = 18.01, R = 8313.0, RdryAir = 287.0, nu = 1.3e-5, KtAtm = 0.024, beta = 0.9;
  double Sp, acoefficient, windSpeedC, Re, Sh, Nu, esatIce, rhoVap, D, Omega, dmdt, psiS, Ce;
  Sp = REF_PI * pow(iceRad, 2.0) * (1 - w->albedo) * c->inShortR;
  acoefficient = beta * c->coeffLAI;
  windSpeedC = c->windSpeed * exp(-acoefficient * 0.4);
  Re = 2 * iceRad * windSpeedC / nu;
  Sh = 1.79 + 0.606 * pow(Re, 0.5);
  Nu = Sh;
  esatIce = 611.15 * exp(22.452 * (w->airTempK - 273.16) / (w->airTempK - 0.61));
  rhoVap = 0.622 * esatIce / (RdryAir * w->airTempK);
  D = 2.06e-5 * pow(w->airTempK / 273.0, 1.75);
  Omega = (1 / (KtAtm * w->airTempK * Nu)) * (1000 * latSubkJ * Mwater / (R * w->airTempK) - 1);
  dmdt = (2.0 * REF_PI * iceRad * (c->rHumidityC / 100 - 1) - Sp * Omega) / (1000 * latSubkJ * Omega + (1 / (D * rhoVap * Sh)));
  psiS = dmdt / ((4.0 / 3.0) * REF_PI * rhoicekg * iceRad * iceRad * iceRad);
  Ce = kc * pow(w->I / w->Imax, -0.4);
  w->Qcs = Ce * w->I * psiS * w->timeSteps;
}

static void set_to_node_snp(Snow *w, const orc_snow_args *sn, double *const *etf, double *const *snf, int i) {
  sn->LiqWE[i] = w->liqWE; sn->IceWE[i] = w->iceWE;
  SF(SnTempC) = w->snTempC; SF(CrustAge) = w->crustAge; SF(DensityAge) = w->densityAge; SF(EvapoTransAge) = w->ETAge;
  SF(SnSub) = w->snSub; SF(SnEvap) = w->snEvap;
  sn->LiqRouted[i] = w->liqRoute;
  SF(DU) = w->dUint; SF(Unode) = w->Utot; SF(Uerror) = w->Uerr;
  if ((w->iceWE + w->liqWE) <= 1e-5) { w->persMaxtemp = 0.0; w->inittimeTemp = 0.0; }
  if (w->snWE > 1e-5) {
    w->persMaxtemp += 1.0;
    if (w->persMaxtemp > w->persMax) {
      w->persMax = w->persMaxtemp;
      if (w->persMaxtemp - 1 < 1) w->inittimeTemp = w->hourly;
    }
  }
  if (w->snWE >= w->peakSnWE) { w->peakSnWE = w->snWE; w->peaktime = w->hourly; w->inittime = w->inittimeTemp; }
  SF(PersTimeMax) = w->persMax; SF(PersTime) = w->persMaxtemp; SF(PeakSWE) = w->peakSnWE;
  SF(PeakPackTime) = (double)w->peaktime; SF(InitPackTime) = (double)w->inittime; SF(InitPackTimeTemp) = (double)w->inittimeTemp;
  SF(CumLHF) += w->L; SF(CumSnSub) += w->snSub; SF(CumSnEvap) += w->snEvap; SF(CumMelt) += w->liqRoute;
  SF(CumSHF) += w->H * w->timeSteps; SF(CumPHF) += w->Prec * w->timeSteps; SF(CumGHF) += w->G * w->timeSteps;
  SF(CumRLin) += w->RLin * w->timeSteps; SF(CumRLout) += w->RLout * w->timeSteps;
  F(CumRSin) += w->RSin * w->timeSteps;
  SF(CumUerror) += w->Uerr * w->timeSteps;
  SF(CumHrsSnow) += w->snOnOff;
  w->L = w->H = w->Prec = w->G = w->RLin = w->RLout = w->RSin = w->dUint = 0.0;
  w->liqRoute = 0.0;
}

/* route liquid above the holding capacity (tSnowPack.cpp:536-553, 655-672) */
static void route_excess_liquid(Snow *w) {
  if (w->liqWE > w->snliqfrac * w->iceWE) {
    if (w->liqWE != w->snWE) {
      w->liqRoute = (w->liqWE - w->snliqfrac * w->iceWE);
      w->liqWE = w->liqWE - w->liqRoute;
      w->snWE = w->liqWE + w->iceWE;
    } else {
      w->liqRoute = w->snWE;
      w->liqWE = 0.0; w->iceWE = 0.0; w->snWE = 0.0;
    }
  }
}

int orc_snow_pack(const orc_et_basin *b, const orc_snow_args *sn, const orc_et_step *s, const orc_met *m,
                  double *const *etf, double *const *snf, const orc_et_shared *sh, int flag, double *members) {
  Snow W, *w = &W;
  if (b->shelter_option >= 1 && b->shelter_option <= 3 && !b->hor_angle_0000) return 1;
  if (b->et_option < 1 || b->et_option > 3) return 1;
  memset(w, 0, sizeof W);
  w->albedo = members[0]; w->Uerr = members[1]; w->snCanWE = members[2];
  w->timeStepm = b->metstep_min; w->timeSteph = w->timeStepm / 60.0; w->timeSteps = 60.0 * w->timeStepm;
  w->minSnTemp = sn->min_sn_temp; w->snliqfrac = sn->snliqfrac; w->hillAlbedoOption = sn->hill_albedo_option;
  w->hourly = sn->hourly_time_step;
  for (int i = 0; i < b->n; i++) {
    EtCell cell, *c = &cell;
    double precip = 0.0, vegHeight = 0;
    w->snOnOff = 0.0;
    met_assembly(b, s, m, etf, sh, i, c);
    c->snf = snf;
    /* getFrNodeSnP, tSnowPack.cpp:999-1044 */
    w->liqWE = sn->LiqWE[i]; w->iceWE = sn->IceWE[i]; w->liqRoute = 0.0;
    w->snWE = w->liqWE + w->iceWE;
    if (w->snWE > 1e-4) { w->snTempC = SF(SnTempC); w->iceTempC = w->snTempC; w->liqTempC = 0.0; }
    else if (c->airTemp > 0.0) { w->snTempC = 0.0; w->iceTempC = 0.0; w->liqTempC = 0.0; }
    else if (c->airTemp <= 0.0) { w->snTempC = c->airTemp; w->iceTempC = c->airTemp; w->liqTempC = c->airTemp; }
    w->crustAge = SF(CrustAge); w->densityAge = SF(DensityAge); w->ETAge = SF(EvapoTransAge);
    w->persMax = SF(PersTimeMax); w->persMaxtemp = SF(PersTime); w->peakSnWE = SF(PeakSWE); w->peakSnWEtemp = SF(PeakSWETemp);
    w->inittime = SF(InitPackTime); w->inittimeTemp = SF(InitPackTimeTemp); w->peaktime = SF(PeakPackTime);
    sn->LiqRouted[i] = 0.0;
    w->snUnload = 0.0;

    if ((w->snWE <= 1e-4) && (c->rain * snow_frac(c) <= 5e-2) && rholiqkg * cmtonaught * SF(IntSWE) < 1e-3) {
      /* no snow anywhere: the ET path */
      evap_method(b, etf, i, c);
      set_to_node(s, etf, i, c);
      components_cell(b, s, m, etf, sh, flag, i);
      w->snTempC = 0.0; w->snWE = 0.0; w->snSub = 0.0; w->snEvap = 0.0;
      w->dUint = w->RLin = w->RLout = w->RSin = w->H = w->L = w->G = w->Prec = 0.0;
      w->ETAge = w->ETAge + w->timeStepm;
      w->liqRoute = 0.0; w->iceWE = 0.0; w->liqWE = 0.0; w->Utot = 0.0; w->Usn = 0.0; w->Uwat = 0.0;
      w->snTempC = 0.0; w->crustAge = 0.0; w->densityAge = 0.0;
      SF(IntSub) = 0.0;
      set_to_node_snp(w, sn, etf, snf, i);
    } else {
      if (b->i_option && there_is_canopy(b, i)) {
        /* callSnowIntercept, tSnowPack.cpp:769-934 */
        double CanStorage = F(CanStorage), prec = sh->Rain[i], Isnow, throughfall, subFrac, unlFrac;
        int fl = 1;
        c->rHumidity = F(RelHumid); c->vPress = F(VapPressure); c->dewTemp = F(DewTemp); c->windSpeed = F(WindSpeed);
        c->atmPress = F(AirPressure); c->inShortR = F(ShortRadIn); c->airTemp = F(AirTemp);
        w->airTempK = c->airTemp + 273.15;
        w->LAI = c->coeffLAI;
        w->Iold = rholiqkg * cmtonaught * SF(IntSWE);
        w->Qcs = 0.0; w->Lm = 0.0;
        if ((prec * snow_frac(c) < 1e-4) && (w->Iold < 1e-3)) {
          w->I = w->Iold = w->Lm = w->Qcs = 0.0;
          evap_method(b, etf, i, c);
          F(NetRad) = 0.0; F(GFlux) = 0.0; F(HFlux) = 0.0; F(LFlux) = 0.0; F(LongRadOut) = 0.0;
          F(SurfTemp) = SF(SnTempC); F(SoilTemp) = SF(SnTempC);
          set_to_node(s, etf, i, c);
          F(ActEvap) = 0.0;
          components_cell(b, s, m, etf, sh, 1, i);
          SF(IntSWE) = 0; SF(IntSnUnload) = 0; SF(IntSub) = 0; SF(IntPrec) = 0;
          sh->EvapSoil[i] = 0.0;
        } else {
          if (CanStorage > 1e-5) { w->Iold += CanStorage; F(CanStorage) = 0.0; fl = 0; }
          w->Imax = 4.4 * w->LAI;
          Isnow = 0.7 * (w->Imax - w->Iold) * (1 - exp(-prec / w->Imax));
          w->I = w->Iold + Isnow;
          throughfall = prec - Isnow;
          if (w->Iold > 0.0 && fl) {
            compute_sub(c, w);
            if (w->airTempK >= 273.16) w->Lm = 5.8e-5 * (w->airTempK - 273.16) * w->timeSteps;
            else w->Lm = 0.0;
          } else { w->Qcs = 0.0; w->Lm = 0.0; }
          w->I += w->Qcs - w->Lm;
          if (w->I < 0.0) {
            if (w->Qcs < 0.0) { subFrac = fabs(w->Qcs) / (fabs(w->Qcs) + w->Lm); unlFrac = w->Lm / (fabs(w->Qcs) + w->Lm); }
            else { subFrac = 0.0; unlFrac = 1.0; }
            w->Qcs -= w->I * subFrac;
            w->Lm += w->I * unlFrac;
            w->I = 0.0;
          }
          w->Iold = w->I;
          SF(IntSWE) = naughttocm * (1 / rholiqkg) * w->I;
          SF(IntSnUnload) = naughttocm * (1 / rholiqkg) * w->Lm * c->coeffV;
          SF(IntSub) = naughttocm * (1 / rholiqkg) * w->Qcs * c->coeffV;
          SF(CumIntSub) += naughttocm * (1 / rholiqkg) * w->Qcs * c->coeffV;
          SF(CumIntUnl) += naughttocm * (1 / rholiqkg) * w->Lm * c->coeffV;
          SF(IntPrec) = Isnow * (1 / rholiqkg) * naughttocm;
          sh->NetPrecipitation[i] = (throughfall * c->coeffV) + (sh->Rain[i] * (1 - c->coeffV));
          F(EvapWetCanopy) = 0.0; sh->EvapDryCanopy[i] = 0.0; sh->EvapSoil[i] = 0.0; F(EvapoTrans) = 0.0; F(PotEvap) = 0.0;
        }
        w->snUnload = SF(IntSnUnload);
        w->snCanWE = SF(IntSWE);
      } else {
        sh->NetPrecipitation[i] = c->rain;
      }
      precip = sh->NetPrecipitation[i];
      precip += w->snUnload * ctom;
      w->snDepthm = cmtonaught * w->snWE / 0.312;
      w->iceWE = w->iceWE * cmtonaught; w->liqWE = w->liqWE * cmtonaught; w->snWE = w->iceWE + w->liqWE;
      vegHeight = (c->coeffH == 0) ? 0.1 : c->coeffH;
      (void)vegHeight;
      if (c->airTemp > 0.0) w->snTempC = 0.0; else w->snTempC = c->airTemp;
      if (w->snWE < 1e-5) {
        w->phfOnOff = 0.0; w->densityAge = 0.0; w->crustAge = 0.0;
        if (w->snTempC == 0.0) update_ripe(c, w, precip); else update_solid(c, w, precip);
        w->snWE = w->iceWE + w->liqWE;
        w->snSub *= naughttocm; w->snEvap *= naughttocm;
        w->L = w->H = w->G = w->Prec = w->Utotold = 0.0;
        route_excess_liquid(w);
        if (w->liqWE < 0) { w->liqWE = 0; w->snWE = w->iceWE; }
        w->Utot = w->dUint = w->iceWE * rhoicekg * cpicekJ * w->snTempC + w->liqWE * rholiqkg * latFreezekJ;
      } else {
        w->phfOnOff = 1.0;
        w->densityAge = (w->snWE * w->densityAge) / (w->snWE + mtoc * precip);
        if (precip * snow_frac(c) > 1e-3) w->crustAge = 0.0;
        if (w->snTempC == 0.0) update_ripe(c, w, precip); else update_solid(c, w, precip);
        w->snWE = w->iceWE + w->liqWE;
        w->snSub *= naughttocm; w->snEvap *= naughttocm;
        w->ETAge = 0.0;
        w->L = latent_hf(c, w, res_fact(c, w));
        w->Prec = precip_hf(c, w);
        if ((w->snWE <= 5e-6) || (w->snTempC < -800.0)) {
          w->liqRoute = 0.0; w->snWE = 0.0; w->iceWE = 0.0; w->liqWE = 0.0; w->Utot = 0.0; w->Usn = 0.0; w->Uwat = 0.0;
          w->snTempC = 0.0; w->crustAge = 0.0; w->densityAge = 0.0;
        } else {
          w->Utot = w->Utotold = w->iceWE * rhoicekg * cpicekJ * w->snTempC + w->liqWE * rholiqkg * latFreezekJ;
          w->albedo = aging_albedo(w);
          snow_eb(c, w, etf, snf, i);
          w->Uerr = (w->Utot - w->Utotold) - w->dUint;
          if (w->Utot < 0.0) {
            w->Usn = w->Utot; w->Uwat = 0.0; w->liqWatCont = 0.0; w->liqWE = 0.0; w->liqTempC = 0.0;
            w->iceWE = w->snWE;
            if (w->iceWE < 0.1 && w->iceWE > 0) w->iceTempC = w->Usn / (cpicekJ * rhoicekg * 0.1);
            else w->iceTempC = w->Usn / (cpicekJ * rhoicekg * w->snWE);
            if (w->iceTempC <= w->minSnTemp) w->iceTempC = w->minSnTemp;
            w->snTempC = w->iceTempC;
          } else {
            w->Uwat = w->Utot; w->Usn = 0.0;
            w->liqWE = w->Uwat / (latFreezekJ * rholiqkg);
            if (w->liqWE >= w->snWE) w->liqWE = w->snWE;
            w->iceWE = w->snWE - w->liqWE;
            route_excess_liquid(w);
            w->snTempC = 0.0; w->iceTempC = 0.0; w->liqTempC = 0.0;
          }
        }
      }
      if (w->snWE <= 5e-6) {
        w->liqRoute += w->snWE;
        w->snWE = 0.0; w->liqWE = 0.0; w->iceWE = 0.0; w->crustAge = 0.0; w->densityAge = 0.0;
        w->Utot = 0.0; w->Usn = 0.0; w->Uwat = 0.0;
      } else {
        w->crustAge += w->timeSteph; w->densityAge += w->timeSteph; w->snOnOff = 1.0;
      }
      w->snWE = naughttocm * w->snWE; w->iceWE = naughttocm * w->iceWE; w->liqWE = naughttocm * w->liqWE;
      w->liqRoute = naughttocm * w->liqRoute;
      sh->EvapSoil[i] = 0.0;
      F(ActEvap) = 0.0;
      F(EvapoTrans) = F(EvapWetCanopy) + sh->EvapDryCanopy[i];
      set_to_node_snp(w, sn, etf, snf, i);
    }
  }
  members[0] = w->albedo; members[1] = w->Uerr; members[2] = w->snCanWE;
  return 0;
}

int orc_snow_n_fields(void) { return ORC_N_SNOW_FIELDS; }
