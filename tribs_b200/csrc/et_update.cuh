// Kernel (1): per-cell surface energy balance / potential evaporation, ET partition and canopy
// interception. One thread per Voronoi cell, everything of a cell in registers, one read and
// one write of each per-cell array per call.
//
// Reference behaviour restated here (src/tHydro/tEvapoTrans.cpp, src/tHydro/tIntercept.cpp):
//   callEvapoPotential 562-831  -> et_potential()      energyBalance 2281-2357 -> energy_balance()
//   met equations 1224-1413     -> MetAir              rtsafe_mod_energy 2391-2474 -> solve_tg()
//   aeroResist/stomResist 1442-1512                    FunctionAndDerivative 2486-2632 -> residual()
//   inLongWave 2004-2037, inShortWave 1751-1930, DirectDiffuse 1939-1993 -> radiation()
//   ForceRestore 2663-2716      -> GroundHeat          betaFunc/betaFuncT 2911-3010 -> stress()
//   setToNode 3039-3076                                 ComputeETComponents 1039-1197 -> et_components()
//   tIntercept::InterceptGray 214-253, InterceptRutter 283-406, storageRungeKutta 431-459
//
// The reference recomputes the air-state functions (latent heat, vapour pressure, density, ...)
// inside every residual evaluation from unchanged inputs; they are pure, so they are evaluated
// once per cell here. Integer powers are written as products, pow(x, 0.5) as sqrt.
#pragma once
#include <stdint.h>
#include "physics.cuh"
#include "../../include/tribs_gpu.h"

namespace tribs {

struct LandClass {          // one row of the land table + values derived once on the host
  double A, B, P, S, K, b;  // interception: Gray storage/coefficient, Rutter throughfall/capacity/drainage
  double Al, H, Kt, Rs, V_raw, V, SE, ST, LAI;
  double rav_num;           // log((zm-d)/zom)*log((zm-d)/zov) for the class's vegetation height
};

struct EtConst {
  int32_t et_option, i_option, gflux_option, shelter_option;
  double metstep_min, etistep_h, tlinke, temp_lapse, max_interstorm;
  double ras_num;           // the same product of logs for bare soil (vegBare = 1)
  double tn_tlk, a1, a2, a3;   // Linke-turbidity polynomials of DirectDiffuse (functions of TLINKE only)
};

// ---- host-side set-up shared by tribs_gpu_et_init and the host check of this header ----------------
inline double et_log_product(double veg) {       // numerator of aeroResist (tEvapoTrans.cpp:1464-1475)
  const double zm = 2.0 + veg, zom = 0.123 * veg, zov = 0.0123 * veg, d = 0.67 * veg;
  return log((zm - d) / zom) * log((zm - d) / zov);
}
inline void et_derived_constants(EtConst &k) {   // needs k.tlinke
  k.ras_num = et_log_product(1.0);
  const double T = k.tlinke;                     // DirectDiffuse, tEvapoTrans.cpp:1974-1984
  k.tn_tlk = -0.015843 + 0.030543 * T + 0.0003797 * pow(T, 2.0);
  const double a1p = 0.26463 - 0.061581 * T + 0.0031408 * pow(T, 2.0);
  k.a1 = (a1p * k.tn_tlk < 0.0022) ? 0.0022 / k.tn_tlk : a1p;
  k.a2 = 2.04020 + 0.018945 * T - 0.011161 * pow(T, 2.0);
  k.a3 = -1.3025 + 0.039231 * T + 0.0085079 * pow(T, 2.0);
}
inline LandClass make_land_class(const double *lp, int et_option) {   // lp[k] = tInvariant::getLandProp(k)
  LandClass L;
  L.A = lp[1]; L.B = lp[2]; L.P = lp[3]; L.S = lp[4]; L.K = lp[5]; L.b = lp[6];
  L.Al = L.H = L.Kt = L.Rs = 0.0;
  double V = 0.0;                                // setCoeffs, tEvapoTrans.cpp:915-952
  if (et_option == 1) { L.Al = lp[7]; L.H = lp[8]; L.Kt = lp[9]; L.Rs = lp[10]; V = lp[11]; }
  else if (et_option == 2 || et_option == 3) { L.Al = lp[7]; V = lp[11]; }
  L.V_raw = lp[11];
  L.V = (V >= 1.0) ? 0.99 : V;
  L.SE = lp[13]; L.ST = lp[14]; L.LAI = lp[12];
  L.rav_num = et_log_product(L.H == 0 ? 0.1 : L.H);
  return L;
}

struct EtStatic {           // device order, length n_pad
  const double *z, *cos_slope, *sin_slope, *aspect, *heat_cond, *heat_cap, *soil_cutoff, *root_cutoff;
  const double *thr, *bedrock, *varea;
  const double *hor_angle_0000;   // OPTRADSHELT 1-3: tShelter's horizon angle at azimuth 0 [rad], else nullptr
  const int32_t *land_class, *boundary, *sweep_of_dev;
  const LandClass *classes;
};

struct MetDev {             // device copies of one tribs_met_t
  const int32_t *cell_record;
  const double *air_temp, *dew_temp, *rel_humid, *vap_press, *atm_press, *wind_speed, *sky_cover, *rad_global, *elev;
};

struct EtStepDev {
  int32_t do_potential, do_components, intercept_flag;
  int32_t met_time;                                // currentTime == met_hour: tWaterBalance::CanopyBalance is due
  double alphaD, sinAlpha, sunaz, Io, cos_alpha;   // cos_alpha = cos(asin(sinAlpha)), once per call
  int32_t hour, first_call, dew_hum_flag, ts_option, sky_flag_from;
  double te_met, te_eti, rs_ratio;                 // rs_ratio = rsRatio[hour] of stomResist()
  MetDev pot, cmp;
  int32_t snow_hourly;                             // batches with OPTSNOW 1: tEvapoTrans::hourlyTimeStep of this callSnowPack
  int32_t pad_;
};

struct EtView { double *f[TRIBS_N_ET_FIELDS]; };

// ---- air state of one met record at one cell ---------------------------------------------------
struct MetAir {
  double airTemp, atmPress, windSpeed, skyCover, inShortR;    // observed (airTemp lapse-corrected)
  double dewTemp, rHumidity, vPress;                          // after vaporPress()
  double dewTempC, vPressC, rHumidityC, atmPressC, windSpeedC;
  double lam, esat, P, cc, psy_raw, rho;

  TRIBS_HD void load(const MetDev &m, int i, double elevation, double temp_lapse, int dew_hum_flag) {
    const int r = m.cell_record[i];
    airTemp = m.air_temp[r];
    airTemp += temp_lapse * (elevation - m.elev[r]);
    dewTemp = m.dew_temp[r]; rHumidity = m.rel_humid[r]; vPress = m.vap_press[r];
    atmPress = m.atm_press[r]; windSpeed = m.wind_speed[r]; skyCover = m.sky_cover[r]; inShortR = m.rad_global[r];
    lam = ((597.3 - airTemp * 0.57) * (4186.8));
    esat = (6.112 * exp((17.67 * airTemp) / (airTemp + 243.5)));
    const double rv = 461.5, eo = 6.112, to = 273.15;
    if (dew_hum_flag == 0) {
      vPressC = esat * rHumidity / 100.0;
      dewTempC = 1.0 / ((1.0 / to) - (log(vPressC / eo) * rv / lam)) - 273.15;
      rHumidityC = rHumidity;
      dewTemp = dewTempC; vPress = vPressC;
    } else if (dew_hum_flag == 1) {
      dewTempC = dewTemp;
      vPressC = eo * exp((lam / rv) * ((1.0 / to) - (1.0 / (dewTemp + 273.15))));
      rHumidityC = 100.0 * vPressC / esat;
      rHumidity = rHumidityC; vPress = vPressC;
    } else {
      vPressC = vPress;
      rHumidityC = 100.0 * vPressC / esat;
      dewTempC = 1.0 / ((1.0 / to) - (log(vPressC / eo) * rv / lam)) - 273.15;
      rHumidity = rHumidityC; dewTemp = dewTempC;
    }
    atmPressC = (fabs(atmPress - 9999.99) < 1.0E-3) ? 1000.0 : atmPress;
    P = atmPressC + vPressC;
    const double airTempK = airTemp + 273.15;
    cc = 100.0 * (lam * esat) / (rv * (airTempK * airTempK));
    psy_raw = ((1013.0 * P) / (0.622 * lam));
    rho = (100.0 * P / (287.6 * airTempK)) * (1.0 - 0.378 * vPressC / P);
    windSpeedC = (windSpeed == 0.0 || fabs(windSpeed - 9999.99) < 1.0E-3) ? 0.1 : windSpeed;
  }
  TRIBS_HD double aero_resist(const LandClass &lc, double ras_num) const {
    const double k2 = 0.41 * 0.41;
    const double rav = lc.rav_num / (windSpeedC * k2);
    const double ras = ras_num / (windSpeedC * k2);
    return (1 - lc.V) * ras + lc.V * rav;
  }
  TRIBS_HD double sky_from_humidity(double rain) const {
    double sc = (rain > 0) ? 10.0 : round(10.0 * (3.2 * rHumidityC / 100.0 - 2.4) / 0.8);
    return sc < 0 ? 0.0 : (sc > 10 ? 10.0 : sc);
  }
};

// ---- soil-moisture stress factors ------------------------------------------------------------------
TRIBS_HD void stress(const LandClass &lc, double thr, double soil_m, double root_m, double soil_cut, double root_cut,
                     double &betaS, double &betaT) {
  double ratio = (soil_m - thr) / (lc.SE - thr);
  double beta = ratio > 1.0 ? 1.0 : (ratio < 0.0 ? 0.0 : ratio);
  betaS = (soil_m <= soil_cut) ? 0.0 : beta;
  ratio = (root_m - thr) / (lc.ST - thr);
  beta = ratio < 1.0 ? ratio : 1.0;
  betaT = (root_m <= root_cut) ? 0.0 : beta;
}

// ---- ground heat flux closure ----------------------------------------------------------------------
struct GroundHeat {
  int option;
  double Tso, Tlo, dt;
  double grad_coef;                 // option 1: sqrt(4 Ks Cs / (pi dt))
  double alpha, dg, w1;             // option 2: force-restore
  TRIBS_HD void setup(int opt, double ks, double cs, double metstep_min) {
    option = opt;
    const double pi = 4.0 * atan(1.0);
    dt = metstep_min * 60.0;
    grad_coef = sqrt((4.0 * ks * cs / (pi * dt)));
    w1 = 2. * pi / 86400.;
    const double d1 = sqrt((2.0 * (ks / cs) / w1));
    const double nd = 0.1 / d1;
    alpha = 1 + 0.943 * nd + 0.223 * (nd * nd) + 0.0168 * (nd * nd * nd) - 0.00527 * ((nd * nd) * (nd * nd));
    dg = (0.5 * cs * d1) / (1.0 + (w1 * dt) / (2.0 * sqrt(365.)));
  }
  // G(Tg) and dG/dTg exactly as the reference forms them (the finite difference of the
  // force-restore branch shifts only the tendency term)
  TRIBS_HD void eval(double Tg, double &G, double &dG) const {
    if (option == 1) {
      G = grad_coef * (Tg - Tso);
      dG = grad_coef;
    } else {
      G = dg * (alpha * ((Tg - Tso) / dt) + w1 * (Tg - Tlo));
      dG = (dg * (alpha * ((Tg + 0.001 - Tso) / dt) + w1 * (Tg - Tlo)) - G) / 0.001;
    }
  }
};

// ---- surface energy balance ------------------------------------------------------------------------
struct EnergyBalance {
  int et_option;
  double v1, inLongR, Rabsb, airTempK, Rah, Rstm, betaS, betaT, V, windSpeedC;
  double rho, P, satVap, es, lam, cc_over_psy, denom, denomrs;
  GroundHeat gh;
  // outputs of the last residual evaluation
  double hFlux, lFlux, gFlux, netRad, potEvap, outLongR;

  TRIBS_HD void residual(double Tg, double &fv, double &dv) {
    const double sigma = 5.67E-8, Es = 0.98, Cp = 1013.0, Ch = 0.0025, rv = 461.5, alpha = 1.26;
    const double Tg2 = Tg * Tg;
    outLongR = v1 * Es * sigma * (Tg2 * Tg2);
    const double Lsoi = outLongR - inLongR;
    double G, dGdTg;
    gh.eval(Tg, G, dGdTg);
    const double Rn = Rabsb - Lsoi;
    double H, LE = 0, Ep = 0, dHdTg, dlEdTg;
    const double dLdTg = 4.0 * Es * sigma * (Tg2 * Tg);
    const double dRndTg = -dLdTg;
    if (et_option == 2) {
      H = rho * Cp * (Tg - airTempK) * Ch * windSpeedC;
      const double tc = Tg - 273.15;
      const double esat = (6.112 * exp((17.67 * tc) / (tc + 243.5)));
      const double ccTs = (lam * esat) / (rv * Tg2);
      const double qhSatTs = (0.622 / P) * esat, qhTa = (0.622 / P) * es;
      Ep = (qhSatTs > qhTa) ? rho * Ch * (qhSatTs - qhTa) * windSpeedC : 0.0;
      LE = Ep * lam * betaS;
      dHdTg = rho * Cp * Ch * windSpeedC;
      dlEdTg = lam * rho * Ch * windSpeedC * (0.622 / P) * ccTs * betaS;
    } else {
      H = rho * Cp * (Tg - airTempK) / Rah;
      dHdTg = rho * Cp / Rah;
      if (et_option == 1) {
        const double dQ = (0.622 / P) * (satVap - es) * rho * lam / Rah;
        const double soiFct = (1.0 - V) * betaS, vegFct = V * betaT;
        const double num = cc_over_psy * (Rn - G) + dQ;
        const double Eps = num / denom / lam;
        Ep = num / denomrs / lam;
        LE = soiFct * lam * Eps + vegFct * lam * Ep;
        Ep *= (denomrs / denom);
        dlEdTg = soiFct * cc_over_psy * (dRndTg - dGdTg) / denom + vegFct * cc_over_psy * (dRndTg - dGdTg) / denomrs;
      } else {
        Ep = (Rn > G) ? (alpha / lam) * (Rn - G) * (cc_over_psy / denom) : 0.0;
        LE = Ep * lam * betaS;
        dlEdTg = betaS * (alpha) * (cc_over_psy / denom) * (dRndTg - dGdTg);
      }
    }
    fv = -Rabsb + Lsoi + H + LE + G;
    dv = dLdTg + dHdTg + dlEdTg + dGdTg;
    hFlux = H; lFlux = LE; netRad = Rn; gFlux = G; potEvap = Ep;
  }

  // safeguarded Newton of the reference: bracket [223.15, 373.15] K, start at the previous surface
  // temperature, bisect when Newton leaves the bracket or stalls, |dx| < 1e-5 ends, <= 100 steps
  TRIBS_HD double solve_tg(double guess) {
    const double x1 = 223.15, x2 = 373.15, xacc = 1.0e-5;
    double fl, fh, f, df, xl, xh;
    residual(x1, fl, df);
    if (fl == 0.0) return x1;
    residual(x2, fh, df);
    if (fh == 0.0) return x2;
    if (fl < 0.0) { xl = x1; xh = x2; } else { xh = x1; xl = x2; }
    double rts = guess, dxold = fabs(x2 - x1), dx = dxold;
    residual(rts, f, df);
    for (int j = 1; j <= 100; j++) {
      if ((((rts - xh) * df - f) * ((rts - xl) * df - f) > 0.0) || (fabs(2.0 * f) > fabs(dxold * df))) {
        dxold = dx;
        dx = 0.5 * (xh - xl);
        rts = xl + dx;
        if (xl == rts) return rts;
      } else {
        dxold = dx;
        dx = f / df;
        const double prev = rts;
        rts -= dx;
        if (prev == rts) return rts;
      }
      if (fabs(dx) < xacc) return rts;
      residual(rts, f, df);
      if (f < 0.0) xl = rts; else xh = rts;
    }
    return guess;
  }
};

// ---- canopy interception ---------------------------------------------------------------------------
struct Canopy {       // the interception fields of one cell
  double StormLength, CumIntercept, CanStorage, InterceptLoss, EvapWetCanopy, NetPrecipitation;
};

TRIBS_HD void intercept_gray(const LandClass &lc, const EtConst &k, double rainfall, Canopy &c) {
  if (rainfall <= 1.0) c.StormLength += k.etistep_h; else c.StormLength = 0.0;
  const double cum = c.CumIntercept;
  const double intercept = (cum <= lc.A) ? rainfall : rainfall * lc.B;
  c.InterceptLoss = intercept;
  c.CumIntercept = (c.StormLength < k.max_interstorm) ? cum + intercept * k.etistep_h : 0.0;
  c.NetPrecipitation = lc.V_raw * (rainfall - intercept) + (1 - lc.V_raw) * rainfall;
}

#if defined(__CUDA_ARCH__) && defined(TRIBS_CONST_MATH)
#define TRIBS_ET_EXP(x) cm::exp_c(x)      // constant-bank coefficients, const_math.cuh
#else
#define TRIBS_ET_EXP(x) exp(x)
#endif
TRIBS_HD double rutter_rate(const LandClass &lc, double C, double R, double Ep) {
  const double drain = lc.K * TRIBS_ET_EXP(lc.b * (C - lc.S));
  return (C < lc.S) ? ((1 - lc.P) * R - (C / lc.S) * Ep - drain) : ((1 - lc.P) * R - Ep - drain);
}

TRIBS_HD void intercept_rutter(const LandClass &lc, const EtConst &k, double rainfall, double Ep, Canopy &c) {
  const double dt = k.etistep_h, prev = c.CanStorage;
  double ctos = c.CanStorage / lc.S;
  if (ctos > 1) ctos = 0;
  if (rainfall <= 1.0) c.StormLength += dt; else c.StormLength = 0.0;
  if (prev > Ep * ctos * dt || rainfall > 0) {
    const double cum = c.CumIntercept;
    double numiter = floor(rainfall * 3.0);
    if (numiter < 50.0) numiter = 50.0;
    const double h = dt / numiter;
    double c0 = prev;
    for (double t0 = 0.0; t0 < dt; t0 += h) {      // classical RK4 on dC/dt
      const double ta = h * rutter_rate(lc, c0, rainfall, Ep);
      const double tb = h * rutter_rate(lc, c0 + ta / 2.0, rainfall, Ep);
      const double tc = h * rutter_rate(lc, c0 + tb / 2.0, rainfall, Ep);
      const double td = h * rutter_rate(lc, c0 + tc, rainfall, Ep);
      c0 += (ta / 6.0 + tb / 3.0 + tc / 3.0 + td / 6.0);
    }
    double cur = c0;
    const double on_canopy = (1 - lc.P) * rainfall * dt;
    if (cur > prev + on_canopy) cur = prev + on_canopy;
    const double avg = (prev + cur) / 2.0;
    double wet_rate = (avg < lc.S) ? (avg / lc.S) * Ep : Ep;
    if (wet_rate < 0) wet_rate = 0.0;
    double loss = on_canopy - (cur - prev);
    if (loss < 0.0) loss = 0.0;
    double wet_vol = wet_rate * dt;
    if (wet_vol > loss) wet_vol = loss;
    if (wet_vol < 0.0) wet_vol = 0.0;
    double drain_vol = loss - wet_vol;
    if (drain_vol < 0.0) drain_vol = 0.0;
    if (cur * 100 / lc.S <= 3) {                  // nearly empty canopy: dump what is left
      if ((wet_vol + cur) <= Ep * dt) wet_vol += cur; else drain_vol += cur;
      cur = 0.0;
    }
    double netp = (drain_vol / dt) + lc.P * rainfall;
    const double interception = (netp <= rainfall) ? rainfall - netp : 0.0;
    netp = lc.V_raw * netp + (1 - lc.V_raw) * rainfall;
    if (cur < 0.0) cur = 0.0;
    c.NetPrecipitation = netp;
    c.InterceptLoss = interception;
    c.CanStorage = cur;
    c.EvapWetCanopy = wet_vol / dt;
    c.CumIntercept = (c.StormLength < k.max_interstorm) ? cum + interception * dt : 0.0;
  } else {
    c.NetPrecipitation = 0.0; c.InterceptLoss = 0.0; c.CanStorage = 0.0; c.CumIntercept = 0.0; c.EvapWetCanopy = 0.0;
  }
}

// ---- the cell update -------------------------------------------------------------------------------
#define EF(name) ev.f[TRIBS_ET_##name][i]

// Radiation sheltering (OPTRADSHELT; tEvapoTrans.cpp:621-686, 1826-1887, aboveHorizon 2080-2270).
TRIBS_HD double shelter_factor(const EtConst &k, const EtView &ev, int i, double cos_s) {
  if (k.shelter_option > 0 && k.shelter_option < 4) return EF(SheltFact);   // tShelter's sky-view factor, static
  return (k.shelter_option == 0) ? 0.5 * (1 + cos_s) : 1.0;
}
// the sector test of aboveHorizon is always true: its first case (azimuth 0) decides
TRIBS_HD double above_horizon(double hor_angle_0000, double alphaD) { return (hor_angle_0000 * (180 / 3.1416) < alphaD) ? 1.0 : 0.0; }
TRIBS_HD double diffuse_view(const EtConst &k, double shelter, double cos_s) {
  if (k.shelter_option > 0 && k.shelter_option < 3) return shelter;
  return (k.shelter_option == 0 || k.shelter_option == 3) ? 0.5 * (1.0 + cos_s) : 1.0;
}

// kStore = false evaluates the cell without leaving anything behind (the snow path probes cells
// that way, snow_update.cuh). Returns the net precipitation below the canopy.
template <bool kStore>
TRIBS_HD double et_sweeps_t(const EtConst &k, const EtStatic &st, const EtStepDev &s, const EtView &ev, const DynView &dv, int i) {
  const LandClass lc = st.classes[st.land_class[i]];
  const double elevation = st.z[i];
  const double rain = dv.f[TRIBS_F_Rain][i];
  double betaS, betaT;
  stress(lc, st.thr[i], dv.f[TRIBS_F_SoilMoisture][i], dv.f[TRIBS_F_RootMoisture][i], st.soil_cutoff[i], st.root_cutoff[i],
         betaS, betaT);
  double rs = lc.Rs * s.rs_ratio;                  // stomResist(): closed stomata at night
  if (s.alphaD < 0.0) rs *= 1000.0;
  double potEvap = 0.0, actEvap = 0.0;

  if (s.do_potential && k.et_option) {
    MetAir a;
    a.load(s.pot, i, elevation, k.temp_lapse, s.dew_hum_flag);
    const double cos_s = st.cos_slope[i], sin_s = st.sin_slope[i];
    // sky-view factor: local slope only (OPTRADSHELT 0), tShelter's value kept in the node (1-3), none (>= 4)
    const double shelter = shelter_factor(k, ev, i, cos_s);
    const bool sky_flag = (st.sweep_of_dev[i] >= s.sky_flag_from) || (fabs(a.skyCover - 9999.99) < 1.0E-3);
    if (sky_flag) a.skyCover = a.sky_from_humidity(rain);
    double surfT, soilT;
    if (s.first_call) {
      const double t0 = s.pot.air_temp[s.pot.cell_record[i]] + 273.15;
      soilT = t0 - 273.15; surfT = t0 - 273.15;
    } else { surfT = EF(SurfTemp); soilT = EF(SoilTemp); }
    if (kStore && k.i_option == 0) dv.f[TRIBS_F_NetPrecipitation][i] = rain;

    EnergyBalance eb;
    eb.et_option = k.et_option;
    eb.gh.setup(k.gflux_option, st.heat_cond[i], st.heat_cap[i], k.metstep_min);
    eb.gh.Tso = surfT + 273.15;
    eb.gh.Tlo = soilT + 273.15;
    // long wave in (v0 = shelter factor for OPTRADSHELT < 3)
    const double v0 = (k.shelter_option < 3) ? shelter : 1.0;
    {
      const double N = a.skyCover / 10.0;
      const double kCloud = 1.0 + 0.17 * (N * N);
      const double Ea = 0.74 + 0.0049 * a.vPressC;
      const double tk = a.airTemp + 273.15, tk2 = tk * tk;
      eb.inLongR = v0 * kCloud * Ea * 5.67E-8 * (tk2 * tk2);
    }
    // short wave on the slope
    double Ics = 0.0, Ids = 0.0, Ir = 0.0, Is = 0.0, sunHour = 0.0, soilAbs = 0.0, vegAbs = 0.0, Isw = 0.0;
    if (s.alphaD > 0.0) {
      const double pi = 4.0 * atan(1.0);
      const double h0 = s.alphaD;
      const double pp0 = exp(-elevation / 8434.5);
      const double Dh0ref = 0.061359 * (0.1594 + 1.123 * h0 + 0.065656 * h0 * h0) / (1 + 28.9344 * h0 + 277.3971 * h0 * h0);
      const double h0ref = h0 + Dh0ref;
      const double m = pp0 / (sin(h0ref * pi / 180.) + 0.50572 * pow((h0ref + 6.07995), -1.6364));
      const double drm = (m <= 20.0) ? 1 / (6.6296 + 1.7513 * m - 0.1202 * m * m + 0.0065 * (m * m * m) - 0.00013 * ((m * m) * (m * m)))
                                     : 1 / (10.4 + 0.718 * m);
      double Ic = s.Io * exp(-0.8662 * k.tlinke * m * drm);
      const double Fdh0 = k.a1 + k.a2 * s.sinAlpha + k.a3 * (s.sinAlpha * s.sinAlpha);
      double Id = s.Io * k.tn_tlk * Fdh0;
      if (s.ts_option > 1) {                     // observed global radiation rescales both parts
        const double clr = (a.inShortR / 1.0);
        Ic = Ic / (Ic * s.sinAlpha + Id) * clr;
        Id = clr - Ic * s.sinAlpha;
      }
      if (k.shelter_option < 4) {
        const double cosi = cos_s * s.sinAlpha + sin_s * s.cos_alpha * cos(s.sunaz - st.aspect[i]);
        if (cosi >= 0) { Ics = Ic * cosi; sunHour = 1.0; }
      } else { Ics = Ic; sunHour = 1.0; }
      if (k.shelter_option == 1 || k.shelter_option == 2) {
        const double seen = above_horizon(st.hor_angle_0000[i], s.alphaD);
        Ics *= seen; sunHour *= seen;
      }
      const double v = diffuse_view(k, shelter, cos_s);
      Ids = Id * v;
      Is = (Ics + Ids);
      if (k.shelter_option == 0) { Ir = lc.Al * Is * (1 - cos_s) * 0.5; Is += Ir; }
      else if (k.shelter_option > 1 && k.shelter_option < 4) { Ir = lc.Al * Is * (0.5 * (1 + cos_s) - shelter); Is += Ir; }
      if (k.et_option == 1) {
        soilAbs = (Is * lc.Kt * lc.V + Is * (1.0 - lc.V)) * (1.0 - lc.Al);
        vegAbs = Is * lc.V * (1.0 - lc.Kt) * (1.0 - lc.Al);
        Isw = soilAbs;
      } else Isw = Is * (1.0 - lc.Al);
    }
    eb.Rabsb = soilAbs;                           // only the Penman-Monteith branch sets it (quirk kept)
    eb.v1 = v0;
    eb.airTempK = a.airTemp + 273.15;
    eb.Rah = a.aero_resist(lc, k.ras_num);
    eb.Rstm = rs;
    eb.betaS = betaS; eb.betaT = betaT; eb.V = lc.V; eb.windSpeedC = a.windSpeedC;
    eb.rho = a.rho; eb.P = a.P; eb.satVap = a.esat; eb.es = a.vPressC; eb.lam = a.lam;
    const double psy = 100.0 * a.psy_raw;
    eb.cc_over_psy = a.cc / psy;
    eb.denom = (eb.cc_over_psy + 1.0);
    eb.denomrs = (eb.cc_over_psy + 1.0 + eb.Rstm / eb.Rah);
    const double Tg = eb.solve_tg(eb.gh.Tso);
    double f, df;
    eb.residual(Tg, f, df);
    const double tlo = eb.gh.Tlo + (eb.gFlux * k.metstep_min * 60.0) /
                       (st.heat_cap[i] * 33.862683 * sqrt((st.heat_cond[i] / st.heat_cap[i] / 3.6361E-05)));
    potEvap = 3600.0 * eb.potEvap;
    actEvap = (k.et_option == 1) ? 3600.0 * (eb.lFlux / a.lam) : potEvap * betaS;

    if (kStore) {
    EF(SurfTemp) = Tg - 273.15; EF(SoilTemp) = tlo - 273.15;
    EF(NetRad) = eb.netRad; EF(GFlux) = eb.gFlux; EF(HFlux) = eb.hFlux; EF(LFlux) = eb.lFlux;
    EF(LongRadIn) = eb.inLongR; EF(LongRadOut) = eb.outLongR;
    EF(ShortRadIn) = (s.ts_option > 1) ? a.inShortR : Isw;
    EF(ShortRadIn_dir) = Ics; EF(ShortRadIn_dif) = (Ids + Ir);
    EF(ShortRadSlope) = (k.et_option == 1) ? Is : 0.0;
    EF(ShortAbsbVeg) = vegAbs; EF(ShortAbsbSoi) = soilAbs;
    EF(AirTemp) = a.airTemp; EF(DewTemp) = a.dewTemp; EF(RelHumid) = a.rHumidity; EF(VapPressure) = a.vPressC;
    EF(SkyCover) = a.skyCover; EF(WindSpeed) = a.windSpeedC; EF(AirPressure) = a.atmPressC;
    EF(CumHrsSun) += sunHour;
    double frac = 0;
    if ((fabs(eb.hFlux) + fabs(eb.lFlux)) > 1.0) frac = fabs(eb.lFlux) / (fabs(eb.hFlux) + fabs(eb.lFlux));
    if (fabs(s.te_met - 1.0) <= 1.0E-6) EF(AvEvapFract) = frac;
    else if (s.te_met > 1.0) EF(AvEvapFract) = (EF(AvEvapFract) * (s.te_met - 1.0) + frac) / s.te_met;
    EF(SheltFact) = shelter;
    EF(CumRSin) += a.inShortR * 3600.0;
    EF(PotEvap) = potEvap; EF(ActEvap) = actEvap;
    }
  } else if (s.do_components && k.et_option) {
    potEvap = EF(PotEvap); actEvap = EF(ActEvap);
  }

  if (!s.do_components) return dv.f[TRIBS_F_NetPrecipitation][i];
  Canopy can;
  can.StormLength = EF(StormLength); can.CumIntercept = EF(CumIntercept); can.CanStorage = EF(CanStorage);
  can.InterceptLoss = EF(InterceptLoss); can.EvapWetCanopy = EF(EvapWetCanopy);
  can.NetPrecipitation = dv.f[TRIBS_F_NetPrecipitation][i];
  const bool canopy = (k.i_option == 1) || (k.i_option == 2 && lc.P < 0.999);
  if (!k.et_option) {
    if (s.intercept_flag && canopy) {
      if (k.i_option == 1) intercept_gray(lc, k, rain, can); else intercept_rutter(lc, k, rain, 0.0, can);
      if (kStore) {
      EF(StormLength) = can.StormLength; EF(CumIntercept) = can.CumIntercept; EF(CanStorage) = can.CanStorage;
      EF(InterceptLoss) = can.InterceptLoss; EF(EvapWetCanopy) = can.EvapWetCanopy;
      dv.f[TRIBS_F_NetPrecipitation][i] = can.NetPrecipitation;
      }
    }
    return can.NetPrecipitation;
  }
  MetAir a;
  a.load(s.cmp, i, elevation, k.temp_lapse, s.dew_hum_flag);
  if (potEvap < 0) { potEvap = 0.0; if (kStore) EF(PotEvap) = 0.0; }
  if (actEvap < 0.0) { actEvap = 0.0; if (kStore) EF(ActEvap) = 0.0; }
  const double ra = a.aero_resist(lc, k.ras_num);
  // the reference mixes a x100 Clausius-Clapeyron slope with an unscaled psychrometric constant here
  const double transFactor = (a.cc + a.psy_raw) / (a.cc + a.psy_raw * (1 + rs / ra));
  double wet = 0.0;
  if (s.intercept_flag && (lc.V > 0) && canopy) {
    if (k.i_option == 1) {
      const double store = can.CumIntercept;
      wet = (store >= potEvap * k.etistep_h) ? potEvap : store / k.etistep_h;
      if (wet > potEvap) wet = potEvap;
      intercept_gray(lc, k, rain, can);
    } else {
      intercept_rutter(lc, k, rain, potEvap, can);
      wet = can.EvapWetCanopy;
    }
  } else {
    can.NetPrecipitation = rain;
  }
  double dry = betaT * ((potEvap - wet) * transFactor);
  double remaining = potEvap - wet - dry;
  if (remaining < 0.0) remaining = 0.0;
  wet *= lc.V; dry *= lc.V;
  const double soil = (st.bedrock[i] <= 0) ? 0.0 : (1 - lc.V) * remaining * betaS;
  const double et = wet + dry + soil;
  if (k.i_option == 1) can.CumIntercept = can.CumIntercept - wet / lc.V * k.etistep_h;
  if (!kStore) return can.NetPrecipitation;
  EF(StormLength) = can.StormLength; EF(CumIntercept) = can.CumIntercept; EF(CanStorage) = can.CanStorage;
  EF(InterceptLoss) = can.InterceptLoss;
  dv.f[TRIBS_F_NetPrecipitation][i] = can.NetPrecipitation;
  EF(EvapWetCanopy) = wet;
  dv.f[TRIBS_F_EvapDryCanopy][i] = dry;
  dv.f[TRIBS_F_EvapSoil][i] = soil;
  EF(EvapoTrans) = et;
  EF(CumTotEvap) += et;
  EF(CumBarEvap) += soil;
  if (fabs(s.te_eti - 1.0) < 1.0E-6) EF(AvET) = et;
  else if (s.te_eti > 1.0) EF(AvET) = (EF(AvET) * (s.te_eti - 1.0) + et) / s.te_eti;
  return can.NetPrecipitation;
}
TRIBS_HD void et_sweeps(const EtConst &k, const EtStatic &st, const EtStepDev &s, const EtView &ev, const DynView &dv, int i) {
  et_sweeps_t<true>(k, st, s, ev, dv, i);
}

// the two sweeps, then the canopy water balance that Simulator::UpdateWaterBalance applies at the
// end of a met step to the values they leave behind (tWaterBalance.cpp:116-140)
TRIBS_HD void et_cell(const EtConst &k, const EtStatic &st, const EtStepDev &s, const EtView &ev, const DynView &dv, int i) {
  et_sweeps(k, st, s, ev, dv, i);
  if (s.met_time) {
    const double del_i = EF(InterceptLoss) - EF(EvapWetCanopy);
    EF(CanopyStorVol) = EF(CanopyStorVol) + del_i * (k.metstep_min / 60.0) * st.varea[i] / 1000.0;
  }
}
#undef EF

}  // namespace tribs
