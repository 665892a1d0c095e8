// Snow pack cell update (tSnowPack::callSnowPack incl. callSnowIntercept). One thread per cell.
// Reference behaviour restated here (src/tHydro/tSnowPack.cpp):
//   getFrNodeSnP 999-1044 / setToNodeSnP 1052-1135 -> load / store of SnowCell
//   updateRipeSnowPack / updateSolidSnowPack 944-990, snowFracCalc 1244-1290, resFactCalc 1373-1480,
//   latentHFCalc / sensibleHFCalc 1177-1236, precipitationHFCalc 1298-1328, agingAlbedo 1336-1365,
//   inShortWaveSn 1574-1748, snowEB 1940-2001, the pack bookkeeping of callSnowPack 440-700.
//   callSnowIntercept 769-934, computeSub 1492-1568 -> snow_canopy_sublimation().
// Cells without snow on the ground, in the air or in the canopy take the ET path (et_sweeps).
//
// With a canopy the reference reads two tSnowPack members before the current cell has written them:
// `albedo` (computeSub) and, under HILLALBOPT 2, `snCanWE` (inShortWaveSn of a cell without canopy).
// They therefore hold what the last cell *before it in sweep order* left there (or the last one of
// the previous call). The device reproduces that with a last-writer scan over sweep positions:
// snow_cell_t<false> ("probe") evaluates a cell without storing anything and reports whether and what
// it would write to the two members; the probe is repeated for the cells whose answer can depend on
// the incoming albedo until no answer changes; snow_cell_t<true> then commits with the final values.
#pragma once
#include "et_update.cuh"

namespace tribs {

struct SnowConst {
  double min_sn_temp, snliqfrac;
  int32_t hill_albedo_option, hourly;
};
struct SnowView { double *f[TRIBS_N_SNOW_FIELDS]; };

// what a cell found in / leaves in the tSnowPack members that outlive it (see the header comment)
struct SnowProbe {
  int32_t alb_writer, can_writer, sensitive;   // sensitive: the cell's results depend on the incoming albedo
  double alb_value, can_value;
};

struct SnowCell {
  double liqWE, iceWE, snWE, snSub, snEvap, liqRoute, Utot, snTempC, crustAge, densityAge, ETAge;
  double H, L, G, Prec, dUint, RLin, RLout, RSin, Uerr, snDepthm, snOnOff, phfOnOff, albedo;
  double persMax, persMaxtemp, peakSnWE, inittime, inittimeTemp, peaktime;
};

namespace snowk {
constexpr double rholiqkg = 1000.0, rhoicekg = 920.0, rhoAir = 1.3, cpicekJ = 2.1, cpwaterkJ = 4.190, cpairkJ = 1.006;
constexpr double latFreezekJ = 334.0, latVapkJ = 2470.0, latSubkJ = 334.0 + 2470.0;
constexpr double naughttokilo = 1e-3, kilotonaught = 1e3, naughttocm = 100.0, cmtonaught = 0.01, mtoc = 0.1;
}  // namespace snowk

TRIBS_HD double snow_fraction(double Ta, double RH) {
  const double Tw = Ta * atan(0.151977 * sqrt(RH + 8.313659)) + atan(Ta + RH) - atan(RH - 1.676331) +
                    0.00391838 * (RH * sqrt(RH)) * atan(0.023101 * RH) - 4.686035;
  if (Tw > 5) return 0.0;
  return 1 / (1 + 6.99 * 1e-5 * exp(2 * (Tw + 3.97)));
}

TRIBS_HD double snow_res_fact(const MetAir &a, const LandClass &lc, const SnowCell &w) {
  const double k2 = 0.41 * 0.41, g = 9.81, Ri_cr = 0.1;
  const double windSpeedC = (a.windSpeed == 0.0 || fabs(a.windSpeed - 9999.99) < 1e-3) ? 0.01 : a.windSpeed;
  double vegHeight = (lc.H <= 0) ? 0.1 : lc.H;
  vegHeight = (vegHeight > w.snDepthm) ? vegHeight - w.snDepthm : 0.1;
  const double windSpeedS = (w.snDepthm < lc.H) ? windSpeedC * exp(-0.5 * lc.LAI * (1.0 - (w.snDepthm / lc.H))) : windSpeedC;
  double zm = 2.0 + vegHeight, zom = 0.123 * vegHeight, zov = 0.0123 * vegHeight, d = 0.67 * vegHeight;
  const double rav = log((zm - d) / zom) * log((zm - d) / zov) / (windSpeedC * k2);
  zm = 2.0 + 0.1; zom = 0.123 * 0.1; zov = 0.0123 * 0.1; d = 0.67 * 0.1;
  const double ras = log((zm - d) / zom) * log((zm - d) / zov) / (windSpeedS * k2);
  double ra = (1 - lc.V) * ras + lc.V * rav;
  const double Ts = (w.liqWE > 1e-5) ? 273.15 : w.snTempC + 273.15;
  const double Ta = a.airTemp + 273.15, T_avg = 0.5 * (Ta + Ts);
  const double zm_eff = fmax(0.1, 2.0 - w.snDepthm), z0 = 0.123 * vegHeight;
  const double RiB = (g * zm_eff * (Ta - Ts)) / (T_avg * windSpeedS * windSpeedS);
  const double Ri_u = 1.0 / (log(zm_eff / z0) + 5.0);
  double C_stab;
  if (RiB < 0.0) C_stab = sqrt(1.0 - 16.0 * RiB);
  else if (RiB <= Ri_u) { const double q = 1.0 - (RiB / Ri_cr); C_stab = 1.0 / (q * q); }
  else { const double q = 1.0 - (Ri_u / Ri_cr); C_stab = 1.0 / (q * q); }
  if (C_stab > 15.0) C_stab = 15.0;
  ra *= C_stab;
  return 1.0 / ra;
}

TRIBS_HD double snow_latent_hf(const MetAir &a, const SnowCell &w, double Kaero) {
  using namespace snowk;
  const double t = (w.liqWE > 1e-5) ? 0.0 : w.snTempC;
  if (t == 0.0) return (latVapkJ * 0.622 * rhoAir * Kaero * (a.vPress - 6.111) / a.atmPress);
  return (latSubkJ * 0.622 * rhoAir * Kaero * (a.vPress - 6.112 * exp((17.67 * t) / (t + 243.5))) / a.atmPress);
}
TRIBS_HD double snow_sensible_hf(const MetAir &a, const SnowCell &w, double Kaero) {
  using namespace snowk;
  const double t = (w.liqWE > 1e-5) ? 0.0 : w.snTempC;
  return (rhoAir * cpairkJ * Kaero * ((a.airTemp + 273.15) - (t + 273.15)));
}
TRIBS_HD double snow_precip_hf(const MetAir &a, double rain, double sfrac, double snUnload) {
  using namespace snowk;
  const double snPrec = (sfrac * (rain + 10.0 * snUnload)) * mtoc, liqPrec = ((1 - sfrac) * (rain + 10.0 * snUnload)) * mtoc;
  if (a.airTemp > 0)
    return (cmtonaught * snPrec * 0 * rholiqkg * cpicekJ + cmtonaught * liqPrec * (latFreezekJ + a.airTemp * rholiqkg * cpwaterkJ)) / 3600;
  return (cmtonaught * snPrec * a.airTemp * rholiqkg * cpicekJ + cmtonaught * liqPrec * latFreezekJ * rholiqkg) / 3600;
}
TRIBS_HD double snow_aging_albedo(const SnowCell &w) {
  double alb = (w.liqWE < 1.0E-5) ? 0.85 * pow(0.96, pow(w.crustAge / 24.0, 0.58)) : 0.85 * pow(0.87, pow(w.crustAge / 24.0, 0.46));
  return alb < 0.45 ? 0.45 : alb;
}
TRIBS_HD void snow_update_pack(const MetAir &a, const LandClass &lc, SnowCell &w, double precip, double sfrac, double time_s) {
  using namespace snowk;
  const double lhf = snow_latent_hf(a, w, snow_res_fact(a, lc, w));
  if (w.snTempC == 0.0) {   // ripe pack: evaporation from the liquid phase
    w.snEvap = (1 - lc.V) * lhf * time_s / (rholiqkg * latVapkJ);
    w.liqWE += cmtonaught * ((mtoc * precip) * (1 - sfrac)) * time_s / 3600;
    if (w.liqWE + w.snEvap <= 0) { w.snEvap = w.liqWE; w.liqWE = 0; } else w.liqWE += w.snEvap;
    w.iceWE += cmtonaught * (mtoc * (precip * sfrac)) * time_s / 3600;
    w.snSub = 0.0;
  } else {                  // frozen pack: sublimation from the ice phase
    w.liqWE += cmtonaught * (mtoc * (precip * (1 - sfrac))) * time_s / 3600;
    w.snEvap = 0.0;
    w.snSub = (1 - lc.V) * lhf * time_s / (rholiqkg * latSubkJ);
    w.iceWE += cmtonaught * (mtoc * (precip * sfrac)) * time_s / 3600;
    if (w.iceWE + w.snSub <= 0) { w.snSub = w.iceWE; w.iceWE = 0; } else w.iceWE += w.snSub;
  }
}
TRIBS_HD void snow_route_excess(SnowCell &w, double snliqfrac) {
  if (w.liqWE > snliqfrac * w.iceWE) {
    if (w.liqWE != w.snWE) {
      w.liqRoute = (w.liqWE - snliqfrac * w.iceWE);
      w.liqWE = w.liqWE - w.liqRoute;
      w.snWE = w.liqWE + w.iceWE;
    } else { w.liqRoute = w.snWE; w.liqWE = 0.0; w.iceWE = 0.0; w.snWE = 0.0; }
  }
}

// computeSub, tSnowPack.cpp:1492-1568: sublimation of intercepted snow over one met step [mm]
TRIBS_HD double snow_canopy_sublimation(const MetAir &a, const LandClass &lc, double albedo, double I, double Imax,
                                        double airTempK, double time_s) {
  using namespace snowk;
  const double PI = 3.1415926;
  const double kc = 0.010, iceRad = 500e-6, Mwater = 18.01, R = 8313.0, RdryAir = 287.0, nu = 1.3e-5, KtAtm = 0.024, beta = 0.9;
  const double Sp = PI * (iceRad * iceRad) * (1 - albedo) * a.inShortR;
  const double windSpeedC = a.windSpeed * exp(-(beta * lc.LAI) * 0.4);
  const double Re = 2 * iceRad * windSpeedC / nu;
  const double Sh = 1.79 + 0.606 * sqrt(Re);
  const double Nu = Sh;
  const double esatIce = 611.15 * exp(22.452 * (airTempK - 273.16) / (airTempK - 0.61));
  const double rhoVap = 0.622 * esatIce / (RdryAir * airTempK);
  const double D = 2.06e-5 * pow(airTempK / 273.0, 1.75);
  const double Omega = (1 / (KtAtm * airTempK * Nu)) * (1000 * latSubkJ * Mwater / (R * airTempK) - 1);
  const double dmdt = (2.0 * PI * iceRad * (a.rHumidityC / 100 - 1) - Sp * Omega) / (1000 * latSubkJ * Omega + (1 / (D * rhoVap * Sh)));
  const double psiS = dmdt / ((4.0 / 3.0) * PI * rhoicekg * iceRad * iceRad * iceRad);
  const double Ce = kc * pow(I / Imax, -0.4);
  return Ce * I * psiS * time_s;
}

#define EF(name) ev.f[TRIBS_ET_##name][i]
#define SNF(name) sv.f[TRIBS_SN_##name][i]

TRIBS_HD void snow_store(SnowCell &w, const SnowConst &sc, const SnowView &sv, const EtView &ev, const DynView &dv, int i, double time_s) {
  dv.f[TRIBS_F_LiqWE][i] = w.liqWE; dv.f[TRIBS_F_IceWE][i] = w.iceWE;
  SNF(SnTempC) = w.snTempC; SNF(CrustAge) = w.crustAge; SNF(DensityAge) = w.densityAge; SNF(EvapoTransAge) = w.ETAge;
  SNF(SnSub) = w.snSub; SNF(SnEvap) = w.snEvap;
  dv.f[TRIBS_F_LiqRouted][i] = w.liqRoute;
  SNF(DU) = w.dUint; SNF(Unode) = w.Utot; SNF(Uerror) = w.Uerr;
  if ((w.iceWE + w.liqWE) <= 1e-5) { w.persMaxtemp = 0.0; w.inittimeTemp = 0.0; }
  if (w.snWE > 1e-5) {
    w.persMaxtemp += 1.0;
    if (w.persMaxtemp > w.persMax) {
      w.persMax = w.persMaxtemp;
      if (w.persMaxtemp - 1 < 1) w.inittimeTemp = sc.hourly;
    }
  }
  if (w.snWE >= w.peakSnWE) { w.peakSnWE = w.snWE; w.peaktime = sc.hourly; w.inittime = w.inittimeTemp; }
  SNF(PersTimeMax) = w.persMax; SNF(PersTime) = w.persMaxtemp; SNF(PeakSWE) = w.peakSnWE;
  SNF(PeakPackTime) = w.peaktime; SNF(InitPackTime) = w.inittime; SNF(InitPackTimeTemp) = w.inittimeTemp;
  SNF(CumLHF) += w.L; SNF(CumSnSub) += w.snSub; SNF(CumSnEvap) += w.snEvap; SNF(CumMelt) += w.liqRoute;
  SNF(CumSHF) += w.H * time_s; SNF(CumPHF) += w.Prec * time_s; SNF(CumGHF) += w.G * time_s;
  SNF(CumRLin) += w.RLin * time_s; SNF(CumRLout) += w.RLout * time_s;
  EF(CumRSin) += w.RSin * time_s;
  SNF(CumUerror) += w.Uerr * time_s;
  SNF(CumHrsSnow) += w.snOnOff;
}

// One cell of callSnowPack. alb_in / can_in: the tSnowPack members `albedo` and `snCanWE` as the
// cell finds them; pr: what it leaves there. kStore = false touches no array.
template <bool kStore>
TRIBS_HD void snow_cell_t(const EtConst &k, const EtStatic &st, const EtStepDev &s, const SnowConst &sc, const EtView &ev,
                          const SnowView &sv, const DynView &dv, int i, double alb_in, double can_in, SnowProbe &pr) {
  using namespace snowk;
  pr.alb_writer = 0; pr.can_writer = 0; pr.sensitive = 0; pr.alb_value = 0.0; pr.can_value = 0.0;
  const LandClass lc = st.classes[st.land_class[i]];
  const double rain = dv.f[TRIBS_F_Rain][i];
  const double time_m = k.metstep_min, time_h = time_m / 60.0, time_s = 60.0 * time_m;
  MetAir a;
  a.load(s.pot, i, st.z[i], k.temp_lapse, s.dew_hum_flag);
  const double sfrac = snow_fraction(a.airTemp, a.rHumidity);
  SnowCell w{};
  w.liqWE = dv.f[TRIBS_F_LiqWE][i]; w.iceWE = dv.f[TRIBS_F_IceWE][i];
  w.snWE = w.liqWE + w.iceWE;
  const double snTempNode = SNF(SnTempC);
  if (w.snWE > 1e-4) w.snTempC = snTempNode;
  else w.snTempC = (a.airTemp > 0.0) ? 0.0 : a.airTemp;
  w.crustAge = SNF(CrustAge); w.densityAge = SNF(DensityAge); w.ETAge = SNF(EvapoTransAge);
  w.persMax = SNF(PersTimeMax); w.persMaxtemp = SNF(PersTime); w.peakSnWE = SNF(PeakSWE);
  w.inittime = SNF(InitPackTime); w.inittimeTemp = SNF(InitPackTimeTemp); w.peaktime = SNF(PeakPackTime);
  w.Uerr = SNF(Uerror);
  w.albedo = alb_in;
  const double intSWE_node = SNF(IntSWE);

  if ((w.snWE <= 1e-4) && (rain * sfrac <= 5e-2) && rholiqkg * cmtonaught * intSWE_node < 1e-3) {
    if (!kStore) return;
    et_sweeps(k, st, s, ev, dv, i);                      // no snow anywhere: kernel (1)'s two sweeps
    SNF(SnLHF) = 0.0; SNF(SnSHF) = 0.0; SNF(SnPHF) = 0.0; SNF(SnGHF) = 0.0; SNF(SnRLin) = 0.0; SNF(SnRLout) = 0.0; SNF(SnRSin) = 0.0;
    w.snTempC = 0.0; w.snWE = 0.0; w.snSub = 0.0; w.snEvap = 0.0;
    w.ETAge = w.ETAge + time_m;
    w.liqRoute = 0.0; w.iceWE = 0.0; w.liqWE = 0.0; w.Utot = 0.0; w.crustAge = 0.0; w.densityAge = 0.0;
    SNF(IntSub) = 0.0;
    snow_store(w, sc, sv, ev, dv, i, time_s);
    return;
  }
  // ---- met assembly of callSnowPack for a snow cell (node fields the ET path would have written)
  const double cos_s = st.cos_slope[i], sin_s = st.sin_slope[i];
  const double shelter = shelter_factor(k, ev, i, cos_s);
  const bool sky_flag = (st.sweep_of_dev[i] >= s.sky_flag_from) || (fabs(a.skyCover - 9999.99) < 1.0E-3);
  if (sky_flag) a.skyCover = a.sky_from_humidity(rain);
  if (kStore) {
    if (k.shelter_option == 0) EF(SheltFact) = shelter;
    EF(AirTemp) = a.airTemp; EF(DewTemp) = a.dewTemp; EF(RelHumid) = a.rHumidity; EF(VapPressure) = a.vPressC;
    EF(SkyCover) = a.skyCover; EF(WindSpeed) = a.windSpeed; EF(AirPressure) = a.atmPress;
    if (s.first_call) {
      const double t0 = s.pot.air_temp[s.pot.cell_record[i]] + 273.15;
      EF(SoilTemp) = t0 - 273.15; EF(SurfTemp) = t0 - 273.15;
    }
  }
  double precip = rain, snUnload = 0.0, snCanWE = can_in;
  bool shortrad_from_met = true;                          // F(ShortRadIn) = observed value unless the ET path ran
  if ((k.i_option == 1) || (k.i_option == 2 && lc.P < 0.999)) {
    // ---- callSnowIntercept, tSnowPack.cpp:769-934
    const double canStorage = EF(CanStorage), prec = rain;
    const double airTempK = a.airTemp + 273.15;
    double Iold = rholiqkg * cmtonaught * intSWE_node;
    if ((prec * sfrac < 1e-4) && (Iold < 1e-3)) {
      // nothing frozen in the canopy: the soil energy balance and the rain interception run under the
      // pack, their fluxes are then cleared (the reference's choice)
      precip = et_sweeps_t<kStore>(k, st, s, ev, dv, i);
      shortrad_from_met = false;
      if (kStore) {
        EF(NetRad) = 0.0; EF(GFlux) = 0.0; EF(HFlux) = 0.0; EF(LFlux) = 0.0; EF(LongRadOut) = 0.0;
        EF(SurfTemp) = snTempNode; EF(SoilTemp) = snTempNode;
        EF(ActEvap) = 0.0;
        SNF(SnLHF) = 0.0; SNF(SnSHF) = 0.0; SNF(SnPHF) = 0.0; SNF(SnGHF) = 0.0; SNF(SnRLin) = 0.0; SNF(SnRLout) = 0.0; SNF(SnRSin) = 0.0;
        SNF(IntSWE) = 0.0; SNF(IntSnUnload) = 0.0; SNF(IntSub) = 0.0; SNF(IntPrec) = 0.0;
        dv.f[TRIBS_F_EvapSoil][i] = 0.0;
      }
      snUnload = 0.0; snCanWE = 0.0;
    } else {
      bool fl = true;
      if (canStorage > 1e-5) { Iold += canStorage; fl = false; if (kStore) EF(CanStorage) = 0.0; }
      const double Imax = 4.4 * lc.LAI;
      const double Isnow = 0.7 * (Imax - Iold) * (1 - exp(-prec / Imax));
      double I = Iold + Isnow, Qcs = 0.0, Lm = 0.0;
      const double throughfall = prec - Isnow;
      if (Iold > 0.0 && fl) {
        Qcs = snow_canopy_sublimation(a, lc, alb_in, I, Imax, airTempK, time_s);
        pr.sensitive = 1;
        Lm = (airTempK >= 273.16) ? 5.8e-5 * (airTempK - 273.16) * time_s : 0.0;
      }
      I += Qcs - Lm;
      if (I < 0.0) {
        double subFrac, unlFrac;
        if (Qcs < 0.0) { subFrac = fabs(Qcs) / (fabs(Qcs) + Lm); unlFrac = Lm / (fabs(Qcs) + Lm); }
        else { subFrac = 0.0; unlFrac = 1.0; }
        Qcs -= I * subFrac;
        Lm += I * unlFrac;
        I = 0.0;
      }
      const double intSWE = naughttocm * (1 / rholiqkg) * I;
      const double unload = naughttocm * (1 / rholiqkg) * Lm * lc.V;
      const double sub = naughttocm * (1 / rholiqkg) * Qcs * lc.V;
      precip = (throughfall * lc.V) + (rain * (1 - lc.V));
      if (kStore) {
        SNF(IntSWE) = intSWE; SNF(IntSnUnload) = unload; SNF(IntSub) = sub;
        SNF(CumIntSub) += sub; SNF(CumIntUnl) += unload;
        SNF(IntPrec) = Isnow * (1 / rholiqkg) * naughttocm;
        dv.f[TRIBS_F_NetPrecipitation][i] = precip;
        EF(EvapWetCanopy) = 0.0; dv.f[TRIBS_F_EvapDryCanopy][i] = 0.0; dv.f[TRIBS_F_EvapSoil][i] = 0.0;
        EF(EvapoTrans) = 0.0; EF(PotEvap) = 0.0;
      }
      snUnload = unload; snCanWE = intSWE;
    }
    pr.can_writer = 1; pr.can_value = snCanWE;
  } else if (kStore) {
    dv.f[TRIBS_F_NetPrecipitation][i] = rain;
  }
  precip += snUnload * 10.0;
  w.snDepthm = cmtonaught * w.snWE / 0.312;
  w.iceWE = w.iceWE * cmtonaught; w.liqWE = w.liqWE * cmtonaught; w.snWE = w.iceWE + w.liqWE;
  w.snTempC = (a.airTemp > 0.0) ? 0.0 : a.airTemp;
  double shortIn = a.inShortR, Ics = 0.0, IdsIr = 0.0;
  if (w.snWE < 1e-5) {                                   // a pack is born this step
    w.phfOnOff = 0.0; w.densityAge = 0.0; w.crustAge = 0.0;
    snow_update_pack(a, lc, w, precip, sfrac, time_s);
    w.snWE = w.iceWE + w.liqWE;
    w.snSub *= naughttocm; w.snEvap *= naughttocm;
    snow_route_excess(w, sc.snliqfrac);
    if (w.liqWE < 0) { w.liqWE = 0; w.snWE = w.iceWE; }
    w.Utot = w.dUint = w.iceWE * rhoicekg * cpicekJ * w.snTempC + w.liqWE * rholiqkg * latFreezekJ;
    if (kStore && shortrad_from_met) EF(ShortRadIn) = a.inShortR;
  } else {
    w.phfOnOff = 1.0;
    w.densityAge = (w.snWE * w.densityAge) / (w.snWE + mtoc * precip);
    if (precip * sfrac > 1e-3) w.crustAge = 0.0;
    snow_update_pack(a, lc, w, precip, sfrac, time_s);
    w.snWE = w.iceWE + w.liqWE;
    w.snSub *= naughttocm; w.snEvap *= naughttocm;
    w.ETAge = 0.0;
    w.L = snow_latent_hf(a, w, snow_res_fact(a, lc, w));
    w.Prec = snow_precip_hf(a, rain, sfrac, snUnload);
    if (kStore && shortrad_from_met) EF(ShortRadIn) = a.inShortR;
    if ((w.snWE <= 5e-6) || (w.snTempC < -800.0)) {
      w.liqRoute = 0.0; w.snWE = 0.0; w.iceWE = 0.0; w.liqWE = 0.0; w.Utot = 0.0; w.snTempC = 0.0; w.crustAge = 0.0; w.densityAge = 0.0;
    } else {
      const double Utotold = w.iceWE * rhoicekg * cpicekJ * w.snTempC + w.liqWE * rholiqkg * latFreezekJ;
      w.Utot = Utotold;
      w.albedo = snow_aging_albedo(w);
      pr.alb_writer = 1; pr.alb_value = w.albedo;
      if (!kStore) return;
      // ---- snowEB
      const double resFact = snow_res_fact(a, lc, w);
      w.H = snow_sensible_hf(a, w, resFact);
      w.L = snow_latent_hf(a, w, resFact);
      w.Prec = w.phfOnOff * snow_precip_hf(a, rain, sfrac, snUnload);
      double Isw = 0.0;
      if (s.alphaD > 0.0) {                               // inShortWaveSn
        const double pi = 4.0 * atan(1.0), h0 = s.alphaD, pp0 = exp(-st.z[i] / 8434.5);
        const double Dh0ref = 0.061359 * (0.1594 + 1.123 * h0 + 0.065656 * h0 * h0) / (1 + 28.9344 * h0 + 277.3971 * h0 * h0);
        const double h0ref = h0 + Dh0ref;
        const double m = pp0 / (sin(h0ref * pi / 180.) + 0.50572 * pow((h0ref + 6.07995), -1.6364));
        const double drm = (m <= 20.0) ? 1 / (6.6296 + 1.7513 * m - 0.1202 * m * m + 0.0065 * (m * m * m) - 0.00013 * ((m * m) * (m * m)))
                                       : 1 / (10.4 + 0.718 * m);
        double Ic = s.Io * exp(-0.8662 * k.tlinke * m * drm);
        double Id = s.Io * k.tn_tlk * (k.a1 + k.a2 * s.sinAlpha + k.a3 * (s.sinAlpha * s.sinAlpha));
        if (s.ts_option > 1) { const double clr = a.inShortR; Ic = Ic / (Ic * s.sinAlpha + Id) * clr; Id = clr - Ic * s.sinAlpha; }
        if (k.shelter_option < 4) {
          const double cosi = 1.0 * (cos_s * s.sinAlpha + sin_s * s.cos_alpha * cos(s.sunaz - st.aspect[i]));
          if (cosi >= 0.0) Ics = Ic * cosi;
        } else Ics = Ic;
        if (k.shelter_option == 1 || k.shelter_option == 2) Ics *= above_horizon(st.hor_angle_0000[i], s.alphaD);
        const double v = diffuse_view(k, shelter, cos_s);
        const double Ids = Id * v;
        double Is = (Ics + Ids), Ir = 0.0;
        double hillalbedo = w.albedo;                     // HILLALBOPT 0
        if (sc.hill_albedo_option == 1) hillalbedo = lc.Al;
        else if (sc.hill_albedo_option == 2 && snCanWE == 0) hillalbedo = lc.V * lc.Al + (1 - lc.V) * w.albedo;
        if (k.shelter_option == 0) { Ir = hillalbedo * Is * (1 - cos_s) * 0.5; Is += Ir; }
        else if (k.shelter_option > 1 && k.shelter_option < 4) { Ir = hillalbedo * Is * (0.5 * (1 + cos_s) - shelter); Is += Ir; }
        const double Iv = Is * exp((lc.Kt - 1) * lc.LAI) * lc.V + Is * (1.0 - lc.V);
        Isw = Iv * (1.0 - w.albedo);
        IdsIr = Ids + Ir;
      }
      shortIn = (s.ts_option > 1) ? a.inShortR : Isw;
      EF(ShortRadIn) = shortIn; EF(ShortRadIn_dir) = Ics; EF(ShortRadIn_dif) = IdsIr;
      w.RSin = naughttokilo * Isw;
      {                                                  // inLongWave
        const double v0 = (k.shelter_option < 3) ? shelter : 1.0, N = a.skyCover / 10.0, kCloud = 1.0 + 0.17 * (N * N);
        const double Ea = 0.74 + 0.0049 * a.vPressC, tk = a.airTemp + 273.15, tk2 = tk * tk;
        w.RLin = naughttokilo * (v0 * kCloud * Ea * 5.67E-8 * (tk2 * tk2));
      }
      {                                                  // snowEB: long wave out, scaled by the sky view under OPTRADSHELT 1-2
        const double v1 = (k.shelter_option > 0 && k.shelter_option < 3) ? shelter : 1.0;
        const double tk = w.snTempC + 273.15, tk2 = tk * tk;
        w.RLout = -naughttokilo * v1 * 0.9 * 5.67e-8 * (tk2 * tk2);
      }
      EF(HFlux) = 0.0; EF(LFlux) = 0.0; EF(LongRadIn) = 0.0; EF(LongRadOut) = 0.0; EF(ShortAbsbVeg) = 0.0; EF(ShortAbsbSoi) = 0.0;
      SNF(SnLHF) = w.L * kilotonaught; SNF(SnSHF) = w.H * kilotonaught; SNF(SnPHF) = w.Prec * kilotonaught;
      SNF(SnGHF) = 0.0 * kilotonaught; SNF(SnRLin) = w.RLin * kilotonaught; SNF(SnRLout) = w.RLout * kilotonaught;
      SNF(SnRSin) = w.RSin * kilotonaught;
      const double Rn = w.RSin + w.RLin + w.RLout;
      w.G = 0;
      w.dUint = (w.H + Rn + w.L + w.Prec + w.G) * time_s;
      w.Utot += w.dUint;
      w.Uerr = (w.Utot - Utotold) - w.dUint;
      if (w.Utot < 0.0) {                                 // frozen pack: temperature follows the energy
        w.liqWE = 0.0; w.iceWE = w.snWE;
        double iceT = (w.iceWE < 0.1 && w.iceWE > 0) ? w.Utot / (cpicekJ * rhoicekg * 0.1) : w.Utot / (cpicekJ * rhoicekg * w.snWE);
        if (iceT <= sc.min_sn_temp) iceT = sc.min_sn_temp;
        w.snTempC = iceT;
      } else {                                            // melt
        w.liqWE = w.Utot / (latFreezekJ * rholiqkg);
        if (w.liqWE >= w.snWE) w.liqWE = w.snWE;
        w.iceWE = w.snWE - w.liqWE;
        snow_route_excess(w, sc.snliqfrac);
        w.snTempC = 0.0;
      }
    }
  }
  if (!kStore) return;
  if (w.snWE <= 5e-6) {
    w.liqRoute += w.snWE;
    w.snWE = 0.0; w.liqWE = 0.0; w.iceWE = 0.0; w.crustAge = 0.0; w.densityAge = 0.0; w.Utot = 0.0;
  } else { w.crustAge += time_h; w.densityAge += time_h; w.snOnOff = 1.0; }
  w.snWE = naughttocm * w.snWE; w.iceWE = naughttocm * w.iceWE; w.liqWE = naughttocm * w.liqWE; w.liqRoute = naughttocm * w.liqRoute;
  dv.f[TRIBS_F_EvapSoil][i] = 0.0;
  EF(ActEvap) = 0.0;
  EF(EvapoTrans) = EF(EvapWetCanopy) + dv.f[TRIBS_F_EvapDryCanopy][i];
  snow_store(w, sc, sv, ev, dv, i, time_s);
}

// callSnowPack for one cell, then the canopy water balance of a met step (as in et_cell)
TRIBS_HD void snow_cell_step(const EtConst &k, const EtStatic &st, const EtStepDev &s, const SnowConst &sc, const EtView &ev,
                             const SnowView &sv, const DynView &dv, int i, double alb_in, double can_in, SnowProbe &pr) {
  snow_cell_t<true>(k, st, s, sc, ev, sv, dv, i, alb_in, can_in, pr);
  if (s.met_time) {
    const double del_i = EF(InterceptLoss) - EF(EvapWetCanopy);
    EF(CanopyStorVol) = EF(CanopyStorVol) + del_i * (k.metstep_min / 60.0) * st.varea[i] / 1000.0;
  }
}
#undef EF
#undef SNF

}  // namespace tribs
