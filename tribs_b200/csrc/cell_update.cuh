// cell_update.cuh — what one thread does for one Voronoi cell in each phase: load the SoA
// slices it needs, run the column physics (physics.cuh), write every reference-visible output.
//
// Reference call sites restated (src/tHydro/tHydroModel.cpp): SetupNodeUSZ 652-703, the
// UnSaturatedZone prologue 779-869 and epilogue 1995-2075, ComputeFluxesEdgesND 2663-2823,
// tCNode::getGwaterChng (src/tCNode/tCNode.cpp:356-370), SaturatedZone 3036-3503.
//
// Index space: "device order" = sweep order re-sorted by dependency level (see schedule.cpp);
// all neighbour indices stored on the device are already remapped, -1 = not an active cell.
#pragma once
#include "physics.cuh"
#include "../../include/tribs_gpu.h"

namespace tribs {

// static SoA, device order, length n_pad
struct StaticView {
  const int32_t *boundary;     // 0 hillslope, 3 stream, -1 padding
  const int32_t *owned;        // 1 = this rank computes the cell (all 1 on a single GPU)
  const int32_t *flow_dest;
  const int32_t *stream_node;
  const double *varea, *z, *bedrock;
  const double *cosv, *sinv;   // |cos|,|sin| of atan(flow slope) with the alpha>0 rule (787-789)
  const double *cos_gw;        // cos(atan(flow slope)) without the rule (2686-2689)
  const double *vel_mm;        // flow-edge Voronoi width * 1000
  const double *hillpath, *streampath, *contr_area;
  const double *ks, *ths, *thr, *lam, *psib, *f, *ar, *uar;
  // Voronoi neighbour CSR
  const int32_t *nbr_rowptr, *nbr_col, *nbr_listpos;
  const double *nbr_len, *nbr_vedglen, *nbr_len_in, *nbr_vedglen_in;
};

// dynamic SoA; Old/New roles are swapped by exchanging pointers on the host
struct DynView {
  double *f[TRIBS_N_FIELDS];
};

struct ModelConst {
  double int_storm_max, surface_depth, root_depth, bas_area, bas_area_flow;
  int32_t et_option, i_option, sn_opt, rdstr_option;
};

struct StepConst {
  double dt, dt_gw, now, elapsed;   // elapsed = (double)tRunTimer::getElapsedSteps(now)
  int32_t hourly_reset;
  double psrf_last;                 // psrf of the last swept cell (read by the saturated zone)
  // forcing applied inside the sweep (pipelined batches): 0 keep the Rain array, 1 uniform, 2 per cell,
  // 3 per rain gauge (rain_cells = the step's gauge values, rain_gauge = gauge of every cell)
  int32_t rain_mode = 0;
  double rain_uniform = 0.0;
  const double *rain_cells = nullptr;   // device order
  const int32_t *rain_gauge = nullptr;  // device order
  int32_t defer_esrf = 0;           // SAT: leave the raw satsrf in esrf, psrf_last is added later
};

TRIBS_HD void load_soil(Column &c, const StaticView &s, int i) {
  c.ks = s.ks[i]; c.ths = s.ths[i]; c.thr = s.thr[i]; c.lam = s.lam[i]; c.psib = s.psib[i];
  c.f = s.f[i]; c.ar = s.ar[i]; c.uar = s.uar[i];
  c.eps = 3 + 2 / c.lam;
  c.dbr = s.bedrock[i];
  c.memo_ok = 0; c.memo_nwt = 0.0; c.memo_val = 0.0;
}
TRIBS_HD void load_old(Column &c, const DynView &d, int i) {
  c.nwt0 = c.nwt1 = d.f[TRIBS_F_NwtOld][i];
  c.mu0 = c.mu1 = d.f[TRIBS_F_MuOld][i];
  c.mi0 = c.mi1 = d.f[TRIBS_F_MiOld][i];
  c.nt0 = c.nt1 = d.f[TRIBS_F_NtOld][i];
  c.nf0 = c.nf1 = d.f[TRIBS_F_NfOld][i];
  c.ru0 = c.ru1 = d.f[TRIBS_F_RuOld][i];
  c.ri0 = c.ri1 = d.f[TRIBS_F_RiOld][i];
}
TRIBS_HD void store_new(const Column &c, const DynView &d, int i) {
  d.f[TRIBS_F_NwtNew][i] = c.nwt1;
  d.f[TRIBS_F_MuNew][i] = c.mu1;
  d.f[TRIBS_F_MiNew][i] = c.mi1;
  d.f[TRIBS_F_NfNew][i] = c.nf1;
  d.f[TRIBS_F_NtNew][i] = c.nt1;
  d.f[TRIBS_F_RiNew][i] = c.ri1;
  d.f[TRIBS_F_RuNew][i] = c.ru1;
}

// soil-moisture diagnostics shared by both zones
TRIBS_HD void moisture_outputs(Column &c, const DynView &d, int i, const ModelConst &m, double elapsed) {
  double th = c.surf_soil_moist(m.surface_depth);
  double surf_sc = th / c.ths;
  d.f[TRIBS_F_SoilMoisture][i] = th;
  d.f[TRIBS_F_SoilMoistureSC][i] = surf_sc;
  d.f[TRIBS_F_SoilMoistureUNSC][i] = c.nwt1 ? c.mu1 / c.nwt1 / c.ths : 1.0;
  th = c.surf_soil_moist(m.root_depth);
  double root_sc = th / c.ths;
  d.f[TRIBS_F_RootMoisture][i] = th;
  d.f[TRIBS_F_RootMoistureSC][i] = root_sc;
  d.f[TRIBS_F_AvSoilMoisture][i] = av_soil_moist_update(d.f[TRIBS_F_AvSoilMoisture][i], surf_sc, root_sc, elapsed);
}

// per-cell partial sums of the basin accumulators (reduced deterministically by the caller)
struct UnsatSums { double stok, tot_rain, tot_gw, tot_moist; };

// ---------------------------------------------------------------------------------- phase UNSAT
// qpin_raw: sum (in sweep order) of the New Qpout of the upslope cells swept before this one.
// Returns the state id; writes Qpout[i] last so a chained consumer may read it afterwards.
struct UnsatCellInputs {
  Column c;
  UnsatEnv env;
  double varea, ractual, evsoi, evveg, qpout_prev_raw;
};

// everything that does not depend on this step's lateral inflow
TRIBS_HD void unsat_cell_load(UnsatCellInputs &u, const StaticView &s, const DynView &d, int i,
                              const ModelConst &m, const StepConst &t) {
  Column &c = u.c;
  load_soil(c, s, i);
  load_old(c, d, i);
  u.varea = s.varea[i];
  u.qpout_prev_raw = d.f[TRIBS_F_Qpout][i];
  c.isv = d.f[TRIBS_F_intstorm][i];
  c.hsrf = c.psrf = c.sbsrf = c.satsrf = 0.0;
  c.g = c.thri = c.thre = 0.0;
  c.cosv = s.cosv[i];
  c.sinv = s.sinv[i];
  u.env.dt = t.dt;
  u.env.int_storm_max = m.int_storm_max;
  u.env.rdstr_option = m.rdstr_option;
  u.env.vel_mm = s.vel_mm[i];
  int dst = s.flow_dest[i];
  // the outlet node is never swept: its Old values stay at their constructor zeros
  u.env.dest_nwt_old = dst >= 0 ? d.f[TRIBS_F_NwtOld][dst] : 0.0;
  u.env.dest_nf_old = dst >= 0 ? d.f[TRIBS_F_NfOld][dst] : 0.0;

  double rain_i;
  if (t.rain_mode == 0) rain_i = d.f[TRIBS_F_Rain][i];
  else {
    rain_i = (t.rain_mode == 1) ? t.rain_uniform : ((t.rain_mode == 2) ? t.rain_cells[i] : t.rain_cells[t.rain_gauge[i]]);
    d.f[TRIBS_F_Rain][i] = rain_i;   // tRainfall::NewRain
  }
  double evsoi = d.f[TRIBS_F_EvapSoil][i], evveg = d.f[TRIBS_F_EvapDryCanopy][i];
  if (evsoi < 0.0) evsoi = 0.0;
  if (evveg < 0.0) evveg = 0.0;
  u.evsoi = evsoi; u.evveg = evveg;
  double ractual;
  if (m.i_option == 0 && m.et_option == 0) ractual = rain_i;
  else if (m.i_option == 0 && m.et_option != 0) ractual = rain_i - evsoi - evveg;
  else if (m.i_option != 0 && m.et_option == 0) {
    if (m.i_option != 1) ractual = rain_i;
    else ractual = d.f[TRIBS_F_NetPrecipitation][i];
  } else
    ractual = d.f[TRIBS_F_NetPrecipitation][i] - evsoi - evveg;
  u.ractual = ractual;   // TotRain uses this value, before the snow override (828 vs 831-838)
  if (m.sn_opt) {
    double snwe = d.f[TRIBS_F_LiqWE][i] + d.f[TRIBS_F_IceWE][i];
    double routewe = d.f[TRIBS_F_LiqRouted][i];
    if ((snwe > 1e-3) || (routewe > 0.)) ractual = 10.0 * routewe - evveg;
  }
  c.r = ractual;
}

// the part that needs the lateral inflow; returns the New Qpout (mm^3/h)
TRIBS_HD double unsat_cell_advance(UnsatCellInputs &u, double qpin_raw, int *state_out) {
  Column &c = u.c;
  c.qpout = u.qpout_prev_raw * 1.0E-6 / u.varea;
  c.qpin = qpin_raw * 1.0E-6 / u.varea;
  c.r1 = (c.r + (c.qpin - c.qpout)) * c.cosv + 0.0;
  *state_out = unsat_advance(c, u.env);
  return c.qpout;
}

TRIBS_HD void unsat_cell_store(UnsatCellInputs &u, const DynView &d, int i, double qpin_raw,
                               const ModelConst &m, const StepConst &t, UnsatSums &sums) {
  Column &c = u.c;
  const double dt = t.dt;
  double esrf = c.psrf;
  esrf = esrf / c.cosv;
  double srf = c.hsrf + c.sbsrf + c.psrf;
  srf /= c.cosv;
  double hsrf = c.hsrf / c.cosv, sbsrf = c.sbsrf / c.cosv, psrf = c.psrf / c.cosv;

  store_new(c, d, i);
  d.f[TRIBS_F_Qpin][i] = qpin_raw;
  d.f[TRIBS_F_intstorm][i] = c.isv;
  double srf_hr = t.hourly_reset ? 0.0 : d.f[TRIBS_F_srf_hr][i];
  d.f[TRIBS_F_srf_hr][i] = srf_hr + srf * dt;
  d.f[TRIBS_F_srf][i] = srf * dt;
  d.f[TRIBS_F_cumsrf][i] += srf * dt;
  d.f[TRIBS_F_sbsrf][i] = sbsrf * dt;
  d.f[TRIBS_F_hsrf][i] = hsrf * dt;
  d.f[TRIBS_F_psrf][i] = psrf * dt;
  d.f[TRIBS_F_esrf][i] = esrf * dt;

  moisture_outputs(c, d, i, m, t.elapsed);

  d.f[TRIBS_F_Recharge][i] = (c.nwt1 - c.nwt0) * c.ths / (c.cosv * dt);
  d.f[TRIBS_F_UnSatFlowOut][i] = c.qpout * 1.0E-6 / u.varea;
  d.f[TRIBS_F_UnSatFlowIn][i] = c.qpin * 1.0E-6 / u.varea;
  if (hsrf > 0.0) d.f[TRIBS_F_hsrfOccur][i] = d.f[TRIBS_F_hsrfOccur][i] + floor(hsrf * 1.0E+3) + 1.0E-6;
  if (sbsrf > 0.0) d.f[TRIBS_F_sbsrfOccur][i] = d.f[TRIBS_F_sbsrfOccur][i] + floor(sbsrf * 1.0E+3) + 1.0E-6;
  if (psrf > 0.0) d.f[TRIBS_F_psrfOccur][i] = d.f[TRIBS_F_psrfOccur][i] + floor(psrf * 1.0E+3) + 1.0E-6;
  if (c.surf_soil_moist(1.0) / c.ths > 0.999) d.f[TRIBS_F_satOccur][i] = d.f[TRIBS_F_satOccur][i] + 1;
  d.f[TRIBS_F_RechDisch][i] = d.f[TRIBS_F_RechDisch][i] + ((c.nwt0 - c.nwt1) + sbsrf * dt * c.cosv / c.ths) * 1.0E-3;

  sums.stok = srf * dt * u.varea / 1000.0;
  sums.tot_rain = (u.ractual * dt * u.varea / 1000.);
  sums.tot_gw = (c.nwt1 - c.nwt0) * c.ths * u.varea / (c.cosv * 1000.0);
  sums.tot_moist = (c.mu1 - c.mu0) * u.varea / (c.cosv * 1000.0);
  c.psrf = psrf;  // the value the saturated zone later sees for the last swept cell
  d.f[TRIBS_F_Qpout][i] = c.qpout;
}

// ---------------------------------------------------------------------------------- phase SAT
// pre-pass: water-table elevation and finite-depth transmissivity of every cell (2686-2717, 2758)
TRIBS_HD void sat_prepare_cell(const StaticView &s, const double *nwt_old, int i, double *wt_elev, double *transm) {
  double nwt = nwt_old[i];
  double cs = s.cos_gw[i];
  if (nwt == 0.0) wt_elev[i] = s.z[i] - ((-s.psib[i]) / (cs * 1000.0));
  else wt_elev[i] = s.z[i] - (nwt / (cs * 1000.0));
  double f = s.f[i];
  transm[i] = (s.ar[i] * s.ks[i] * (m_exp(-f * nwt) - m_exp(-f * s.bedrock[i])) / f);
}

template <int MAXC>
TRIBS_HD double sat_gather(const StaticView &s, const DynView &d, int i, const double *wt_elev,
                           const double *transm, int *overflow) {
  // The reference sorts a cell's edge fluxes, then sums them (tCNode::getGwaterChng). The sorted list lives in
  // registers: every slot is addressed statically (a bubble pass of min / max per new flux keeps it ascending, unused
  // slots hold +inf), so nothing is indexed dynamically and nothing goes to local memory.
  double q[MAXC];
#pragma unroll
  for (int k = 0; k < MAXC; k++) q[k] = 1.0e308 * 10.0;   // +inf
  int nq = 0;
  const double elev_i = wt_elev[i], nwt_i = d.f[TRIBS_F_NwtOld][i], dbr_i = s.bedrock[i];
  const double t_i = transm[i], area_i = s.varea[i];
  int best = -1;
  bool best_pos = false;
  for (int k = s.nbr_rowptr[i]; k < s.nbr_rowptr[i + 1]; k++) {
    int j = s.nbr_col[k];
    if (j < 0) continue;  // outlet and closed-boundary nodes take no part (2681-2684)
    double elev_j = wt_elev[j], nwt_j = d.f[TRIBS_F_NwtOld][j];
    double flux = 0.0;
    bool out_pos = (elev_i > elev_j && nwt_i <= dbr_i && nwt_j <= dbr_i);
    if (out_pos) {  // i -> j, origin i
      double wts = (elev_i - elev_j) / s.nbr_len[k];
      double width = s.nbr_vedglen[k] * 1000.0;
      if (wts > 1.0) wts = 1.0;
      flux = t_i * width * wts;
      if (width) {
        double shape = s.varea[j] / (width * width * 10.0E-6);
        if (shape <= 0.1) flux *= shape;
      }
    } else {
      double dbr_j = s.bedrock[j];
      if (elev_j > elev_i && nwt_j <= dbr_j && nwt_i <= dbr_j) {  // j -> i, origin j
        double wts = (elev_j - elev_i) / s.nbr_len_in[k];
        double width = s.nbr_vedglen_in[k] * 1000.0;
        if (wts > 1.0) wts = 1.0;
        double qo = transm[j] * width * wts;
        if (width) {
          double shape = area_i / (width * width * 10.0E-6);
          if (shape <= 0.1) qo *= shape;
        }
        flux = -qo;
      }
    }
    if (flux > 0.0 || flux < 0.0) {
      if (nq < MAXC) {
        nq++;
        double x = flux;
#pragma unroll
        for (int k = 0; k < MAXC; k++) {     // ascending: the smaller of (slot, carried) stays, the larger moves on
          const double lo = fmin(q[k], x);
          x = fmax(q[k], x);
          q[k] = lo;
        }
      } else
        *overflow = 1;
    }
    int lp = s.nbr_listpos[k];
    if (lp > best) { best = lp; best_pos = out_pos; }
  }
  double gw = 0.0;
#pragma unroll
  for (int k = 0; k < MAXC; k++)
    if (k < nq) gw += q[k];
  if (best >= 0) {
    // the last edge of the cell in edge-list order decides which transmissivity is kept (2807, 2821)
    d.f[TRIBS_F_Transmissivity][i] = best_pos ? t_i : s.ar[i] * s.ks[i] * m_exp(-s.f[i] * nwt_i) / s.f[i];
  }
  return gw;
}

struct SatSums { double stok, tot_gw, tot_moist; };

TRIBS_HD int sat_cell(const StaticView &s, const DynView &d, int i, double gw, const ModelConst &m,
                      const StepConst &t, SatSums &sums) {
  Column c;
  load_soil(c, s, i);
  load_old(c, d, i);
  c.hsrf = c.psrf = c.sbsrf = c.satsrf = 0.0;
  c.g = c.thri = c.thre = 0.0;
  c.qpin = c.qpout = c.isv = c.r = c.r1 = 0.0;
  c.cosv = s.cosv[i];
  c.sinv = s.sinv[i];
  const double area = s.varea[i], dtgw = t.dt_gw;
  int st = sat_advance(c, gw, area, dtgw);

  double esrf = t.psrf_last + c.satsrf;  // stale psrf member of the unsaturated sweep (3439)
  esrf = esrf / c.cosv;
  const double satsrf_raw = c.satsrf;
  double satsrf = c.satsrf / c.cosv;
  double srf = satsrf;

  store_new(c, d, i);
  d.f[TRIBS_F_gwchange][i] = gw;
  d.f[TRIBS_F_srf_hr][i] += srf * dtgw;
  double srf_tot = d.f[TRIBS_F_srf][i] + srf * dtgw;
  d.f[TRIBS_F_srf][i] = srf_tot;
  d.f[TRIBS_F_cumsrf][i] += srf_tot;  // reference double-counts here (3460-3461)
  d.f[TRIBS_F_satsrf][i] = satsrf * dtgw;
  d.f[TRIBS_F_esrf][i] = t.defer_esrf ? satsrf_raw : esrf * dtgw;
  if (satsrf > 0.0) d.f[TRIBS_F_satsrfOccur][i] = d.f[TRIBS_F_satsrfOccur][i] + floor(satsrf * 1.0E+3) + 1.0E-6;
  if (c.surf_soil_moist(1.0) / c.ths > 0.999) d.f[TRIBS_F_satOccur][i] = d.f[TRIBS_F_satOccur][i] + 1;
  d.f[TRIBS_F_RechDisch][i] = d.f[TRIBS_F_RechDisch][i] + ((c.nwt0 - c.nwt1) + satsrf * dtgw * c.cosv / c.ths) * 1.0E-3;
  moisture_outputs(c, d, i, m, t.elapsed);
  d.f[TRIBS_F_Recharge][i] = (c.nwt1 - c.nwt0) * c.ths / (c.cosv * dtgw);
  sums.stok = srf * dtgw * area / 1000.0;
  sums.tot_moist = (c.mu1 - c.mu0) * area / (c.cosv * 1000.0);
  sums.tot_gw = (c.nwt1 - c.nwt0) * c.ths * area / (c.cosv * 1000.0);
  return st;
}

}  // namespace tribs
