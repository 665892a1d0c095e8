// TEST SCAFFOLDING (never shipped, never linked into libtribs_gpu.so).
// Compiles the device physics headers (tribs_b200/csrc/physics.cuh, cell_update.cuh) with the
// host compiler so that, on a box without a GPU, the exact source the CUDA kernels run can be
// compared bit-for-bit with the oracle / the reference dumps. Cells are visited sequentially
// in sweep order, which is one valid schedule of the device sweep.
#include "../../tribs_b200/csrc/cell_update.cuh"
#include "../../tribs_b200/csrc/snow_update.cuh"
#include "../../tribs_b200/csrc/channel_solve.cuh"
#include <vector>
#include <algorithm>
#include <cstring>

using namespace tribs;

extern "C" {

// sf order: varea z bedrock cosv sinv cos_gw vel_mm hillpath streampath contr_area ks ths thr lam psib f ar uar
// si order: boundary flow_dest stream_node nbr_rowptr nbr_col nbr_listpos
// ef order: nbr_len nbr_vedglen nbr_len_in nbr_vedglen_in
static std::vector<int32_t> g_ones;
static StaticView make_static(double **sf, int32_t **si, double **ef) {
  StaticView s;
  s.owned = nullptr;   // single owner on the host: never read by the functions compiled here
  s.varea = sf[0]; s.z = sf[1]; s.bedrock = sf[2]; s.cosv = sf[3]; s.sinv = sf[4]; s.cos_gw = sf[5];
  s.vel_mm = sf[6]; s.hillpath = sf[7]; s.streampath = sf[8]; s.contr_area = sf[9];
  s.ks = sf[10]; s.ths = sf[11]; s.thr = sf[12]; s.lam = sf[13]; s.psib = sf[14]; s.f = sf[15];
  s.ar = sf[16]; s.uar = sf[17];
  s.boundary = si[0]; s.flow_dest = si[1]; s.stream_node = si[2];
  s.nbr_rowptr = si[3]; s.nbr_col = si[4]; s.nbr_listpos = si[5];
  s.nbr_len = ef[0]; s.nbr_vedglen = ef[1]; s.nbr_len_in = ef[2]; s.nbr_vedglen_in = ef[3];
  return s;
}

int hc_n_fields() { return TRIBS_N_FIELDS; }

// one unsaturated sweep in sweep order; glob[0..4] += Stok, TotRain, TotGWchange, TotMoist; glob[4]=psrf_last
void hc_unsat(int n, double **sf, int32_t **si, double **ef, double **fields, const double *mc,
              const int32_t *mi, double dt, double now, double elapsed, int hourly_reset,
              double *glob, long long *state_counts) {
  StaticView s = make_static(sf, si, ef);
  DynView d;
  for (int k = 0; k < TRIBS_N_FIELDS; k++) d.f[k] = fields[k];
  ModelConst m{mc[0], mc[1], mc[2], mc[3], mc[4], mi[0], mi[1], mi[2], mi[3]};
  StepConst t;
  t.dt = dt; t.dt_gw = 0.0; t.now = now; t.elapsed = elapsed; t.hourly_reset = hourly_reset; t.psrf_last = 0.0;
  std::vector<double> qpin(n, 0.0);
  for (int i = 0; i < n; i++) {
    UnsatCellInputs u;
    unsat_cell_load(u, s, d, i, m, t);
    int st;
    double qo = unsat_cell_advance(u, qpin[i], &st);
    UnsatSums sums;
    unsat_cell_store(u, d, i, qpin[i], m, t, sums);
    if (st >= 0 && state_counts) state_counts[st]++;
    int dst = s.flow_dest[i];
    if (dst >= 0) qpin[dst] += qo;
    glob[0] += sums.stok; glob[1] += sums.tot_rain; glob[2] += sums.tot_gw; glob[3] += sums.tot_moist;
    glob[4] = u.c.psrf;
  }
  // contributions that arrive after their target was swept stay visible until Reset()
  for (int i = 0; i < n; i++) d.f[TRIBS_F_Qpin][i] = qpin[i];
}

void hc_sat(int n, double **sf, int32_t **si, double **ef, double **fields, const double *mc,
            const int32_t *mi, double dtgw, double now, double elapsed, double psrf_last,
            double *glob, long long *state_counts, int *overflow) {
  StaticView s = make_static(sf, si, ef);
  DynView d;
  for (int k = 0; k < TRIBS_N_FIELDS; k++) d.f[k] = fields[k];
  ModelConst m{mc[0], mc[1], mc[2], mc[3], mc[4], mi[0], mi[1], mi[2], mi[3]};
  StepConst t;
  t.dt = 0.0; t.dt_gw = dtgw; t.now = now; t.elapsed = elapsed; t.hourly_reset = 0; t.psrf_last = psrf_last;
  std::vector<double> elev(n), tr(n);
  for (int i = 0; i < n; i++) sat_prepare_cell(s, d.f[TRIBS_F_NwtOld], i, elev.data(), tr.data());
  for (int i = 0; i < n; i++) {
    double gw = sat_gather<32>(s, d, i, elev.data(), tr.data(), overflow);
    SatSums sums;
    int st = sat_cell(s, d, i, gw, m, t, sums);
    if (state_counts) state_counts[10 + st]++;
    glob[0] += sums.stok; glob[2] += sums.tot_gw; glob[3] += sums.tot_moist;
  }
}

double hc_lambert_w(double z) { return lambert_w(z); }

// const_math.cuh (TRIBS_CONST_MATH): the same source the device compiles, IEEE operations only
void hc_cm_exp(const double *x, double *y, int n) { for (int i = 0; i < n; i++) y[i] = cm::exp_c(x[i]); }
void hc_cm_log(const double *x, double *y, int n) { for (int i = 0; i < n; i++) y[i] = cm::log_c(x[i]); }
void hc_cm_pow(const double *x, const double *e, double *y, int n) { for (int i = 0; i < n; i++) y[i] = cm::pow_c(x[i], e[i]); }

// ---- kernel (1) and the snow pack: one call of the cell loop, cells in sweep order ----------------
// ci: et_option i_option gflux_option shelter_option | cd: metstep_min etistep_h tlinke temp_lapse max_interstorm
// est: z flow_slope aspect heat_cond heat_cap soil_cutoff root_cutoff thr bedrock varea hor_angle_0000
// si: do_potential do_components intercept_flag hour first_call dew_hum_flag ts_option sky_flag_from met_time
// sd: alphaD sinAlpha sunaz Io te_met te_eti
// met_pot / met_cmp: air_temp dew_temp rel_humid vap_press atm_press wind_speed sky_cover rad_global elev
// snow: NULL, or {min_sn_temp, snliqfrac, hill_albedo_option, hourly, albedo member, snCanWE member, probe rounds}:
// run callSnowPack instead; entries 4..6 are updated
int hc_n_et_fields() { return TRIBS_N_ET_FIELDS; }
int hc_n_snow_fields() { return TRIBS_N_SNOW_FIELDS; }
void hc_et_step(int n, const int32_t *ci, const double *cd, double **est, const int32_t *land_class, const double *land_table,
                int n_classes, const int32_t *si, const double *sd, const int32_t *cell_record, double **met_pot,
                double **met_cmp, double **etf, double **fields, double *snow, double **snf) {
  EtConst k{};
  k.et_option = ci[0]; k.i_option = ci[1]; k.gflux_option = ci[2]; k.shelter_option = ci[3];
  k.metstep_min = cd[0]; k.etistep_h = cd[1]; k.tlinke = cd[2]; k.temp_lapse = cd[3]; k.max_interstorm = cd[4];
  et_derived_constants(k);
  std::vector<LandClass> cls(n_classes);
  for (int q = 0; q < n_classes; q++) cls[q] = make_land_class(land_table + (size_t)q * 15, k.et_option);
  std::vector<double> cs(n), sn(n);
  std::vector<int32_t> bnd(n, 0), order(n);
  for (int i = 0; i < n; i++) { const double sl = fabs(atan(est[1][i])); cs[i] = cos(sl); sn[i] = sin(sl); order[i] = i; }
  EtStatic st{};
  st.z = est[0]; st.cos_slope = cs.data(); st.sin_slope = sn.data(); st.aspect = est[2]; st.heat_cond = est[3];
  st.heat_cap = est[4]; st.soil_cutoff = est[5]; st.root_cutoff = est[6]; st.thr = est[7]; st.bedrock = est[8]; st.varea = est[9];
  st.hor_angle_0000 = est[10];
  st.land_class = land_class; st.boundary = bnd.data(); st.sweep_of_dev = order.data(); st.classes = cls.data();
  static const double rsRatio[24] = {3.837, 3.589, 3.21, 2.43, 1.617, 1.196, 1.067, 1.014, 0.995, 0.976, 0.976, 1.0,
                                     1.053, 1.167, 1.354, 1.637, 2.043, 2.66, 3.215, 3.507, 3.689, 3.818, 3.923, 4.024};
  EtStepDev s{};
  s.do_potential = si[0] && k.et_option; s.do_components = si[1]; s.intercept_flag = si[2]; s.hour = si[3]; s.first_call = si[4];
  s.dew_hum_flag = si[5]; s.ts_option = si[6]; s.sky_flag_from = si[7]; s.met_time = si[8];
  s.alphaD = sd[0]; s.sinAlpha = sd[1]; s.sunaz = sd[2]; s.Io = sd[3]; s.te_met = sd[4]; s.te_eti = sd[5];
  s.cos_alpha = cos(asin(s.sinAlpha));
  s.rs_ratio = rsRatio[s.hour];
  auto met = [&](double **m) {
    MetDev d{};
    d.cell_record = cell_record;
    d.air_temp = m[0]; d.dew_temp = m[1]; d.rel_humid = m[2]; d.vap_press = m[3]; d.atm_press = m[4];
    d.wind_speed = m[5]; d.sky_cover = m[6]; d.rad_global = m[7]; d.elev = m[8];
    return d;
  };
  s.pot = met(met_pot);
  s.cmp = met(met_cmp ? met_cmp : met_pot);
  EtView ev;
  for (int f = 0; f < TRIBS_N_ET_FIELDS; f++) ev.f[f] = etf[f];
  DynView dv;
  for (int f = 0; f < TRIBS_N_FIELDS; f++) dv.f[f] = fields[f];
  if (!snow) {
    for (int i = 0; i < n; i++) et_cell(k, st, s, ev, dv, i);
    return;
  }
  SnowConst sc{snow[0], snow[1], (int32_t)snow[2], (int32_t)snow[3]};
  SnowView sv;
  for (int f = 0; f < TRIBS_N_SNOW_FIELDS; f++) sv.f[f] = snf[f];
  s.do_potential = 1; s.do_components = 1; s.intercept_flag = k.i_option != 0; s.met_time = 1; s.cmp = s.pot;
  SnowProbe pr;
  if (k.i_option == 0) {
    for (int i = 0; i < n; i++) snow_cell_step(k, st, s, sc, ev, sv, dv, i, 0.88, 0.0, pr);
    return;
  }
  // the protocol of snow_step_with_canopy (tribs_gpu.cu) with host loops for the kernels and the scan:
  // store-free probe rounds until the answers repeat, last-writer scans, one commit
  std::vector<int32_t> key_alb[2], key_can[2], last_alb(n + 1), last_can(n + 1), sensitive(n);
  std::vector<double> val_alb[2], val_can[2];
  for (int b = 0; b < 2; b++) { key_alb[b].assign(n + 1, 0); key_can[b].assign(n + 1, 0); val_alb[b].assign(n, 0.0); val_can[b].assign(n, 0.0); }
  double *members = snow + 4;
  auto scan = [&](const std::vector<int32_t> &key, std::vector<int32_t> &last) {
    int32_t run = 0;
    for (int p = 0; p <= n; p++) { last[p] = run; run = std::max(run, key[p]); }
  };
  auto carried = [&](const std::vector<int32_t> &last, const std::vector<double> &val, double member, int pos) {
    return last[pos] > 0 ? val[last[pos] - 1] : member;
  };
  int n_sensitive = 0;
  for (int i = 0; i < n; i++) {
    snow_cell_t<false>(k, st, s, sc, ev, sv, dv, i, members[0], 0.0, pr);
    sensitive[i] = pr.sensitive; n_sensitive += pr.sensitive;
    key_alb[0][i] = pr.alb_writer ? i + 1 : 0; val_alb[0][i] = pr.alb_value;
    key_can[0][i] = pr.can_writer ? i + 1 : 0; val_can[0][i] = pr.can_value;
  }
  scan(key_alb[0], last_alb);
  int buf = 0, rounds = 1;
  while (n_sensitive > 0) {
    const int rd = buf, wr = buf ^ 1;
    bool changed = false;
    for (int i = 0; i < n; i++) {
      if (!sensitive[i]) {
        key_alb[wr][i] = key_alb[rd][i]; val_alb[wr][i] = val_alb[rd][i]; key_can[wr][i] = key_can[rd][i]; val_can[wr][i] = val_can[rd][i];
        continue;
      }
      snow_cell_t<false>(k, st, s, sc, ev, sv, dv, i, carried(last_alb, val_alb[rd], members[0], i), 0.0, pr);
      const int32_t ka = pr.alb_writer ? i + 1 : 0, kc = pr.can_writer ? i + 1 : 0;
      if (ka != key_alb[rd][i] || kc != key_can[rd][i] || memcmp(&pr.alb_value, &val_alb[rd][i], 8) || memcmp(&pr.can_value, &val_can[rd][i], 8))
        changed = true;
      key_alb[wr][i] = ka; val_alb[wr][i] = pr.alb_value; key_can[wr][i] = kc; val_can[wr][i] = pr.can_value;
    }
    rounds++;
    buf = wr;
    if (!changed) break;
    scan(key_alb[buf], last_alb);
  }
  scan(key_can[buf], last_can);
  for (int i = 0; i < n; i++)
    snow_cell_step(k, st, s, sc, ev, sv, dv, i, carried(last_alb, val_alb[buf], members[0], i),
                   carried(last_can, val_can[buf], members[1], i), pr);
  if (last_alb[n] > 0) members[0] = val_alb[buf][last_alb[n] - 1];
  if (last_can[n] > 0) members[1] = val_can[buf][last_can[n] - 1];
  snow[6] = rounds;
}

// Channel router: the two linear solvers of channel_solve.cuh on one tridiagonal system (0-based sub / dia / sup, rhs f).
// x_ref: the reference's order (pivoted band elimination, one thread); x_part: the partition form, lanes looped.
void hc_chan_solve(int N, const double *sub, const double *dia, const double *sup, const double *f, double *x_ref, double *x_part, int *parts) {
  std::vector<double> A1(N + 2), A2(N + 2), A3(N + 2), B(N + 2), V(N + 2), W(N + 2);
  for (int i = 0; i < N; i++) {
    const int k = i + 1;
    if (k == 1) { A1[1] = dia[0]; A2[1] = sup[0]; A3[1] = 0.0; }
    else { A1[k] = sub[i]; A2[k] = dia[i]; A3[k] = (k == N) ? 999. : sup[i]; }
    B[k] = f[i];
  }
  chan_band_eliminate(A1.data(), A2.data(), A3.data(), B.data(), N);
  for (int k = 1; k <= N; k++) A1[k] = 1.0 / A1[k];
  chan_band_back(A1.data(), A2.data(), A3.data(), B.data(), N);
  for (int i = 0; i < N; i++) x_ref[i] = B[i + 1];
  const int P = chan_part_count(N);
  *parts = P;
  for (int i = 0; i < N; i++) B[i + 1] = f[i];
  for (int p = 0; p < P; p++) {
    int first, last;
    chan_part_range(N, P, p, &first, &last);
    chan_part_chunk(sub, dia, sup, A1.data(), A2.data(), A3.data(), B.data(), V.data(), W.data(), first, last, p > 0, p < P - 1);
  }
  std::vector<double> M((size_t)(2 * P) * (2 * P + 1)), X(2 * P);
  chan_part_interface(B.data(), V.data(), W.data(), N, P, M.data(), X.data());
  for (int i = 0; i < N; i++) x_part[i] = chan_part_assemble(B.data(), V.data(), W.data(), X.data(), N, P, i + 1);
}
}
