// TEST SCAFFOLDING (never shipped, never linked into libtribs_gpu.so).
// Compiles the device physics headers (tribs_b200/csrc/physics.cuh, cell_update.cuh) with the
// host compiler so that, on a box without a GPU, the exact source the CUDA kernels run can be
// compared bit-for-bit with the oracle / the reference dumps. Cells are visited sequentially
// in sweep order, which is one valid schedule of the device sweep.
#include "../../tribs_b200/csrc/cell_update.cuh"
#include <vector>
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
  for (int i = 0; i < n; i++) sat_prepare_cell(s, d, i, elev.data(), tr.data());
  for (int i = 0; i < n; i++) {
    double gw = sat_gather<32>(s, d, i, elev.data(), tr.data(), overflow);
    SatSums sums;
    int st = sat_cell(s, d, i, gw, m, t, sums);
    if (state_counts) state_counts[10 + st]++;
    glob[0] += sums.stok; glob[2] += sums.tot_gw; glob[3] += sums.tot_moist;
  }
}

double hc_lambert_w(double z) { return lambert_w(z); }
}
