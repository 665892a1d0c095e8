/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * tribs_oracle: plain-C, single-thread CPU restatement of the tRIBS per-timestep
 * Voronoi-cell water balance (reference: src/tHydro/tHydroModel.cpp,
 * src/tFlowNet/tKinemat.cpp:920-1125, src/tFlowNet/tFlowNet.cpp:735-867,
 * src/tFlowNet/tFlowResults.cpp:857-1161, src/tCNode/tCNode.cpp:356-370).
 * It works on flattened arrays in the reference's sweep (list) order.
 *
 * Parity pinning: validated against the reference itself (oracle/_ref/tribs_ref_harness,
 * the unmodified reference sources compiled in place) — see tests/test_oracle_vs_ref.py and
 * the committed fixtures under tests/golden/ that the harness generated.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this.
 * The product (tribs_b200/csrc, libtribs_gpu.so) never links or calls it.
 */
#ifndef TRIBS_ORACLE_H
#define TRIBS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- static per-cell FP64 arrays (index into orc_basin.sf) ---- */
#define ORC_STATIC_F64(X) \
  X(varea) X(z) X(flow_slope) X(flow_vedglen) X(bedrock) X(hillpath) X(streampath) \
  X(contr_area) X(Ks) X(ThetaS) X(ThetaR) X(PoreSize) X(AirEBubPres) X(DecayF) \
  X(SatAnRatio) X(UnsatAnRatio)

/* ---- static per-cell int32 arrays (orc_basin.si) ---- */
#define ORC_STATIC_I32(X) X(boundary) X(flow_dest) X(stream_node)

/* ---- per-directed-edge arrays of the Voronoi neighbour CSR ---- */
#define ORC_EDGE_F64(X) X(nbr_len) X(nbr_vedglen) X(nbr_len_in) X(nbr_vedglen_in)
#define ORC_EDGE_I32(X) X(nbr_col) X(nbr_listpos)

/* ---- dynamic per-cell FP64 arrays (orc_state.f) ---- */
#define ORC_STATE_F64(X) \
  X(NwtOld) X(NwtNew) X(MuOld) X(MuNew) X(MiOld) X(MiNew) X(NtOld) X(NtNew) \
  X(NfOld) X(NfNew) X(RuOld) X(RuNew) X(RiOld) X(RiNew) \
  X(Qpin) X(Qpout) X(intstorm) X(gwchange) \
  X(srf) X(hsrf) X(sbsrf) X(psrf) X(satsrf) X(esrf) X(srf_hr) X(cumsrf) \
  X(Rain) X(NetPrecipitation) X(EvapSoil) X(EvapDryCanopy) \
  X(SoilMoisture) X(SoilMoistureSC) X(SoilMoistureUNSC) X(RootMoisture) X(RootMoistureSC) \
  X(AvSoilMoisture) X(Recharge) X(UnSatFlowIn) X(UnSatFlowOut) X(Transmissivity) \
  X(hsrfOccur) X(sbsrfOccur) X(psrfOccur) X(satsrfOccur) X(satOccur) X(RechDisch) \
  X(FlowVelocity) X(TTime) X(Qstrm) X(LiqWE) X(IceWE) X(LiqRouted) X(UnSaturatedStorage) X(SaturatedStorage)

/* ---- basin accumulators (orc_state.glob) ---- */
#define ORC_GLOBALS(X) \
  X(Stok) X(TotRain) X(TotGWchange) X(TotMoist) X(psrf_last) \
  X(fSoi100) X(fTop100) X(fClm100) X(fGW100) X(dM100) X(dMRt) X(mTh100) X(mThRt) \
  X(hillvel_last) X(Qstrm_outlet)

/* ---- tFlowResults arrays (orc_results.a[k][bin]) ---- */
#define ORC_RESULTS(X) \
  X(mhydro) X(HsrfRout) X(SbsrfRout) X(PsrfRout) X(SatsrfRout) X(crr) X(rmax) X(rmin) \
  X(frac) X(msm) X(msmRt) X(msmU) X(mgw) X(sat) X(qunsat)

#define ORC_ENUM(name) ORC_##name,
enum { ORC_STATIC_F64(ORC_ENUM) ORC_N_STATIC_F64 };
#define ORC_ENUMI(name) ORCI_##name,
enum { ORC_STATIC_I32(ORC_ENUMI) ORC_N_STATIC_I32 };
#define ORC_ENUME(name) ORCE_##name,
enum { ORC_EDGE_F64(ORC_ENUME) ORC_N_EDGE_F64 };
#define ORC_ENUMEI(name) ORCEI_##name,
enum { ORC_EDGE_I32(ORC_ENUMEI) ORC_N_EDGE_I32 };
#define ORC_ENUMS(name) ORCS_##name,
enum { ORC_STATE_F64(ORC_ENUMS) ORC_N_STATE_F64 };
#define ORC_ENUMG(name) ORCG_##name,
enum { ORC_GLOBALS(ORC_ENUMG) ORC_N_GLOBALS };
#define ORC_ENUMR(name) ORCR_##name,
enum { ORC_RESULTS(ORC_ENUMR) ORC_N_RESULTS };

typedef struct {
  int32_t n;                         /* active cells */
  int32_t n_edges;                   /* directed spokes in the CSR */
  double *sf[ORC_N_STATIC_F64];      /* each of length n */
  int32_t *si[ORC_N_STATIC_I32];
  int32_t *nbr_rowptr;               /* n+1 */
  double *ef[ORC_N_EDGE_F64];        /* each of length n_edges */
  int32_t *ei[ORC_N_EDGE_I32];
  /* options / scalars */
  double BasArea;                    /* tHydroModel::BasArea */
  double BasArea_flow;               /* tFlowNet::BasArea   */
  double IntStormMAX;
  double surfaceSoilDepth, rootZoneDepth;
  int32_t EToption, Ioption, SnOpt, RdstrOption;
  /* routing */
  double velcoef, velratio, flowexp, baseflow, kincoef, dtReff, hillvel0;
  int32_t outlet_node;               /* sweep index of OutletNode or -1 */
  double outlet_contr_area;
  /* timer */
  double tstep, gwstep, etistep, output_interval;
} orc_basin;

typedef struct {
  double *f[ORC_N_STATE_F64];        /* each of length n */
  double glob[ORC_N_GLOBALS];
} orc_state;

/* sparse (TimeInd, Qeff) stack of one stream cell, ascending TimeInd */
typedef struct {
  int32_t n, cap;
  int32_t *tind;
  double *q;
} orc_stack;

typedef struct {
  int32_t limit;                     /* bins per array */
  double *a[ORC_N_RESULTS];
  orc_stack *stacks;                 /* n+1 entries: stream cells by sweep index, [n] = the outlet node */
} orc_results;

/* one reference time step is: time += tstep; [forcing]; unsat; [resetgw; sat]; route; reset */
void orc_unsat(const orc_basin *b, orc_state *s, double dt, double current_time);
void orc_reset_gw(const orc_basin *b, orc_state *s);
void orc_sat(const orc_basin *b, orc_state *s, double dtgw, double current_time);
void orc_route(const orc_basin *b, orc_state *s, orc_results *r, double current_time);
void orc_reset(const orc_basin *b, orc_state *s);
/* tWaterBalance::UnSaturatedBalance / SaturatedBalance / CanopyBalance / BasinStorage as
 * Simulator::UpdateWaterBalance calls them (tSimul.cpp:576-585, tWaterBalance.cpp:116-345).
 * gw_time: fmod(currentTime, GWSTEP) == 0; met_time: currentTime == met_hour. The three ET arrays
 * and canopy_stor_vol may be NULL (ET off): treated as zeros / not updated. basin6 = BasinStorages. */
void orc_balance(const orc_basin *b, orc_state *s, int gw_time, int met_time, double metstep_h,
                 const double *intercept_loss, const double *evap_wet_canopy, const double *evapotrans,
                 const double *can_storage, double *canopy_stor_vol, double *basin6);
void orc_retrieve_qeff(const orc_basin *b, orc_results *r, double current_time, double *out);
void orc_state_counts(long long *out16, int reset); /* [0..9] unsat states, [10..13] GW states */

/* scalar helpers exposed for known-answer tests (soil = Ks,Ths,Thr,PoreInd,Psib,F,Ar,UAr,bedrock) */
double orc_kat_lambertw(double z);
double orc_kat_total_moist(const double *soil, double nwt);
double orc_kat_upper_moist(const double *soil, double nf, double nwt);
double orc_kat_lower_moist(const double *soil, double nf, double nwt);
double orc_kat_newton1(const double *soil, double dm, double nwt);
double orc_kat_newton2(const double *soil, double dm, double nwt, double nf, int flag);
double orc_kat_recharge(const double *soil, double dm, double nf);
double orc_kat_unsat_lat(const double *soil, double nf, double ru, double vel);
double orc_kat_sat_lat(const double *soil, double nt, double nf, double ru, double vel);
double orc_kat_z1z2(const double *soil, double z1, double z2, double nwt);

int orc_sizes(int which); /* 0..6: counts of the seven field tables above, for binding checks */

#ifdef __cplusplus
}
#endif
#endif
