/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * tribs_oracle_et: plain-C, single-thread CPU restatement of tRIBS's per-cell potential
 * evaporation / surface energy balance, ET partition and canopy interception
 * (reference: src/tHydro/tEvapoTrans.cpp:562-831, 915-1197, 1224-1512, 1751-2046, 2281-2716,
 * 2790-3076, 3962-4010; src/tHydro/tIntercept.cpp:174-459, 468-563), on flattened arrays in the
 * reference's sweep order. Host-side pieces of the same classes (calendar, sun geometry,
 * station files) are mirrored in Python (oracle/pyoracle.py, tribs_b200/model.py).
 *
 * Parity pinning: validated against the reference itself (oracle/_ref/tribs_ref_harness, phase
 * 'm' dumps) — tests/test_oracle.py and the fixture tests/golden/met.npz.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this.
 */
#ifndef TRIBS_ORACLE_ET_H
#define TRIBS_ORACLE_ET_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* per-cell fields written by the ET / interception sweeps (tCNode members of the same names) */
#define ORC_ET_FIELDS(X) \
  X(PotEvap) X(ActEvap) X(SurfTemp) X(SoilTemp) X(NetRad) X(GFlux) X(HFlux) X(LFlux) \
  X(LongRadIn) X(LongRadOut) X(ShortRadIn) X(ShortRadIn_dir) X(ShortRadIn_dif) X(ShortRadSlope) \
  X(ShortAbsbVeg) X(ShortAbsbSoi) X(AirTemp) X(DewTemp) X(RelHumid) X(VapPressure) X(SkyCover) \
  X(WindSpeed) X(AirPressure) X(CumHrsSun) X(AvEvapFract) X(SheltFact) X(CumRSin) \
  X(EvapWetCanopy) X(EvapoTrans) X(CumTotEvap) X(CumBarEvap) X(AvET) X(InterceptLoss) \
  X(CanStorage) X(CumIntercept) X(StormLength) X(CanopyStorVol)

#define ORC_ET_ENUM(name) ORCET_##name,
enum { ORC_ET_FIELDS(ORC_ET_ENUM) ORC_N_ET_FIELDS };

typedef struct {
  int32_t n;
  const double *z, *flow_slope, *aspect, *ThetaS, *ThetaR, *VolHeatCond, *SoilHeatCap;
  const double *soil_cutoff, *root_cutoff, *bedrock;
  const int32_t *land_class;      /* [n] row of land_table */
  const double *land_table;       /* rows of 15: column k = tInvariant::getLandProp(k), k = 1..14 */
  int32_t et_option, i_option, gflux_option, shelter_option;
  double metstep_min;             /* tEvapoTrans::timeStep (METSTEP, minutes) */
  double etistep_h;               /* tRunTimer::getEtIStep() */
  double tlinke, temp_lapse;
  double max_interstorm;          /* tIntercept::maxInterStormPeriod = INTSTORMMAX/2 */
  const double *hor_angle_0000;   /* OPTRADSHELT 1-3: tCNode::getHorAngle0000() [rad] (tShelter); NULL otherwise */
} orc_et_basin;

/* one met record per station (or per cell for gridded data), values of one hour */
typedef struct {
  int32_t n_records;
  const int32_t *cell_record;     /* [n] record of each cell */
  const double *air_temp, *dew_temp, *rel_humid, *vap_press, *atm_press, *wind_speed, *sky_cover;
  const double *rad_global, *elev;
} orc_met;

typedef struct {
  double alphaD, sinAlpha, sunaz, Io;   /* tEvapoTrans::SetSunVariables */
  int32_t hour;                         /* currentTime[3] */
  int32_t first_call;                   /* hourlyTimeStep == 0 */
  int32_t dew_hum_flag, ts_option;
  int32_t sky_flag_from;                /* cells at sweep index >= this use the computed sky cover */
  double te_met, te_eti;                /* tRunTimer::getElapsedMETSteps / getElapsedETISteps */
} orc_et_step;

/* forcing/state arrays shared with the water-balance sweeps */
typedef struct {
  double *Rain, *NetPrecipitation, *EvapSoil, *EvapDryCanopy, *SoilMoisture, *RootMoisture;
} orc_et_shared;

/* tEvapoTrans::callEvapoPotential; returns 0, or the (1-based) cell whose options are unsupported */
int orc_et_potential(const orc_et_basin *b, const orc_et_step *s, const orc_met *m,
                     double *const *etf, const orc_et_shared *sh);
/* tEvapoTrans::callEvapoTrans(Intercept, flag) */
int orc_et_components(const orc_et_basin *b, const orc_et_step *s, const orc_met *m,
                      double *const *etf, const orc_et_shared *sh, int flag);
int orc_et_n_fields(void);

/* ---------------------------------------------------------------- snow (tSnowPack::callSnowPack)
 * Single-layer snow pack energy/mass balance with canopy snow interception
 * (src/tHydro/tSnowPack.cpp:249-757, 769-934, 944-2001). The reference keeps some working values in
 * class members that are NOT reset for every cell (`albedo` is read by computeSub before the
 * cell's own agingAlbedo; `Uerr` is written only when the energy balance ran): they carry over
 * from the previous cell in sweep order and from the previous call. `members` holds them between
 * calls: [0] albedo (start 0.88), [1] Uerr (start 0), [2] snCanWE (start 0). */
#define ORC_SNOW_FIELDS(X) \
  X(SnTempC) X(CrustAge) X(DensityAge) X(EvapoTransAge) X(SnSub) X(SnEvap) X(DU) X(Unode) X(Uerror) \
  X(SnLHF) X(SnSHF) X(SnGHF) X(SnPHF) X(SnRLout) X(SnRLin) X(SnRSin) \
  X(IntSWE) X(IntPrec) X(IntSnUnload) X(IntSub) X(CumIntSub) X(CumIntUnl) \
  X(CumLHF) X(CumMelt) X(CumSHF) X(CumSnSub) X(CumSnEvap) X(CumPHF) X(CumRLin) X(CumRLout) X(CumGHF) \
  X(CumUerror) X(CumHrsSnow) X(PersTimeMax) X(PersTime) X(PeakSWE) X(PeakSWETemp) X(InitPackTime) \
  X(InitPackTimeTemp) X(PeakPackTime)
#define ORC_SNOW_ENUM(name) ORCSN_##name,
enum { ORC_SNOW_FIELDS(ORC_SNOW_ENUM) ORC_N_SNOW_FIELDS };

typedef struct {
  double min_sn_temp, snliqfrac;   /* MINSNTEMP, SNLIQFRAC */
  int32_t hill_albedo_option;      /* HILLALBOPT */
  int32_t hourly_time_step;        /* tEvapoTrans::hourlyTimeStep of this call */
  double *LiqWE, *IceWE, *LiqRouted;   /* arrays shared with the water-balance sweeps */
} orc_snow_args;

int orc_snow_pack(const orc_et_basin *b, const orc_snow_args *sn, const orc_et_step *s, const orc_met *m,
                  double *const *etf, double *const *snf, const orc_et_shared *sh, int flag, double *members);
int orc_snow_n_fields(void);

#ifdef __cplusplus
}
#endif
#endif
