/* qd_b200.h -- C ABI of libqd_b200: the B200-native Qingdai per-timestep loop.
 *
 * Drop-in boundary.  Every entry point replaces one piece of the
 * reference's Python hot path (paths relative to the reference checkout):
 *
 *   qd_laplacian / qd_hyperdiffuse / qd_advect   <- pygcm/jax_compat.py:111,135,190 (the backend
 *        seam imported at pygcm/dynamics.py:15, pygcm/ocean.py:24) == dynamics.py:90-212
 *   qd_shapiro / qd_zonal_bandstop               <- pygcm/dynamics.py:215-258, ocean.py:154-164
 *   qd_gaussian                                  <- scipy.ndimage.gaussian_filter call sites
 *                                                   pygcm/physics.py:44,69,111,159,330; run_simulation.py:1931
 *   qd_divergence / qd_vorticity                 <- pygcm/grid.py:41-88
 *   qd_median_pos / qd_wsum                      <- np.median(x[x>0]) (physics.py:300,
 *                                                   run_simulation.py:1873, dynamics.py:348); energy.py:524
 *   qd_atmos_step                                <- SpectralModel.time_step pygcm/dynamics.py:260-667
 *   qd_ocean_step                                <- WindDrivenSlabOcean.step pygcm/ocean.py:265-533
 *   qd_loop_step                                 <- loop body scripts/run_simulation.py:1760-2344
 *   qd_route_event                               <- RiverRouting.step pygcm/routing.py:241-331
 *   qd_wsum / qd_minmax                          <- energy.py:494-538, hydrology.py:270-340, ocean.py:535-561
 *
 * Conventions: plain pointers and sizes only.  All fields are C-contiguous float64
 * [batch][n_lat][n_lon] (row 0 = south pole, column 0 = 0E, last column = 360E duplicate);
 * masks are uint8.  "dev" pointers are device memory owned by the caller (PyTorch); the
 * library allocates only its private tables at qd_create.  Every function returns 0 on
 * success or a negative qd_status; qd_last_error() gives the message.  Work is enqueued on
 * the stream set by qd_set_stream (default: the legacy default stream) and is asynchronous
 * unless stated.  There is no CPU fallback: without a CUDA device qd_create fails.
 */
#ifndef QD_B200_H
#define QD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qd_ctx qd_ctx;

typedef enum {
  QD_OK = 0,
  QD_E_INVALID = -1,   /* bad argument */
  QD_E_CUDA = -2,      /* CUDA runtime error (see qd_last_error) */
  QD_E_NODEVICE = -3,  /* no CUDA device: the product path refuses to run */
  QD_E_UNBOUND = -4,   /* a field the call needs was never bound */
  QD_E_STATE = -5      /* call sequence error */
} qd_status;

/* ---- float64 field slots (index into the caller's [QD_F_COUNT][B][nlat][nlon] block) ---- */
typedef enum {
  /* atmosphere prognostics (dynamics.py:56-88) */
  QD_F_U = 0, QD_F_V, QD_F_H, QD_F_TS, QD_F_Q, QD_F_CLOUD, QD_F_HICE,
  /* atmosphere diagnostics kept by the reference as attributes */
  QD_F_ISR, QD_F_ISR_A, QD_F_ISR_B, QD_F_OLR, QD_F_EFLUX, QD_F_PCOND, QD_F_LH, QD_F_LHREL, QD_F_CLOUD_EFF,
  /* ocean prognostics (ocean.py:85-94) */
  QD_F_UO, QD_F_VO, QD_F_ETA, QD_F_SST,
  /* land reservoirs (run_simulation.py:1289-1290) */
  QD_F_WLAND, QD_F_SSNOW,
  /* static maps */
  QD_F_FRICTION, QD_F_BASE_ALBEDO, QD_F_ELEVATION, QD_F_CS_MAP, QD_F_OROG_NX, QD_F_OROG_NY,
  /* per-step products of the loop body */
  QD_F_PRECIP, QD_F_ALBEDO, QD_F_TEQ, QD_F_QNET, QD_F_CSNOW, QD_F_RLAND,
  /* ecology sub-daily (adapter.py:140-186) */
  QD_F_EDAY, QD_F_FCANOPY, QD_F_ALPHA_ECO, QD_F_LAI_SNAP,
  /* private scratch (ping-pong partners and stencil intermediates) */
  QD_F_X0, QD_F_X1, QD_F_X2, QD_F_X3, QD_F_X4, QD_F_X5, QD_F_X6, QD_F_X7, QD_F_X8, QD_F_X9, QD_F_X10,
  QD_F_COUNT
} qd_field_id;

/* ---- uint8 mask slots ---- */
typedef enum { QD_M_LAND = 0, QD_M_ICE, QD_M_GLACIER, QD_M_COUNT } qd_mask_id;

/* ---- per-row tables [QD_R_COUNT][nlat], filled by the host with the reference's NumPy
 *      expressions so that metric terms are bit-identical (grid.py:27-39,95; dynamics.py:104,
 *      164,490,516-518,557-570; ocean.py:82,332-334,344-347; routing.py:176-200) ---- */
typedef enum {
  QD_R_LAT_DEG = 0, QD_R_COS, QD_R_SIN, QD_R_FCOR, QD_R_W,
  QD_R_COS_ADV_ATM,   /* max(1e-6, cos)          */
  /* a cosine table used by the Laplacian is followed by its 1/c and 1/c^2 rows (the stencil kernels
   * multiply by reciprocals: fp64 division is ~20 instructions on sm_100 and the step is otherwise
   * fp64-issue bound; differences to true division are <= 1 ulp per operation) and by the three
   * centred-stencil coefficient rows ap, am, bl of the del^4 kernels; the library fills all five
   * companion rows itself (qd_create, qd_set_rows*, qd_user_row*) */
  QD_R_COS_ADV_HALF,  /* max(cos, 0.5)           */
  QD_R_ICOS_HALF, QD_R_ICOS2_HALF, QD_R_H4AP_HALF, QD_R_H4AM_HALF, QD_R_H4BL_HALF,
  QD_R_COS_LAP_ATM,   /* max(cos, 0.2)           */
  QD_R_ICOS_LAP_ATM, QD_R_ICOS2_LAP_ATM, QD_R_H4AP_LAP_ATM, QD_R_H4AM_LAP_ATM, QD_R_H4BL_LAP_ATM,
  QD_R_COS_CAP,       /* max(cos, 1e-6)          */
  QD_R_FSAFE,         /* regularised Coriolis    */
  QD_R_K4_U, QD_R_K4_V, QD_R_K4_H, QD_R_K4_Q, QD_R_K4_C,
  QD_R_OC_S4DX4,      /* sigma4 * dx_min^4 (ocean; divided by sub_dt on device) */
  QD_R_OC_SPONGE,     /* polar_gain * s^2        */
  QD_R_POLAR,         /* 1.0 where |lat| >= QD_POLAR_LAT_THRESH */
  QD_R_AREA,          /* cell area m^2 (routing) */
  QD_R_INV_ACOS_HALF, /* 1/(a*max(cos,0.5)): ocean pressure-gradient metric (ocean.py:309) */
  QD_R_INV_ACOS_CAP,  /* 1/(a*max(cos,1e-6)): divergence / vorticity metric (grid.py:66,87) */
  QD_R_IAC_ADV_ATM,   /* RN(1/(a*max(1e-6,cos))): exact-division helper of the atmospheric departure points (dynamics.py:104) */
  QD_R_COUNT
} qd_row_id;

/* ---- per-column tables [QD_C_COUNT][nlon] ---- */
typedef enum { QD_C_LON_RAD = 0, QD_C_SIN_LON, QD_C_COS_LON, QD_C_COUNT } qd_col_id;

/* ---- parameters: one float64 vector per ensemble member, indexed by qd_param_id
 *      (booleans/ints are stored as 0/1 or integral doubles; "unset" overrides are NaN).
 *      Names follow qingdai_b200/params.py:QDParams. ---- */
typedef enum {
  QD_P_G = 0, QD_P_TAU_RAD, QD_P_GH_NEWTON, QD_P_ENERGY_W, QD_P_MOM_PRIMITIVE,
  QD_P_SW_A0, QD_P_SW_KC, QD_P_LW_EPS0, QD_P_LW_KC, QD_P_T_FLOOR, QD_P_C_SFC,
  QD_P_CLOUD_COUPLE, QD_P_RH0, QD_P_K_Q, QD_P_K_P, QD_P_PCOND_REF,
  QD_P_LW_V2, QD_P_HICE_REF, QD_P_EPS_OCEAN, QD_P_EPS_LAND, QD_P_EPS_ICE,
  QD_P_LW_TAU0, QD_P_LW_KTAU, QD_P_GH_LOCK, QD_P_GH_FACTOR_LW, QD_P_C_H, QD_P_CP_AIR,
  QD_P_SEAICE, QD_P_T_FREEZE, QD_P_RHO_I, QD_P_L_F, QD_P_CS_OCEAN, QD_P_CS_LAND, QD_P_CS_ICE,
  QD_P_POLAR_FIX_S, QD_P_POLAR_FIX_N, QD_P_ATM_H, QD_P_DIFF_FACTOR,
  QD_P_C_E, QD_P_RHO_A, QD_P_H_MBL, QD_P_L_V, QD_P_P0,
  QD_P_EVAP_OCEAN, QD_P_EVAP_LAND, QD_P_EVAP_ICE, QD_P_TAU_COND,
  /* ocean */
  QD_P_OC_H, QD_P_OC_RHO_W, QD_P_OC_CP_W, QD_P_OC_G, QD_P_OC_CD, QD_P_OC_R_BOT, QD_P_OC_RHO_A,
  QD_P_OC_VCAP, QD_P_OC_TAU_SCALE, QD_P_OC_K_H, QD_P_OC_CFL, QD_P_OC_MAX_U, QD_P_OC_MEAN4,
  QD_P_OC_ADV_ALPHA, QD_P_OC_USE_QNET, QD_P_OC_ICE_QFAC, QD_P_OC_ETA_CAP, QD_P_OC_POLAR_FIX,
  QD_P_OC_TS_MIN, QD_P_OC_TS_MAX, QD_P_OC_K4_U, QD_P_OC_K4_V, QD_P_OC_K4_ETA, QD_P_OC_DX_MIN,
  /* loop physics */
  QD_P_D_CRIT, QD_P_K_PRECIP, QD_P_ALPHA_WATER, QD_P_ALPHA_ICE, QD_P_ALPHA_CLOUD,
  QD_P_USE_TOPO_ALBEDO, QD_P_OROG, QD_P_K_OROG, QD_P_BETA_DIV, QD_P_P_FALLBACK, QD_P_PQ_MIN,
  QD_P_P_BLEND, QD_P_PREF, QD_P_CMAX, QD_P_W_MEM, QD_P_W_P, QD_P_W_SRC, QD_P_CLOUD_FLOOR,
  QD_P_CLOUD_ADVECT, QD_P_CLOUD_ADV_ALPHA, QD_P_CLOUD_SMOOTH_SIGMA,
  QD_P_LAPSE_ENABLE, QD_P_LAPSE_KPM, QD_P_LAND_ELEV_MAX, QD_P_POLAR_ICE_THICK_MAX, QD_P_RHO_SNOW,
  QD_P_GLACIER_FRAC, QD_P_GLACIER_SWE, QD_P_HAS_ELEVATION,
  /* hydrology */
  QD_P_RUNOFF_TAU_DAYS, QD_P_WLAND_CAP, QD_P_SNOW_THRESH, QD_P_SNOW_MELT_RATE, QD_P_SNOW_T_BAND,
  QD_P_SNOW_DEGREE_DAY, QD_P_SNOW_DDF, QD_P_SNOW_MELT_TREF, QD_P_SWE_ENABLE, QD_P_SWE_REF,
  QD_P_SWE_MAX, QD_P_SNOW_ALBEDO_FRESH,
  /* ecology */
  QD_P_ECO_ENABLE, QD_P_ECO_W_LAI, QD_P_ECO_SOIL_REFLECT, QD_P_ECO_ALPHA_LEAF,
  /* host-evaluated sums (np.sum over the 2-D weight arrays, energy.py:522, ocean.py:374) */
  QD_P_WSUM_ALL, QD_P_OC_WSUM_OCEAN, QD_P_OC_ANY_OCEAN,
  /* host-evaluated reciprocals of parameter-only divisors */
  QD_P_OC_INV_RHO_H, QD_P_OC_INV_RHO_CP_H,
  /* smallest s2 = uo^2 + vo^2 whose correctly rounded square root exceeds QD_OCEAN_MAX_U: the per-cell test
   * sqrt(s2) > cap of ocean.py:412 becomes s2 >= threshold with identical outcomes and no square root on the common path */
  QD_P_OC_SPEED2_CAP,
  QD_P_COUNT
} qd_param_id;

/* ---- per-member device scalars written by reduction kernels (readable via qd_get_scalars) ---- */
typedef enum {
  QD_S_SUM_PQW = 0, QD_S_SUM_PRAWW, QD_S_MED_POS, QD_S_CNT_POS, QD_S_PREF, QD_S_CNT_PRECIP,
  QD_S_PREF_ATM, QD_S_CNT_PCOND, QD_S_MAX_UOCEAN, QD_S_MAX_VA, QD_S_ETA_NUM, QD_S_SUB_DT,
  QD_S_NSUB, QD_S_WSUM, QD_S_WSUM_OCEAN, QD_S_TMP0, QD_S_TMP1, QD_S_TMP2, QD_S_TMP3,
  /* ecology canopy-cache clock (population.py:57-71,272-276): accumulated hours, next recompute,
   * cache present, "recomputed in this step" flag */
  QD_S_ECO_HOURS, QD_S_ECO_NEXT, QD_S_ECO_CACHED, QD_S_ECO_FLAG,
  QD_S_COUNT
} qd_scalar_id;

/* ---- per-step forcing scalars (forcing.py:78-136, orbital.py:33-52), computed by the host with
 *      the reference's NumPy expressions ---- */
typedef struct {
  double t;                                  /* model time, s */
  double flux_a, sin_delta_a, cos_delta_a, alpha_a;
  double flux_b, sin_delta_b, cos_delta_b, alpha_b;
  double theta;                              /* (t*omega) mod 2pi */
} qd_forcing_t;

/* ---- step configuration: switches and cadences of one step.  The library keeps the two
 *      counters the reference keeps (SpectralModel._step_counter dynamics.py:451, pre-incremented;
 *      WindDrivenSlabOcean._step ocean.py:281) and applies the cadence tests itself. ---- */
typedef struct {
  double dt;
  int has_albedo;        /* time_step(..., albedo=array): energy branch live (dynamics.py:326) */
  int diff_enable;       /* QD_DIFF_ENABLE && filter in {hyper4, combo} (dynamics.py:542) */
  int diff_every, k4_nsub, apply_q, apply_cloud;
  int shapiro_every, shapiro_n, shapiro_q, shapiro_cloud;   /* 0 = off (dynamics.py:612-626) */
  int spec_every; double spec_cutoff, spec_damp;            /* 0 = off (dynamics.py:629-637) */
  int oc_diff_every, oc_k4_nsub, oc_shapiro_n, oc_shapiro_every;   /* ocean.py:341,359 */
  int oc_has_q, oc_has_ice;                                 /* Q_net / ice_mask arguments given */
  int with_ocean, with_hydrology, with_routing, with_eco;   /* loop composition */
  int loop_with_albedo;  /* opt-in deviation: the loop passes albedo into the atmosphere step */
  int store_isr_ab;      /* also store the per-star insolation fields */
} qd_step_cfg_t;

/* ------------------------------------------------------------------ lifecycle */
int  qd_create(int nlat, int nlon, int batch, int device,
               double a, double dlat, double dlon,
               double a_sq, double dlon_sq,   /* a**2 and dlon**2 as the host's Python evaluates them */
               const double* rows_host /* [QD_R_COUNT][nlat] */,
               const double* cols_host /* [QD_C_COUNT][nlon] */,
               const double* params_host /* [batch][QD_P_COUNT] */,
               qd_ctx** out);
int  qd_destroy(qd_ctx* ctx);
const char* qd_last_error(const qd_ctx* ctx);
int  qd_version(void);
int  qd_set_stream(qd_ctx* ctx, void* cuda_stream);
int  qd_synchronize(qd_ctx* ctx);
/* caller-owned device storage: fields [QD_F_COUNT][B][nlat][nlon] f64, masks [QD_M_COUNT][B][nlat][nlon] u8 */
int  qd_bind(qd_ctx* ctx, double* fields_dev, uint8_t* masks_dev);
int  qd_set_params(qd_ctx* ctx, const double* params_host);     /* re-snapshot (sync) */
int  qd_set_rows(qd_ctx* ctx, const double* rows_host);         /* e.g. K4 rows after a dt change; every member */
/* one member's row table: the K4, ocean sponge and polar rows follow that member's QD_* parameters
 * (dynamics.py:557-570, ocean.py:332-347, run_simulation.py:1956) -- ensemble parameter sweeps */
int  qd_set_rows_member(qd_ctx* ctx, int member, const double* rows_host);
int  qd_get_scalars(qd_ctx* ctx, double* out_host /* [batch][QD_S_COUNT] */);   /* sync */
/* host <-> device field transfer through the C ABI (sync); member b or -1 for all members */
int  qd_upload_field(qd_ctx* ctx, int field, int member, const double* host);
int  qd_download_field(qd_ctx* ctx, int field, int member, double* host);
int  qd_upload_mask(qd_ctx* ctx, int mask, int member, const uint8_t* host);
int  qd_download_mask(qd_ctx* ctx, int mask, int member, uint8_t* host);

/* ------------------------------------------------------------------ operators (device pointers,
 * each [B][nlat][nlon]; out must not alias in unless stated) */
int  qd_laplacian(qd_ctx* ctx, const double* in_dev, double* out_dev, const double* cos_rows_dev);
int  qd_hyperdiffuse(qd_ctx* ctx, double* f_dev /* in place */, double* scratch_dev,
                     const double* k4_rows_dev, double k4_scale, double dt, int nsub,
                     const double* cos_rows_dev);
int  qd_advect(qd_ctx* ctx, const double* in_dev, const double* u_dev, const double* v_dev,
               double* out_dev, double dt, const double* cos_rows_dev);
int  qd_shapiro(qd_ctx* ctx, double* f_dev /* in place */, double* scratch_dev, int n);
/* weights = the 2*radius+1 normalised taps as scipy's _gaussian_kernel1d evaluates them (host NumPy) */
int  qd_gaussian(qd_ctx* ctx, double* f_dev /* in place */, double* scratch_dev, int radius, int wrap,
                 const double* weights_host);
/* taps used inside qd_loop_step: which=0 -> sigma=1 'reflect' (physics.py:44,69,111,159,330),
 * which=1 -> QD_CLOUD_SMOOTH_SIGMA 'wrap' (run_simulation.py:1931) */
int  qd_set_gauss(qd_ctx* ctx, int which, int radius, int wrap, const double* weights_host);
int  qd_zonal_bandstop(qd_ctx* ctx, double* f_dev /* in place */, double cutoff, double damp);
int  qd_divergence(qd_ctx* ctx, const double* u_dev, const double* v_dev, double* out_dev);
int  qd_vorticity(qd_ctx* ctx, const double* u_dev, const double* v_dev, double* out_dev);
int  qd_median_pos(qd_ctx* ctx, const double* in_dev, double empty_value, double* out_host /* [B] */); /* sync */
int  qd_wsum(qd_ctx* ctx, const double* in_dev, double* out_host /* [B] sum(x*w) */);                    /* sync */
/* Exact-median kernel launches and first-digit speculation hits (csrc/qd_select.cuh) per call site since qd_create:
 * out[site][0] = launches (x members), out[site][1] = hits; sites: 0 positive precipitation part, 1 precipitation,
 * 2 P_cond, 3 qd_median_pos. */
int  qd_median_stats(qd_ctx* ctx, long long* out /* [4][2] */);                                            /* sync */
/* device row tables for the operator calls above */
const double* qd_row_dev(qd_ctx* ctx, int row_id);
/* upload an arbitrary [nlat] row table into one of 6 user row slots (0/1 are used by the *_host operator
 * forms, 2..4 hold the QD_OCEAN_K4_U/V/ETA overrides) (the library appends the five
 * companion rows the Laplacian kernels expect right behind it), returns its device pointer */
const double* qd_user_row(qd_ctx* ctx, int slot, const double* rows_host);   /* every member's copy */
int  qd_user_row_member(qd_ctx* ctx, int slot, int member, const double* rows_host);

/* host-buffer convenience forms of the three jax_compat seam kernels (H2D + kernel + D2H, sync);
 * arrays are single-member [nlat][nlon] */
int  qd_laplacian_host(qd_ctx* ctx, const double* in, double* out, const double* cos_rows);
int  qd_hyperdiffuse_host(qd_ctx* ctx, const double* in, double* out, const double* k4_map /* [nlat][nlon] or NULL */,
                          double k4_scalar, double dt, int nsub, const double* cos_rows);
int  qd_advect_host(qd_ctx* ctx, const double* in, const double* u, const double* v, double* out,
                    double dt, const double* cos_rows);

/* ------------------------------------------------------------------ step level (bound fields) */
int  qd_atmos_step(qd_ctx* ctx, const qd_step_cfg_t* cfg);   /* Teq in QD_F_TEQ, albedo in QD_F_ALBEDO */
int  qd_ocean_step(qd_ctx* ctx, const qd_step_cfg_t* cfg);   /* winds QD_F_U/V, Q in QD_F_QNET, ice in QD_M_ICE */
/* WindDrivenSlabOcean.step(dt, u_atm, v_atm, ...) with the winds as caller-owned device arrays [B][nlat][nlon]
 * (ocean.py:265): the atmosphere's QD_F_U/V are not touched */
int  qd_ocean_step_winds(qd_ctx* ctx, const qd_step_cfg_t* cfg, const double* u_atm_dev, const double* v_atm_dev);
int  qd_loop_step(qd_ctx* ctx, const qd_step_cfg_t* cfg, const qd_forcing_t* forcing, int nsteps);
int  qd_last_nsub(qd_ctx* ctx, int* out_host /* [B] */);      /* sync */
/* 2 (default): a whole loop step is one CUDA graph (kernels, memsets, cooperative selects, WHILE node
 * for the ocean's data-dependent sub-step loop); 1: stream launches + WHILE-node graph for the ocean
 * loop only; 0: stream launches and a host loop with one scalar read-back per step */
int  qd_use_graphs(qd_ctx* ctx, int enable);
/* captured graphs currently cached / captures that failed and fell back to stream mode (a 2x slowdown nobody should
 * discover by accident: bench.py asserts failed == 0).  Graphs are dropped and re-captured whenever an entry point changes
 * something they bake in (qd_bind, qd_set_params, qd_set_gauss, qd_route_setup, qd_eco_bind, qd_band_connect). */
int  qd_graph_status(qd_ctx* ctx, int* live, int* failed);
int  qd_set_counters(qd_ctx* ctx, int atm_counter, int ocean_counter, int has_cloud_eff);
int  qd_get_counters(qd_ctx* ctx, int* atm_counter, int* ocean_counter, int* has_cloud_eff);
int  qd_minmax(qd_ctx* ctx, const double* in_dev, double* out_host /* [B][2] */);   /* sync */
int  qd_set_gauss2d(qd_ctx* ctx, int mode);                   /* 0: two one-axis passes; 1 (default): fused tiles, TMA box loads; 2: generic tile kernel; 3: no TMA (A/B runs) */
int  qd_set_h4_stream(qd_ctx* ctx, int enable);               /* 0: force the tile kernel for del^4 (tests); default 1 */
/* 1: the ocean's CFL sub-steps (ocean.py:305-444) run as two kernels -- a warp-streaming kernel that fuses momentum,
 * del^4 of (uo, vo, eta), continuity, the SST gather and the current hygiene, plus a closing kernel -- instead of four.
 * Halves the DRAM traffic of a sub-step; on B200 it is fp64-issue bound and measured ~10 % SLOWER than the four-kernel
 * form, so it is opt-in.  Default 0. */
int  qd_set_ocean_fused(qd_ctx* ctx, int enable);
int  qd_launch_count(qd_ctx* ctx, long long* out);            /* kernels launched so far */
/* Self-test of the device exp / tanh (csrc/qd_math.cuh: libdevice's algorithms with constant-bank coefficients):
 * out_host[0..n) = the CUDA library call, out_host[n..2n) = the routine the kernels use, for which = 0 (exp) / 1 (tanh).
 * The two halves must be bit-identical (tests/test_gpu.py).  Host check build: both halves are libm. */
int  qd_math_check(qd_ctx* ctx, const double* x_host, long long n, double* out_host /* [2][n] */, int which);   /* sync */
/* per-kernel device time: CUDA events on the launching stream around every launch while enabled */
int  qd_profile(qd_ctx* ctx, int enable);
int  qd_profile_report(qd_ctx* ctx, char* buf, int buflen);   /* "name count total_ms" lines; sync */

/* ------------------------------------------------------------------ routing (routing.py:211-335) */
/* ------------------------------------------------------------------ latitude bands (one domain over the GPUs of a node)
 * BASELINE configs[4] / SURVEY 8e: rank r of `world` computes a contiguous block of latitude rows (full longitude
 * circles); halos, partial sums and median histograms travel through per-rank exchange buffers that the peers map
 * (CUDA IPC) and write directly over NVLink.  Call order on every rank: qd_create, qd_bind, qd_band_init,
 * qd_band_export -> exchange the 64-byte handles out of band (the host side uses torch.distributed) ->
 * qd_band_connect.  Afterwards qd_loop_step runs the partitioned step; fields keep their full-size layout and a
 * rank's own rows [own0, own1) are authoritative.  batch must be 1; routing and the ecology coupling are not
 * partitioned. */
int  qd_band_init(qd_ctx* ctx, int rank, int world, int halo_rows);
int  qd_band_export(qd_ctx* ctx, void* handle64);
int  qd_band_connect(qd_ctx* ctx, const void* handles /* [world][64] in rank order */);
int  qd_band_info(qd_ctx* ctx, int* own0, int* own1, int* halo_rows, int* error_word /* != 0: a bounded wait expired */);
/* Tuning aid: `iters` back-to-back halo exchanges of the first `nfields` field slots (values are exchanged as they
 * are: the halo rows of those fields are overwritten with the neighbours' current rows); *ms_out = device time of the
 * loop on this rank.  Every rank of the band group must call it with the same arguments. */
int  qd_band_exchange_bench(qd_ctx* ctx, int nfields, int iters, float* ms_out);                /* sync */

/* ------------------------------------------------------------------ ecology sub-daily
 * Replaces EcologyAdapter.step_subdaily (adapter.py:140-186) + PopulationManager.step_subdaily /
 * _should_recompute_canopy / _recompute_canopy_cache / canopy_reflectance_factor
 * (population.py:252-286,831-841,895-915) and get_surface_albedo_bands (population.py:875-892).
 * lai_layers_dev: caller-owned [B][n_layers][nlat][nlon] f64 = LAI_layers_SK flattened over species x layers
 * (the daily ecology stays host Python and re-uploads it when it changes). */
#define QD_ECO_MAX_BANDS 64
int  qd_eco_bind(qd_ctx* ctx, const double* lai_layers_dev, int n_layers, double k_canopy,
                 double update_every_hours, double lai_delta, int substep_every_nphys);
/* clocks as PopulationManager.__init__ leaves them (hours=0, next=update_every_hours, no cache), the LAI
 * snapshot = total LAI, adapter step counter = 0 */
int  qd_eco_reset(qd_ctx* ctx, double hours, double next_hours, int cached, int step_count);
int  qd_eco_state(qd_ctx* ctx, int* step_count, int* have_alpha, int set);   /* cadence counters of step_subdaily (adapter.py:150-156): read (set=0) / restore (set=1); checkpoints */
/* one EcologyAdapter.step_subdaily call outside the fused loop: isr_dev [B][nlat][nlon]; alpha_dev receives the
 * land alpha map (NaN on ocean) when the call is on the QD_ECO_SUBSTEP_EVERY_NPHYS cadence (*produced = 1) */
int  qd_eco_subdaily(qd_ctx* ctx, const double* isr_dev, double dt, double* alpha_dev, int* produced);
/* A_b^surface [B][nb][nlat][nlon] (NaN on ocean) from the cached canopy factor */
int  qd_eco_bands(qd_ctx* ctx, int nb, const double* r_eff_host, double soil_ref, double* out_dev);

/* ------------------------------------------------------------------ individual pool sub-steps
 * IndividualPool.try_substep (pygcm/ecology/individuals.py:142-191) with the NB-band split of the dual-star insolation
 * (spectral.dual_star_insolation_to_bands, spectral.py:388-426).  cell_host[n] = j * nlon + i of each individual's
 * sampled cell, ab_host[n][nb], tol_host[n]; spec_a / spec_b / t_ray [nb] = per-star blackbody band weights and the
 * Rayleigh factors.  qd_indiv_substep reads the context's per-star insolation fields QD_F_ISR_A / QD_F_ISR_B. */
int  qd_indiv_setup(qd_ctx* ctx, int n_indiv, int nb, const int* cell_host, const double* ab_host, const double* tol_host,
                    const double* spec_a, const double* spec_b, const double* t_ray);
int  qd_indiv_substep(qd_ctx* ctx, const double* soil_dev, double soil_scalar, double period, double day_length);
int  qd_indiv_state(qd_ctx* ctx, double* e_day_host, double* stress_host, int upload);

/* ------------------------------------------------------------------ routing-network builder (host C++)
 * scripts/generate_hydrology_maps.py:85-273: pit_fill, compute_flow_to_index, identify_lakes, compute_lake_outlets,
 * topo_sort_flow_order with their sequential semantics (bit-identical outputs).  dist[3][nlat][3][3]: centre distance
 * to the neighbour at (dj, di), evaluated by the caller with the reference's NumPy expression (:65-82). */
int  qd_net_build(int nlat, int nlon, double* elev_inout, const uint8_t* land_mask, const double* dist, int pit_iters,
                  double pit_eps, int64_t* flow_to, int64_t* flow_order, int64_t* n_order, uint8_t* lake_mask,
                  int32_t* lake_id, int32_t* lake_outlet, int* n_lakes, int* sweeps);

/* ------------------------------------------------------------------ global diagnostics (one launch, warp-shuffle reductions)
 * energy.compute_energy_diagnostics (energy.py:494-538), hydrology.diagnose_water_closure (hydrology.py:270-340),
 * WindDrivenSlabOcean.diagnostics (ocean.py:535-561).  out_host: [B][qd_diag_count()] doubles: slot 0 = sum of the
 * area weights, then weighted SUMS (divide by slot 0 + 1e-15 for the reference's means) of T_s, h, q, cloud, h_ice,
 * W_land, S_snow, E, precip, R_land, albedo, SST, I, R, OLR, SW_sfc, LW_sfc, SH, LH, ocean KE, then eta min / max,
 * max |U_ocean|, max |u|, T_s min / max (qd_diag.cuh: QD_D_*). */
int  qd_diag_count(void);
int  qd_diag(qd_ctx* ctx, double* out_host);

/* ------------------------------------------------------------------ phytoplankton tracer transport
 * PhytoManager.advect_diffuse (pygcm/ecology/phyto.py:496-547, called every physics step at
 * scripts/run_simulation.py:2256-2258): conc_dev [S][nlat][nlon] f64 in place; uo_dev / vo_dev [nlat][nlon] or NULL
 * for the context's own ocean currents. */
int  qd_phyto_advect_diffuse(qd_ctx* ctx, double* conc_dev, int n_species, const double* uo_dev, const double* vo_dev,
                             double dt, double adv_alpha, double k_h);

int  qd_route_setup(qd_ctx* ctx, int n_order, const int64_t* flow_order_host,
                    const int64_t* flow_to_host /* [ncell] */, const uint8_t* land_host,
                    const uint8_t* lake_host, const int32_t* lake_id_host,
                    int n_lakes, const int64_t* lake_outlet_host);
int  qd_route_levels(qd_ctx* ctx);                            /* depth of the donor DAG, -1 if not set up */
int  qd_route_accumulate(qd_ctx* ctx, double dt);             /* buffer += where(land, R*area*dt, 0) */
/* one event for one member (sync): flow accumulation [ncell] kg, ocean inflow kg, per-cell residual
 * [ncell], the routed input buffer [ncell], lake stores [n_lakes]; clears the member's buffer */
int  qd_route_event(qd_ctx* ctx, int member, double* flow_accum_kg_host, double* ocean_inflow_kg,
                    double* residual_host, double* input_host, double* lake_store_kg_host);
int  qd_route_buffer(qd_ctx* ctx, int member, double* host /* [ncell] */, int upload);   /* sync */

#ifdef __cplusplus
}
#endif
#endif /* QD_B200_H */
