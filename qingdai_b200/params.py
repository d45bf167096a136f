"""Parameter snapshot for the Qingdai hot path.

The reference re-reads ``QD_*`` environment variables with ``os.getenv`` inside the step
(e.g. pygcm/dynamics.py:330-341,484,534-536,549,581,611-631,650; pygcm/ocean.py:49-82,
350-352,380,389,399,438; pygcm/energy.py:66-74,122,127,149-151,187-188,381-385;
pygcm/humidity.py:71-82; pygcm/hydrology.py:65-80; scripts/run_simulation.py:1593-1627,
1777,1878,1892-1925).  A compiled path cannot do that per cell, so the values are
snapshotted once into :class:`QDParams` (and from there into the C ``qd_params_t``).
Defaults are the *code* defaults of the reference (they win over its docs).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, fields, replace
from typing import Mapping, Optional


def _f(env: Mapping[str, str], key: str, default: float) -> float:
    try:
        return float(env.get(key, default))
    except (TypeError, ValueError):
        return float(default)


def _i(env: Mapping[str, str], key: str, default: int) -> int:
    try:
        return int(env.get(key, default))
    except (TypeError, ValueError):
        return int(default)


def _opt(env: Mapping[str, str], key: str) -> Optional[float]:
    v = env.get(key)
    if v in (None, "", "None", "none", "null"):
        return None
    try:
        return float(v)
    except ValueError:
        return None


@dataclass
class QDParams:
    # ---- atmosphere core (dynamics.py:22-40, :304-322, :464-467) ----
    g: float = 9.81
    H: float = 8000.0
    tau_rad: float = 10 * 24 * 3600.0
    gh_newton: float = 0.40            # SpectralModel(greenhouse_factor=QD_GH_FACTOR->0.40) run_simulation.py:1267
    energy_w: float = 0.0              # QD_ENERGY_W dynamics.py:316
    mom_scheme: str = "geos"           # QD_MOM_SCHEME dynamics.py:484
    # ---- energy (energy.py:44-74, :101-234, :291-420; dynamics.py:330-401,472) ----
    sw_a0: float = 0.06
    sw_kc: float = 0.20
    lw_eps0: float = 0.70
    lw_kc: float = 0.20
    t_floor: float = 150.0
    c_sfc: float = 2.0e7
    cloud_couple: bool = True
    rh0: float = 0.6
    k_q: float = 0.3
    k_p: float = 0.4
    pcond_ref: Optional[float] = None
    lw_v2: bool = True
    hice_ref: float = 0.5
    eps_ocean: float = 0.98
    eps_land: float = 0.96
    eps_ice: float = 0.99
    lw_tau0: float = 6.0
    lw_ktau: float = 1.0
    gh_lock: bool = True
    gh_factor_lw: float = 0.582        # energy.py:127,224 default when QD_GH_FACTOR unset
    C_H: float = 1.5e-3
    cp_air: float = 1004.0
    seaice_enabled: bool = True
    t_freeze: float = 271.35
    rho_i: float = 917.0
    L_f: float = 3.34e5
    Cs_ocean: float = 1000.0 * 4200.0 * 50.0
    Cs_land: float = 3.0e6
    Cs_ice: float = 5.0e6
    polar_fix_s: bool = True
    polar_fix_n: bool = True
    atm_H: float = 800.0               # QD_ATM_H default = h_mbl dynamics.py:472
    # ---- numerics (dynamics.py:534-650) ----
    diff_enable: bool = True
    filter_type: str = "combo"
    diff_every: int = 1
    sigma4: float = 0.02
    k4_u: Optional[float] = None
    k4_v: Optional[float] = None
    k4_h: Optional[float] = None
    k4_q: Optional[float] = None
    k4_c: Optional[float] = None
    k4_nsub: int = 1
    diff_q: bool = False
    diff_cloud: bool = False
    shapiro_every: int = 6
    shapiro_n: int = 2
    spec_every: int = 0
    spec_cutoff: float = 0.75
    spec_damp: float = 0.5
    diff_factor: float = 0.998
    # ---- humidity (humidity.py:37-82; dynamics.py:81) ----
    C_E: float = 1.3e-3
    rho_a: float = 1.2
    h_mbl: float = 800.0
    L_v: float = 2.5e6
    p0: float = 1.0e5
    ocean_evap_scale: float = 1.0
    land_evap_scale: float = 0.5
    ice_evap_scale: float = 0.05
    tau_cond: float = 1800.0
    q_init_rh: float = 0.5
    # ---- ocean (ocean.py:49-82, :380-399, :438, :519-533) ----
    oc_H: float = 50.0
    oc_rho_w: float = 1000.0
    oc_cp_w: float = 4200.0
    oc_g: float = 9.81
    oc_CD: float = 1.5e-3
    oc_r_bot: float = 2.0e-5
    oc_rho_a: float = 1.2
    oc_vcap: float = 15.0
    oc_tau_scale: float = 0.2
    oc_polar_lat0: float = 70.0
    oc_polar_gain: float = 5.0e-5
    oc_K_h: float = 5.0e3
    oc_sigma4: float = 0.02
    oc_k4_nsub: int = 1
    oc_diff_every: int = 1
    oc_shapiro_n: int = 0
    oc_shapiro_every: int = 8
    oc_cfl: float = 0.5
    oc_max_u: float = 3.0
    oc_outlier: str = "mean4"
    oc_adv_alpha: float = 0.7
    oc_use_qnet: bool = True
    oc_ice_qfac: float = 0.2
    oc_eta_cap: float = 5.0
    oc_polar_fix: bool = True
    oc_ts_min: float = 150.0
    oc_ts_max: float = 340.0
    oc_k4_u: Optional[float] = None
    oc_k4_v: Optional[float] = None
    oc_k4_eta: Optional[float] = None
    # ---- script loop physics (run_simulation.py:1605-1627,1777,1878-1925; physics.py:336-338) ----
    D_crit: float = -1e-7
    k_precip: float = 1e5
    alpha_water: float = 0.1
    alpha_ice: float = 0.6
    alpha_cloud: float = 0.5
    use_topo_albedo: bool = True
    orog_enabled: bool = False
    k_orog: float = 7e-4
    beta_div: float = 0.4
    p_fallback: bool = True
    pq_min: float = 1e-8
    p_blend: float = 0.6
    pref: Optional[float] = None
    cmax: float = 0.95
    w_mem: float = 0.4
    w_p: float = 0.4
    w_src: float = 0.2
    cloud_floor: float = 0.8
    cloud_advect: bool = True
    cloud_adv_alpha: float = 0.7
    cloud_smooth_sigma: float = 0.2
    lapse_enable: bool = True
    lapse_kpm: float = 6.5
    land_elev_max: float = 10000.0
    polar_ice_thick_max: float = 4500.0
    polar_lat_thresh: float = 60.0
    rho_snow: float = 300.0
    glacier_frac: float = 0.60
    glacier_swe: float = 50.0
    # ---- hydrology (hydrology.py:27-80,172) ----
    runoff_tau_days: float = 10.0
    wland_cap: Optional[float] = None
    snow_thresh: float = 273.15
    snow_melt_rate: float = 5.0
    snow_t_band: float = 1.5
    snow_melt_mode: str = "degree_day"
    snow_ddf: float = 3.0
    snow_melt_tref: float = 273.15
    swe_enable: bool = True
    swe_ref: float = 15.0
    swe_max: Optional[float] = None
    snow_albedo_fresh: float = 0.70
    hydro_dt_hours: float = 6.0
    # ---- ecology sub-daily (adapter.py:140-186; run_simulation.py:2089) ----
    eco_lai_albedo_weight: float = 1.0
    eco_soil_reflect: float = 0.20

    def replace(self, **kw) -> "QDParams":
        return replace(self, **kw)

    @classmethod
    def from_env(cls, env: Optional[Mapping[str, str]] = None, **overrides) -> "QDParams":
        """Snapshot every ``QD_*`` variable the hot path consults (env defaults to os.environ)."""
        e = os.environ if env is None else env
        rho_w = _f(e, "QD_RHO_W", 1000.0)
        cp_w = _f(e, "QD_CP_W", 4200.0)
        mld = _f(e, "QD_MLD_M", 50.0)
        h_mbl = _f(e, "QD_MBL_H", 800.0)
        lapse = _f(e, "QD_LAPSE_K_KPM", 6.5)
        p = cls(
            gh_newton=_f(e, "QD_GH_FACTOR", 0.40),
            energy_w=_f(e, "QD_ENERGY_W", 0.0),
            mom_scheme=str(e.get("QD_MOM_SCHEME", "geos")).lower(),
            sw_a0=_f(e, "QD_SW_A0", 0.06), sw_kc=_f(e, "QD_SW_KC", 0.20),
            lw_eps0=_f(e, "QD_LW_EPS0", 0.70), lw_kc=_f(e, "QD_LW_KC", 0.20),
            t_floor=_f(e, "QD_T_FLOOR", 150.0), c_sfc=_f(e, "QD_CS", 2.0e7),
            cloud_couple=_i(e, "QD_CLOUD_COUPLE", 1) == 1,
            rh0=_f(e, "QD_RH0", 0.6), k_q=_f(e, "QD_K_Q", 0.3), k_p=_f(e, "QD_K_P", 0.4),
            pcond_ref=_opt(e, "QD_PCOND_REF"),
            lw_v2=_i(e, "QD_LW_V2", 1) == 1, hice_ref=_f(e, "QD_HICE_REF", 0.5),
            eps_ocean=_f(e, "QD_EPS_OCEAN", 0.98), eps_land=_f(e, "QD_EPS_LAND", 0.96),
            eps_ice=_f(e, "QD_EPS_ICE", 0.99),
            lw_tau0=_f(e, "QD_LW_TAU0", 6.0), lw_ktau=_f(e, "QD_LW_KTAU", 1.0),
            gh_lock=_i(e, "QD_GH_LOCK", 1) == 1, gh_factor_lw=_f(e, "QD_GH_FACTOR", 0.582),
            C_H=_f(e, "QD_CH", 1.5e-3), cp_air=_f(e, "QD_CP_A", 1004.0),
            seaice_enabled=_i(e, "QD_USE_SEAICE", 1) == 1,
            t_freeze=_f(e, "QD_T_FREEZE", 271.35), rho_i=_f(e, "QD_RHO_ICE", 917.0),
            L_f=_f(e, "QD_LF", 3.34e5),
            Cs_ocean=rho_w * cp_w * mld, Cs_land=_f(e, "QD_CS_LAND", 3e6), Cs_ice=_f(e, "QD_CS_ICE", 5e6),
            polar_fix_s=_i(e, "QD_POLAR_FREEZE_FIX", 1) == 1,
            polar_fix_n=_i(e, "QD_POLAR_FREEZE_FIX_N", 1) == 1,
            atm_H=_f(e, "QD_ATM_H", h_mbl),
            diff_enable=_i(e, "QD_DIFF_ENABLE", 1) == 1,
            filter_type=str(e.get("QD_FILTER_TYPE", "combo")).lower(),
            diff_every=_i(e, "QD_DIFF_EVERY", 1), sigma4=_f(e, "QD_SIGMA4", 0.02),
            k4_u=_opt(e, "QD_K4_U"), k4_v=_opt(e, "QD_K4_V"), k4_h=_opt(e, "QD_K4_H"),
            k4_q=_opt(e, "QD_K4_Q"), k4_c=_opt(e, "QD_K4_CLOUD"),
            k4_nsub=_i(e, "QD_K4_NSUB", 1),
            diff_q=_i(e, "QD_DIFF_Q", 0) == 1, diff_cloud=_i(e, "QD_DIFF_CLOUD", 0) == 1,
            shapiro_every=_i(e, "QD_SHAPIRO_EVERY", 6), shapiro_n=_i(e, "QD_SHAPIRO_N", 2),
            spec_every=_i(e, "QD_SPEC_EVERY", 0), spec_cutoff=_f(e, "QD_SPEC_CUTOFF", 0.75),
            spec_damp=_f(e, "QD_SPEC_DAMP", 0.5), diff_factor=_f(e, "QD_DIFF_FACTOR", 0.998),
            C_E=_f(e, "QD_CE", 1.3e-3), rho_a=_f(e, "QD_RHO_A", 1.2), h_mbl=h_mbl,
            L_v=_f(e, "QD_LV", 2.5e6), p0=_f(e, "QD_P0", 1.0e5),
            ocean_evap_scale=_f(e, "QD_OCEAN_EVAP_SCALE", 1.0),
            land_evap_scale=_f(e, "QD_LAND_EVAP_SCALE", 0.5),
            ice_evap_scale=_f(e, "QD_ICE_EVAP_SCALE", 0.05),
            tau_cond=_f(e, "QD_TAU_COND", 1800.0), q_init_rh=_f(e, "QD_Q_INIT_RH", 0.5),
            oc_H=_f(e, "QD_OCEAN_H_M", mld), oc_rho_w=rho_w, oc_cp_w=cp_w,
            oc_CD=_f(e, "QD_CD", 1.5e-3), oc_r_bot=_f(e, "QD_R_BOT", 2.0e-5),
            oc_rho_a=_f(e, "QD_RHO_A", 1.2), oc_vcap=_f(e, "QD_WIND_STRESS_VCAP", 15.0),
            oc_tau_scale=_f(e, "QD_TAU_SCALE", 0.2),
            oc_polar_lat0=_f(e, "QD_POLAR_SPONGE_LAT", 70.0),
            oc_polar_gain=_f(e, "QD_POLAR_SPONGE_GAIN", 5.0e-5),
            oc_K_h=_f(e, "QD_KH_OCEAN", 5.0e3), oc_sigma4=_f(e, "QD_SIGMA4_OCEAN", 0.02),
            oc_k4_nsub=_i(e, "QD_OCEAN_K4_NSUB", 1), oc_diff_every=_i(e, "QD_OCEAN_DIFF_EVERY", 1),
            oc_shapiro_n=_i(e, "QD_OCEAN_SHAPIRO_N", 0), oc_shapiro_every=_i(e, "QD_OCEAN_SHAPIRO_EVERY", 8),
            oc_cfl=_f(e, "QD_OCEAN_CFL", 0.5), oc_max_u=_f(e, "QD_OCEAN_MAX_U", 3.0),
            oc_outlier=str(e.get("QD_OCEAN_OUTLIER", "mean4")).strip().lower(),
            oc_adv_alpha=_f(e, "QD_OCEAN_ADV_ALPHA", 0.7),
            oc_use_qnet=_i(e, "QD_OCEAN_USE_QNET", 1) == 1,
            oc_ice_qfac=_f(e, "QD_OCEAN_ICE_QFAC", 0.2), oc_eta_cap=_f(e, "QD_ETA_CAP", 5.0),
            oc_polar_fix=_i(e, "QD_OCEAN_POLAR_FIX", 1) == 1,
            oc_ts_min=_f(e, "QD_TS_MIN", 150.0), oc_ts_max=_f(e, "QD_TS_MAX", 340.0),
            oc_k4_u=_opt(e, "QD_OCEAN_K4_U"), oc_k4_v=_opt(e, "QD_OCEAN_K4_V"),
            oc_k4_eta=_opt(e, "QD_OCEAN_K4_ETA"),
            use_topo_albedo=_i(e, "QD_USE_TOPO_ALBEDO", 1) == 1,
            orog_enabled=_i(e, "QD_OROG", 0) == 1, k_orog=_f(e, "QD_OROG_K", 7e-4),
            beta_div=_f(e, "QD_P_BETADIV", 0.4),
            p_fallback=_i(e, "QD_P_HYBRID_FALLBACK", 1) == 1,
            pq_min=_f(e, "QD_PQ_MIN", 1e-8), p_blend=_f(e, "QD_P_BLEND", 0.6),
            pref=_opt(e, "QD_PREF"), cmax=_f(e, "QD_CMAX", 0.95),
            w_mem=_f(e, "QD_W_MEM", 0.4), w_p=_f(e, "QD_W_P", 0.4), w_src=_f(e, "QD_W_SRC", 0.2),
            cloud_floor=_f(e, "QD_CLOUD_FROM_P_FLOOR", 0.8),
            cloud_advect=_i(e, "QD_CLOUD_ADVECT", 1) == 1,
            cloud_adv_alpha=_f(e, "QD_CLOUD_ADV_ALPHA", 0.7),
            cloud_smooth_sigma=_f(e, "QD_CLOUD_SMOOTH_SIGMA", 0.2),
            lapse_enable=_i(e, "QD_LAPSE_ENABLE", 1) == 1, lapse_kpm=lapse,
            land_elev_max=_f(e, "QD_LAND_ELEV_MAX_M", 10000.0),
            polar_ice_thick_max=_f(e, "QD_POLAR_ICE_THICK_MAX_M", 4500.0),
            polar_lat_thresh=_f(e, "QD_POLAR_LAT_THRESH", 60.0),
            rho_snow=_f(e, "QD_RHO_SNOW", 300.0),
            glacier_frac=_f(e, "QD_GLACIER_FRAC", 0.60), glacier_swe=_f(e, "QD_GLACIER_SWE_MM", 50.0),
            runoff_tau_days=_f(e, "QD_RUNOFF_TAU_DAYS", 10.0), wland_cap=_opt(e, "QD_WLAND_CAP"),
            snow_thresh=_f(e, "QD_SNOW_THRESH", 273.15), snow_melt_rate=_f(e, "QD_SNOW_MELT_RATE", 5.0),
            snow_t_band=_f(e, "QD_SNOW_T_BAND", 1.5),
            snow_melt_mode=str(e.get("QD_SNOW_MELT_MODE", "degree_day")).strip().lower(),
            snow_ddf=_f(e, "QD_SNOW_DDF_MM_PER_K_DAY", 3.0),
            snow_melt_tref=_f(e, "QD_SNOW_MELT_TREF", 273.15),
            swe_enable=_i(e, "QD_SWE_ENABLE", 1) == 1, swe_ref=_f(e, "QD_SWE_REF_MM", 15.0),
            swe_max=_opt(e, "QD_SWE_MAX_MM"), snow_albedo_fresh=_f(e, "QD_SNOW_ALBEDO_FRESH", 0.70),
            hydro_dt_hours=_f(e, "QD_HYDRO_DT_HOURS", 6.0),
            eco_lai_albedo_weight=_f(e, "QD_ECO_LAI_ALBEDO_WEIGHT", 1.0),
            eco_soil_reflect=_f(e, "QD_ECO_SOIL_REFLECT", 0.20),
        )
        if overrides:
            names = {f.name for f in fields(cls)}
            bad = set(overrides) - names
            if bad:
                raise TypeError(f"unknown QDParams fields: {sorted(bad)}")
            p = replace(p, **overrides)
        return p
