"""Engine: owns the device state of one (grid, ensemble) and drives libqd_b200 through the C ABI.

PyTorch is plumbing here: it owns the float64 field block ``[QD_F_COUNT, B, n_lat, n_lon]`` and the
uint8 mask block in HBM and provides the CUDA stream; every arithmetic operation of the hot path
happens inside the hand-written kernels of ``csrc/`` (no torch ops, no CPU fallback).

Row / column metric tables and the parameter vector are evaluated on the host with the
reference's NumPy expressions (cited inline) so metric terms are bit-identical.
"""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import fields as dc_fields
from typing import Optional, Sequence

import numpy as np
import torch

from . import constants as const
from ._binding import ENUM, NF, NM, NR, NC, NP, NS, Forcing, Library, StepCfg, default_library
from .params import QDParams

F = {k[5:].lower(): v for k, v in ENUM.items() if k.startswith("QD_F_") and k != "QD_F_COUNT"}
M = {k[5:].lower(): v for k, v in ENUM.items() if k.startswith("QD_M_") and k != "QD_M_COUNT"}
R = {k[5:].lower(): v for k, v in ENUM.items() if k.startswith("QD_R_") and k != "QD_R_COUNT"}
S = {k[5:].lower(): v for k, v in ENUM.items() if k.startswith("QD_S_") and k != "QD_S_COUNT"}
P = {k[5:].lower(): v for k, v in ENUM.items() if k.startswith("QD_P_") and k != "QD_P_COUNT"}


def _ptr(a):
    if isinstance(a, torch.Tensor):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


def gaussian_taps(sigma, truncate=4.0):
    """scipy.ndimage._gaussian_kernel1d (order 0): radius and normalised taps, evaluated by NumPy."""
    r = int(truncate * float(sigma) + 0.5)
    x = np.arange(-r, r + 1)
    w = np.exp(-0.5 / (float(sigma) * float(sigma)) * x ** 2)
    return r, np.ascontiguousarray(w / w.sum(), dtype=np.float64)


def row_tables(nlat, nlon, p: QDParams, dt):
    """[QD_R_COUNT, nlat] metric rows.  Every expression is the reference's (file:line cited)."""
    a = const.PLANET_RADIUS
    lat = np.linspace(-90, 90, nlat)                                   # grid.py:27
    lon = np.linspace(0, 360, nlon)                                    # grid.py:29
    dlat = np.deg2rad(lat[1] - lat[0])                                 # grid.py:38
    dlon = np.deg2rad(lon[1] - lon[0])                                 # grid.py:39
    lat_rad = np.deg2rad(lat)
    cos = np.cos(lat_rad)
    rows = np.zeros((NR, nlat), dtype=np.float64)
    rows[R["lat_deg"]] = lat
    rows[R["cos"]] = cos
    rows[R["sin"]] = np.sin(lat_rad)
    f = 2 * const.PLANET_OMEGA * np.sin(lat_rad)                       # grid.py:95
    rows[R["fcor"]] = f
    rows[R["w"]] = np.maximum(cos, 0.0)                                # energy.py:520-521
    rows[R["cos_adv_atm"]] = np.maximum(1e-6, cos)                     # dynamics.py:104
    rows[R["cos_adv_half"]] = np.maximum(cos, 0.5)                     # ocean.py:82, run_simulation.py:1145
    rows[R["cos_lap_atm"]] = np.maximum(cos, 0.2)                      # dynamics.py:164
    for base in ("half", "lap_atm"):                                   # reciprocal rows for the stencil kernels
        c_ = rows[R["cos_adv_half" if base == "half" else "cos_lap_atm"]]
        rows[R["icos_" + base]] = 1.0 / c_
        rows[R["icos2_" + base]] = 1.0 / (c_ ** 2)
    rows[R["cos_cap"]] = np.maximum(cos, 1e-6)                         # dynamics.py:490, grid.py:52
    f_min = 2.0 * const.PLANET_OMEGA * np.sin(np.deg2rad(5.0))         # dynamics.py:516-518
    sgn = np.where(f >= 0.0, 1.0, -1.0)
    rows[R["fsafe"]] = np.where(np.abs(f) < f_min, sgn * f_min, f)
    cos3 = np.maximum(cos, 1e-3)                                       # dynamics.py:559-563
    dx_min = np.minimum(a * dlat, a * dlon * cos3)
    base = p.sigma4 * (dx_min ** 4) / max(1e-12, dt)
    for name, scale, ov in (("k4_u", None, p.k4_u), ("k4_v", None, p.k4_v), ("k4_h", 0.5, p.k4_h),
                            ("k4_q", 0.5, p.k4_q), ("k4_c", 0.25, p.k4_c)):      # dynamics.py:566-570
        rows[R[name]] = float(ov) if ov is not None else (base if scale is None else scale * base)
    cosh = np.maximum(cos, 0.5)
    rows[R["oc_s4dx4"]] = p.oc_sigma4 * (np.minimum(a * dlat, a * dlon * cosh) ** 4)   # ocean.py:344-347
    lat_abs = np.abs(np.rad2deg(lat_rad))                              # ocean.py:332-334
    s = np.clip((lat_abs - p.oc_polar_lat0) / max(1e-6, 90.0 - p.oc_polar_lat0), 0.0, 1.0)
    rows[R["oc_sponge"]] = p.oc_polar_gain * (s ** 2)
    rows[R["polar"]] = (np.abs(lat) >= p.polar_lat_thresh).astype(np.float64)          # run_simulation.py:1956
    dphi = np.deg2rad(abs(lat[1] - lat[0]))                            # routing.py:176-200
    dlam = np.deg2rad(abs(lon[1] - lon[0]))
    band = np.sin(np.clip(lat_rad + 0.5 * dphi, -0.5 * np.pi, 0.5 * np.pi)) - np.sin(np.clip(lat_rad - 0.5 * dphi, -0.5 * np.pi, 0.5 * np.pi))
    rows[R["area"]] = (a * a) * dlam * band
    rows[R["inv_acos_half"]] = 1.0 / (a * rows[R["cos_adv_half"]])
    rows[R["inv_acos_cap"]] = 1 / (a * rows[R["cos_cap"]])            # grid.py:66 evaluates exactly this factor
    rows[R["iac_adv_atm"]] = 1.0 / (a * rows[R["cos_adv_atm"]])       # correctly rounded reciprocal of the divisor in dynamics.py:104
    cols = np.zeros((NC, nlon), dtype=np.float64)
    lon_rad = np.deg2rad(lon)
    cols[ENUM["QD_C_LON_RAD"]] = lon_rad
    cols[ENUM["QD_C_SIN_LON"]] = np.sin(lon_rad)
    cols[ENUM["QD_C_COS_LON"]] = np.cos(lon_rad)
    return rows, cols, float(dlat), float(dlon), lat, lon


_NAN = float("nan")


def speed2_threshold(cap):
    """Smallest double t with sqrt(t) > cap (IEEE sqrt is correctly rounded and monotone, on the host as on the device):
    ``sqrt(uo*uo + vo*vo) > cap`` (ocean.py:412) and ``uo*uo + vo*vo >= t`` select exactly the same cells."""
    import math
    cap = float(cap)
    if cap != cap or cap == math.inf:
        return math.inf                      # nothing exceeds a NaN / infinite cap
    if cap < 0.0:
        return 0.0                           # every non-NaN speed exceeds a negative cap
    t = min(cap * cap, 1.7976931348623157e308)
    while t > 0.0 and math.sqrt(t) > cap:
        t = math.nextafter(t, 0.0)
    while not math.sqrt(t) > cap:
        t = math.nextafter(t, math.inf)
    return t


def param_vector(p: QDParams, nlat, nlon, land_mask=None, has_elevation=False, eco_alpha_leaf=0.0, eco_enable=False):
    """One [QD_P_COUNT] float64 vector from a QDParams snapshot (+ host-evaluated sums)."""
    a = const.PLANET_RADIUS
    v = np.zeros(NP, dtype=np.float64)

    def opt(x):
        return _NAN if x is None else float(x)
    wm, wp, ws = p.w_mem, p.w_p, p.w_src                              # run_simulation.py:1892-1898
    wsum = wm + wp + ws
    if wsum <= 0:
        wm, wp, ws, wsum = 0.5, 0.4, 0.1, 1.0
    wm /= wsum
    wp /= wsum
    ws /= wsum
    lat = np.linspace(-90, 90, nlat)
    lon = np.linspace(0, 360, nlon)
    dlat = np.deg2rad(lat[1] - lat[0])
    dlon = np.deg2rad(lon[1] - lon[0])
    cosh = np.maximum(np.cos(np.deg2rad(lat)), 0.5)
    dx_min = min(a * dlat, a * dlon * max(1e-3, float(np.min(cosh))))  # ocean.py:293-296
    w2d = np.maximum(np.cos(np.deg2rad(np.meshgrid(lon, lat)[1])), 0.0)
    wsum_all = float(np.sum(w2d) + 1e-15)                              # physics.py:345-346
    if land_mask is not None:
        ocean = (np.asarray(land_mask) == 0)
        wsum_ocean = float(np.sum(w2d * ocean))                        # ocean.py:372-374
        any_ocean = float(bool(np.any(ocean)))
    else:
        wsum_ocean, any_ocean = float(np.sum(w2d)), 1.0
    vals = dict(
        g=p.g, tau_rad=p.tau_rad, gh_newton=p.gh_newton, energy_w=p.energy_w,
        mom_primitive=float(p.mom_scheme == "primitive"),
        sw_a0=p.sw_a0, sw_kc=p.sw_kc, lw_eps0=p.lw_eps0, lw_kc=p.lw_kc, t_floor=p.t_floor, c_sfc=p.c_sfc,
        cloud_couple=float(p.cloud_couple), rh0=p.rh0, k_q=p.k_q, k_p=p.k_p, pcond_ref=opt(p.pcond_ref),
        lw_v2=float(p.lw_v2), hice_ref=p.hice_ref, eps_ocean=p.eps_ocean, eps_land=p.eps_land, eps_ice=p.eps_ice,
        lw_tau0=p.lw_tau0, lw_ktau=p.lw_ktau, gh_lock=float(p.gh_lock), gh_factor_lw=p.gh_factor_lw,
        c_h=p.C_H, cp_air=p.cp_air, seaice=float(p.seaice_enabled), t_freeze=p.t_freeze, rho_i=p.rho_i, l_f=p.L_f,
        cs_ocean=p.Cs_ocean, cs_land=p.Cs_land, cs_ice=p.Cs_ice,
        polar_fix_s=float(p.polar_fix_s), polar_fix_n=float(p.polar_fix_n), atm_h=p.atm_H, diff_factor=p.diff_factor,
        c_e=p.C_E, rho_a=p.rho_a, h_mbl=p.h_mbl, l_v=p.L_v, p0=p.p0,
        evap_ocean=p.ocean_evap_scale, evap_land=p.land_evap_scale, evap_ice=p.ice_evap_scale, tau_cond=p.tau_cond,
        oc_h=p.oc_H, oc_rho_w=p.oc_rho_w, oc_cp_w=p.oc_cp_w, oc_g=p.oc_g, oc_cd=p.oc_CD, oc_r_bot=p.oc_r_bot,
        oc_rho_a=p.oc_rho_a, oc_vcap=p.oc_vcap, oc_tau_scale=p.oc_tau_scale, oc_k_h=p.oc_K_h, oc_cfl=p.oc_cfl,
        oc_max_u=p.oc_max_u, oc_mean4=float(p.oc_outlier == "mean4"), oc_adv_alpha=p.oc_adv_alpha,
        oc_use_qnet=float(p.oc_use_qnet), oc_ice_qfac=p.oc_ice_qfac, oc_eta_cap=p.oc_eta_cap,
        oc_polar_fix=float(p.oc_polar_fix), oc_ts_min=p.oc_ts_min, oc_ts_max=p.oc_ts_max,
        oc_k4_u=opt(p.oc_k4_u), oc_k4_v=opt(p.oc_k4_v), oc_k4_eta=opt(p.oc_k4_eta), oc_dx_min=dx_min,
        d_crit=p.D_crit, k_precip=p.k_precip, alpha_water=p.alpha_water, alpha_ice=p.alpha_ice, alpha_cloud=p.alpha_cloud,
        use_topo_albedo=float(p.use_topo_albedo), orog=float(p.orog_enabled), k_orog=p.k_orog, beta_div=p.beta_div,
        p_fallback=float(p.p_fallback), pq_min=p.pq_min, p_blend=p.p_blend, pref=opt(p.pref), cmax=p.cmax,
        w_mem=wm, w_p=wp, w_src=ws, cloud_floor=p.cloud_floor, cloud_advect=float(p.cloud_advect),
        cloud_adv_alpha=p.cloud_adv_alpha, cloud_smooth_sigma=p.cloud_smooth_sigma,
        lapse_enable=float(p.lapse_enable), lapse_kpm=p.lapse_kpm, land_elev_max=p.land_elev_max,
        polar_ice_thick_max=p.polar_ice_thick_max, rho_snow=p.rho_snow, glacier_frac=p.glacier_frac,
        glacier_swe=p.glacier_swe, has_elevation=float(bool(has_elevation)),
        runoff_tau_days=p.runoff_tau_days, wland_cap=(p.wland_cap if p.wland_cap else 0.0),
        snow_thresh=p.snow_thresh, snow_melt_rate=p.snow_melt_rate, snow_t_band=p.snow_t_band,
        snow_degree_day=float(p.snow_melt_mode == "degree_day"), snow_ddf=p.snow_ddf, snow_melt_tref=p.snow_melt_tref,
        swe_enable=float(p.swe_enable), swe_ref=p.swe_ref, swe_max=(p.swe_max if p.swe_max else 0.0),
        snow_albedo_fresh=p.snow_albedo_fresh,
        eco_enable=float(eco_enable), eco_w_lai=p.eco_lai_albedo_weight, eco_soil_reflect=p.eco_soil_reflect,
        eco_alpha_leaf=float(eco_alpha_leaf),
        wsum_all=wsum_all, oc_wsum_ocean=wsum_ocean, oc_any_ocean=any_ocean,
        oc_inv_rho_h=1.0 / (p.oc_rho_w * p.oc_H), oc_inv_rho_cp_h=1.0 / (p.oc_rho_w * p.oc_cp_w * p.oc_H),
        oc_speed2_cap=speed2_threshold(p.oc_max_u),
    )
    missing = set(P) - set(vals)
    extra = set(vals) - set(P)
    assert not missing and not extra, (missing, extra)
    for k, x in vals.items():
        v[P[k]] = x
    return v


_ENGINES = weakref.WeakValueDictionary()


def engine_for_grid(grid, **kw):
    """The engine shared by every drop-in object built on the same grid object (the script builds
    SpectralModel, WindDrivenSlabOcean, RiverRouting ... on one SphericalGrid)."""
    key = id(grid)
    eng = _ENGINES.get(key)
    if eng is None or eng.shape != (grid.n_lat, grid.n_lon):
        eng = Engine(grid.n_lat, grid.n_lon, **kw)
        _ENGINES[key] = eng
        grid._qd_engine = eng          # keep it alive as long as the grid
    return eng


class Engine:
    def __init__(self, nlat, nlon, batch=1, params: Optional[Sequence[QDParams] | QDParams] = None,
                 dt=300.0, device=None, lib: Optional[Library] = None, band=None):
        self.lib = lib or default_library()
        self.nlat, self.nlon, self.batch = int(nlat), int(nlon), int(batch)
        self.shape = (self.nlat, self.nlon)
        if self.lib.host_emulation:
            self.device = torch.device("cpu")
        else:
            if not torch.cuda.is_available():
                raise RuntimeError("qingdai_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
            self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if params is None:
            params = QDParams.from_env()
        self.params = list(params) if isinstance(params, (list, tuple)) else [params] * self.batch
        assert len(self.params) == self.batch
        self.dt = dt
        self._land = [None] * self.batch
        self._has_elev = False
        self._eco_leaf = 0.0
        self._eco_on = False
        rows, cols, self.dlat, self.dlon, self.lat, self.lon = row_tables(self.nlat, self.nlon, self.params[0], dt)
        self._rows = rows
        self._params_version = 0
        a = const.PLANET_RADIUS
        pv = self._param_block()
        self.fields = torch.zeros((NF, self.batch, self.nlat, self.nlon), dtype=torch.float64, device=self.device)
        self.masks = torch.zeros((NM, self.batch, self.nlat, self.nlon), dtype=torch.uint8, device=self.device)
        ctx = C.c_void_p()
        dev_index = 0 if self.device.type == "cpu" else (self.device.index or 0)
        rc = self.lib.qd_create(self.nlat, self.nlon, self.batch, dev_index, a, self.dlat, self.dlon,
                                a ** 2, self.dlon ** 2, _ptr(rows), _ptr(cols), _ptr(pv), C.byref(ctx))
        if rc != 0:
            raise RuntimeError(f"qd_create failed with status {rc} (no CUDA device / out of memory?)")
        self.ctx = ctx
        self._chk(self.lib.qd_bind(self.ctx, _ptr(self.fields), _ptr(self.masks)), "qd_bind")
        self._check_uniform_switches()
        for b in range(1, self.batch):                                          # K4 / sponge / polar rows are per member
            if self.params[b] is not self.params[0]:
                rb, *_ = row_tables(self.nlat, self.nlon, self.params[b], dt)
                self._chk(self.lib.qd_set_rows_member(self.ctx, b, _ptr(rb)), "qd_set_rows_member")
        if self.device.type == "cuda":
            self._chk(self.lib.qd_set_stream(self.ctx, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), "qd_set_stream")
        self.band = None
        if band is not None and int(band[1]) > 1:
            self._connect_band(*band)
        r1, w1 = gaussian_taps(1.0)
        self._chk(self.lib.qd_set_gauss(self.ctx, 0, r1, 0, _ptr(w1)), "qd_set_gauss")
        self._cloud_sigma = None
        self._upload_cloud_gauss()
        import os
        if os.environ.get("QD_B200_GAUSS2D"):                 # tuning / A-B runs: variant of the fused Gaussian (qd_set_gauss2d)
            self._chk(self.lib.qd_set_gauss2d(self.ctx, int(os.environ["QD_B200_GAUSS2D"])), "qd_set_gauss2d")
        self._finalizer = weakref.finalize(self, self.lib.qd_destroy, self.ctx)

    # ---------------------------------------------------------------- latitude bands (SURVEY 8e, configs[4])
    def _connect_band(self, rank, world, halo_rows=16):
        """One domain over `world` GPUs: rank r computes a block of latitude rows.  The library allocates its
        exchange buffer; the 64-byte IPC handles are swapped here through torch.distributed (setup only --
        the data path is peer-to-peer stores issued by the step's own kernels)."""
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("latitude bands need an initialised torch.distributed process group")
        rank, world = int(rank), int(world)
        self._chk(self.lib.qd_band_init(self.ctx, rank, world, int(halo_rows)), "qd_band_init")
        mine = np.zeros(64, dtype=np.uint8)
        self._chk(self.lib.qd_band_export(self.ctx, _ptr(mine)), "qd_band_export")
        handles = [None] * world
        dist.all_gather_object(handles, mine.tobytes())
        blob = np.frombuffer(b"".join(handles), dtype=np.uint8).copy()
        self._chk(self.lib.qd_band_connect(self.ctx, _ptr(blob)), "qd_band_connect")
        dist.barrier()
        self.band = (rank, world)

    def band_info(self):
        """(own0, own1, halo_rows, error_word) of this rank; the whole grid without bands."""
        a, b, h, e = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._chk(self.lib.qd_band_info(self.ctx, C.byref(a), C.byref(b), C.byref(h), C.byref(e)), "qd_band_info")
        return a.value, b.value, h.value, e.value

    def gather_rows(self, name, member=0):
        """Full field assembled from every rank's own rows (diagnostics / tests; off the hot path)."""
        x = self.get(name, member)
        if self.band is None:
            return x
        import torch.distributed as dist
        r0, r1, _, _ = self.band_info()
        parts = [None] * self.band[1]
        dist.all_gather_object(parts, (r0, r1, x[r0:r1].copy()))
        out = np.empty_like(x)
        for a, b, rows in parts:
            out[a:b] = rows
        return out

    def _upload_cloud_gauss(self):
        """Taps of the cloud tracer's wrap Gaussian (run_simulation.py:1931, QD_CLOUD_SMOOTH_SIGMA); re-sent when the sigma changes."""
        sig = self.params[0].cloud_smooth_sigma
        if sig == self._cloud_sigma:
            return
        rc_, wc = gaussian_taps(sig if sig > 0 else 0.2)
        self._chk(self.lib.qd_set_gauss(self.ctx, 1, rc_, 1, _ptr(wc)), "qd_set_gauss")
        self._cloud_sigma = sig

    # launch structure (which kernels run, cadences) is shared by the members of one batch; every
    # continuous parameter (P vector, K4 / sponge / polar rows) is per member
    _SWITCHES = ("diff_enable", "filter_type", "diff_every", "k4_nsub", "diff_q", "diff_cloud", "shapiro_every", "shapiro_n",
                 "spec_every", "spec_cutoff", "spec_damp", "oc_diff_every", "oc_k4_nsub", "oc_shapiro_n",
                 "oc_shapiro_every", "cloud_smooth_sigma", "k4_q", "k4_c", "cloud_advect", "orog_enabled")

    def _check_uniform_switches(self):
        p0 = self.params[0]
        for p in self.params[1:]:
            for name in self._SWITCHES:
                if getattr(p, name) != getattr(p0, name):
                    raise ValueError(f"ensemble members of one batch must share the launch-structure parameter {name!r}")

    # ---------------------------------------------------------------- plumbing
    def _chk(self, rc, what=""):
        self.lib.check(self.ctx, rc, what)

    def _param_block(self):
        return np.ascontiguousarray(np.stack([
            param_vector(p, self.nlat, self.nlon, self._land[b], self._has_elev, self._eco_leaf, self._eco_on)
            for b, p in enumerate(self.params)]))

    def refresh_params(self, dt=None):
        """Re-snapshot parameters (and K4 rows when dt changed) into the device tables."""
        if dt is not None and dt != self.dt:
            self.dt = dt
        self._params_version = getattr(self, "_params_version", 0) + 1
        self._check_uniform_switches()
        for b, p in enumerate(self.params):                                     # K4 / sponge / polar rows are per member
            rows, *_ = row_tables(self.nlat, self.nlon, p, self.dt)
            if b == 0:
                self._rows = rows
            self._chk(self.lib.qd_set_rows_member(self.ctx, b, _ptr(rows)), "qd_set_rows_member")
        pv = self._param_block()
        self._chk(self.lib.qd_set_params(self.ctx, _ptr(pv)), "qd_set_params")
        self._upload_cloud_gauss()
        p0 = self.params[0]
        for slot, name in ((2, "oc_k4_u"), (3, "oc_k4_v"), (4, "oc_k4_eta")):    # ocean.py:350-352
            if any((getattr(p, name) is None) != (getattr(p0, name) is None) for p in self.params):
                raise ValueError("QD_OCEAN_K4_* overrides must be set for every ensemble member or for none")
            if getattr(p0, name) is not None:
                for b, p in enumerate(self.params):
                    r = np.full(self.nlat, float(getattr(p, name)))
                    self._chk(self.lib.qd_user_row_member(self.ctx, slot, b, _ptr(r)), "qd_user_row_member")

    def set_params(self, params):
        self.params = list(params) if isinstance(params, (list, tuple)) else [params] * self.batch
        self.refresh_params()

    def use_graphs(self, enable=True):
        """Ocean sub-step loop as a CUDA-graph WHILE node (default) or as a host loop with a read-back."""
        self._chk(self.lib.qd_use_graphs(self.ctx, int(enable) if not isinstance(enable, bool) else (2 if enable else 0)), "qd_use_graphs")

    def graph_status(self):
        """{"live": cached CUDA graphs, "failed": captures that fell back to stream mode} (qd_graph_status)."""
        a, b = C.c_int(), C.c_int()
        self._chk(self.lib.qd_graph_status(self.ctx, C.byref(a), C.byref(b)), "qd_graph_status")
        return {"live": a.value, "failed": b.value}

    def sync(self):
        self._chk(self.lib.qd_synchronize(self.ctx), "qd_synchronize")

    def launches(self):
        n = C.c_longlong()
        self._chk(self.lib.qd_launch_count(self.ctx, C.byref(n)), "qd_launch_count")
        return int(n.value)

    # ---------------------------------------------------------------- field access
    def tensor(self, name):
        """Device view [B, nlat, nlon] of a float64 field."""
        return self.fields[F[name]]

    def get(self, name, member=0):
        """Host copy of one member's field (device -> host through the C ABI)."""
        out = np.empty(self.shape, dtype=np.float64)
        self._chk(self.lib.qd_download_field(self.ctx, F[name], member, _ptr(out)), "qd_download_field")
        return out

    def set(self, name, value, member=None):
        """Upload a host array into a field (all members when member is None)."""
        arr = np.ascontiguousarray(np.broadcast_to(np.asarray(value, dtype=np.float64), self.shape))
        members = range(self.batch) if member is None else [member]
        for b in members:
            self._chk(self.lib.qd_upload_field(self.ctx, F[name], b, _ptr(arr)), "qd_upload_field")

    def get_mask(self, name, member=0):
        out = np.empty(self.shape, dtype=np.uint8)
        self._chk(self.lib.qd_download_mask(self.ctx, M[name], member, _ptr(out)), "qd_download_mask")
        return out

    def set_mask(self, name, value, member=None):
        arr = np.ascontiguousarray(np.asarray(value).astype(np.uint8))
        members = range(self.batch) if member is None else [member]
        for b in members:
            self._chk(self.lib.qd_upload_mask(self.ctx, M[name], b, _ptr(arr)), "qd_upload_mask")
            if name == "land":
                self._land[b] = np.array(arr, copy=True)
        if name == "land":
            self.refresh_params()

    def set_elevation(self, elevation, member=None):
        """Elevation map + the static slope unit vectors of physics.py:134-152."""
        self._has_elev = elevation is not None
        if elevation is None:
            self.refresh_params()
            return
        elev = np.asarray(elevation, dtype=np.float64)
        a = const.PLANET_RADIUS
        cos_lat = np.maximum(np.cos(np.deg2rad(np.meshgrid(self.lon, self.lat)[1])), 1e-6)
        dx = a * cos_lat * self.dlon
        dy = a * self.dlat
        dHdx = (np.roll(elev, -1, axis=1) - np.roll(elev, 1, axis=1)) / (2.0 * dx)
        dHdy = (np.roll(elev, -1, axis=0) - np.roll(elev, 1, axis=0)) / (2.0 * dy)
        dHdy[0, :] = 0.0
        dHdy[-1, :] = 0.0
        gn = np.sqrt(dHdx ** 2 + dHdy ** 2)
        eps = 1e-12
        nx = np.where(gn > eps, dHdx / (gn + eps), 0.0)
        ny = np.where(gn > eps, dHdy / (gn + eps), 0.0)
        self.set("elevation", elev, member)
        self.set("orog_nx", nx, member)
        self.set("orog_ny", ny, member)
        self.refresh_params()

    def bind_eco(self, lai_dev, n_layers, k_canopy, every_hours, lai_delta, substep_every=1):
        """Hand the LAI layers [B, S*K, nlat, nlon] (device tensor, kept alive here) to the library."""
        self._lai_dev = lai_dev
        ptr = _ptr(lai_dev) if lai_dev is not None else C.c_void_p(0)
        self._chk(self.lib.qd_eco_bind(self.ctx, ptr, int(n_layers), float(k_canopy), float(every_hours), float(lai_delta), int(substep_every)), "qd_eco_bind")

    def eco_reset(self, hours, next_hours, cached=False, step_count=0):
        self._chk(self.lib.qd_eco_reset(self.ctx, float(hours), float(next_hours), int(bool(cached)), int(step_count)), "qd_eco_reset")

    def set_eco(self, enabled, alpha_leaf_scalar=0.0):
        self._eco_on, self._eco_leaf = bool(enabled), float(alpha_leaf_scalar)
        self.refresh_params()

    DIAG = ("wsum", "ts", "h", "q", "cloud", "hice", "wland", "ssnow", "eflux", "precip", "rland", "albedo", "sst", "I", "R", "OLR",
            "SW_sfc", "LW_sfc", "SH", "LH", "KE_ocean", "eta_min", "eta_max", "uocean_max", "uabs_max", "ts_min", "ts_max")

    def diag(self):
        """One launch of the device diagnostics kernel (csrc/qd_diag.cuh): list of dicts, one per member.  Weighted
        sums come back as the reference's area-weighted MEANS, sum(x w) / (sum(w) + 1e-15) (energy.py:522-527)."""
        n = int(self.lib.qd_diag_count())
        assert n == len(self.DIAG)
        raw = np.empty((self.batch, n), dtype=np.float64)
        self._chk(self.lib.qd_diag(self.ctx, _ptr(raw)), "qd_diag")
        raw[:, 1:21] /= (raw[:, :1] + 1e-15)                 # x / (sum(w) + 1e-15), element by element as before
        return [dict(zip(self.DIAG, row)) for row in raw.tolist()]

    def median_stats(self):
        """[[launches, first-digit speculation hits]] of the exact-median kernel per call site (positive precipitation
        part, precipitation, P_cond, qd_median_pos)."""
        out = np.zeros((4, 2), dtype=np.int64)
        self._chk(self.lib.qd_median_stats(self.ctx, _ptr(out)), "qd_median_stats")
        return out

    def scalars(self):
        out = np.empty((self.batch, NS), dtype=np.float64)
        self._chk(self.lib.qd_get_scalars(self.ctx, _ptr(out)), "qd_get_scalars")
        return out

    # ---------------------------------------------------------------- step configuration
    def step_cfg(self, dt, has_albedo=False, with_ocean=True, with_hydrology=True, with_routing=False,
                 with_eco=False, loop_with_albedo=False, oc_has_q=True, oc_has_ice=True, store_isr_ab=True):
        self._check_uniform_switches()        # cadences / filter switches come from member 0: a batch that disagrees must not run
        p = self.params[0]
        rows = self._rows
        c = StepCfg()
        c.dt = float(dt)
        c.has_albedo = int(has_albedo)
        c.diff_enable = int(p.diff_enable and p.filter_type in ("hyper4", "combo"))
        c.diff_every = max(1, int(p.diff_every))
        c.k4_nsub = int(p.k4_nsub)
        c.apply_q = int(bool(np.any(rows[R["k4_q"]] > 0.0)) or p.diff_q)            # dynamics.py:589-590
        c.apply_cloud = int(bool(np.any(rows[R["k4_c"]] > 0.0)) or p.diff_cloud)
        sh_on = p.filter_type in ("shapiro", "combo", "hyper4") and p.shapiro_every > 0
        c.shapiro_every = int(p.shapiro_every) if sh_on else 0
        c.shapiro_n = int(p.shapiro_n)
        c.shapiro_q = int(p.diff_q)
        c.shapiro_cloud = int(p.diff_cloud)
        c.spec_every = int(p.spec_every) if (p.filter_type in ("spectral", "combo") and p.spec_every > 0) else 0
        c.spec_cutoff, c.spec_damp = float(p.spec_cutoff), float(p.spec_damp)
        c.oc_diff_every, c.oc_k4_nsub = int(p.oc_diff_every), int(p.oc_k4_nsub)
        c.oc_shapiro_n, c.oc_shapiro_every = int(p.oc_shapiro_n), int(p.oc_shapiro_every)
        c.oc_has_q, c.oc_has_ice = int(oc_has_q), int(oc_has_ice)
        c.with_ocean, c.with_hydrology = int(with_ocean), int(with_hydrology)
        c.with_routing, c.with_eco = int(with_routing), int(with_eco)
        c.loop_with_albedo, c.store_isr_ab = int(loop_with_albedo), int(store_isr_ab)
        return c

    def _dt_guard(self, dt):
        if dt != self.dt:
            self.refresh_params(dt)

    def set_counters(self, atm, ocean, has_cloud_eff):
        self._chk(self.lib.qd_set_counters(self.ctx, int(atm), int(ocean), int(has_cloud_eff)), "qd_set_counters")

    def counters(self):
        a, o, c = C.c_int(), C.c_int(), C.c_int()
        self._chk(self.lib.qd_get_counters(self.ctx, C.byref(a), C.byref(o), C.byref(c)), "qd_get_counters")
        return a.value, o.value, c.value

    # ---------------------------------------------------------------- steps
    def atmos_step(self, dt, has_albedo=False):
        self._dt_guard(dt)
        cfg = self.step_cfg(dt, has_albedo=has_albedo)
        self._chk(self.lib.qd_atmos_step(self.ctx, C.byref(cfg)), "qd_atmos_step")

    def ocean_step(self, dt, has_q=True, has_ice=True, winds=None):
        """One WindDrivenSlabOcean.step; ``winds`` = (u, v) device tensors [B, nlat, nlon] to drive it with instead of the
        atmosphere's own wind fields."""
        self._dt_guard(dt)
        cfg = self.step_cfg(dt, oc_has_q=has_q, oc_has_ice=has_ice)
        if winds is None:
            self._chk(self.lib.qd_ocean_step(self.ctx, C.byref(cfg)), "qd_ocean_step")
        else:
            self._chk(self.lib.qd_ocean_step_winds(self.ctx, C.byref(cfg), _ptr(winds[0]), _ptr(winds[1])), "qd_ocean_step_winds")

    def loop_steps(self, forcings: Sequence[Forcing], dt, **cfg_kw):
        """Run len(forcings) fused loop steps (run_simulation.py:1760-2344 without plotting/daily ecology)."""
        self._dt_guard(dt)
        p = self.params[0]
        key = (float(dt), self._params_version, tuple(sorted(cfg_kw.items())),
               p.diff_enable, p.filter_type, p.diff_every, p.k4_nsub, p.diff_q, p.diff_cloud, p.shapiro_every, p.shapiro_n,
               p.spec_every, p.spec_cutoff, p.spec_damp, p.oc_diff_every, p.oc_k4_nsub, p.oc_shapiro_n, p.oc_shapiro_every)
        if getattr(self, "_loop_cfg_key", None) != key:          # the step configuration only changes with dt / params / switches
            self._loop_cfg, self._loop_cfg_key = self.step_cfg(dt, **cfg_kw), key
        cfg = self._loop_cfg
        n = len(forcings)
        if n == 1:
            arr = self.__dict__.get("_forcing1")
            if arr is None:
                arr = self._forcing1 = (Forcing * 1)()
            arr[0] = forcings[0]
        else:
            arr = (Forcing * n)(*forcings)
        self._chk(self.lib.qd_loop_step(self.ctx, C.byref(cfg), arr, n), "qd_loop_step")

    def last_nsub(self):
        out = np.zeros(self.batch, dtype=np.int32)
        self._chk(self.lib.qd_last_nsub(self.ctx, _ptr(out)), "qd_last_nsub")
        return out

    # ---------------------------------------------------------------- operator seam (NumPy in / NumPy out)
    def _stage(self, k, arr=None):
        t = self.fields[F[f"x{k}"]][0]
        if arr is not None:
            t.copy_(torch.from_numpy(np.ascontiguousarray(np.asarray(arr, dtype=np.float64))))
        return t

    def op_divvort(self, u, v, vort=False):
        assert self.batch == 1
        tu, tv, to = self._stage(5, u), self._stage(6, v), self._stage(7)
        fn = self.lib.qd_vorticity if vort else self.lib.qd_divergence
        self._chk(fn(self.ctx, _ptr(tu), _ptr(tv), _ptr(to)), "qd_divergence")
        self.sync()
        return to.cpu().numpy().copy()

    def _user_row(self, slot, rows):
        r = np.ascontiguousarray(np.asarray(rows, dtype=np.float64).reshape(-1))
        assert r.size == self.nlat, "row table must have n_lat entries"
        ptr = self.lib.qd_user_row(self.ctx, slot, _ptr(r))
        if not ptr:
            raise RuntimeError("qd_user_row failed")
        return C.c_void_p(ptr)

    @staticmethod
    def _rows_of(coslat, nlat):
        c = np.asarray(coslat, dtype=np.float64)
        return c[:, 0] if c.ndim == 2 else c.reshape(nlat)

    def op_laplacian(self, Fh, cos_rows):
        """jax_compat.laplacian_sphere (jax_compat.py:111-132) through the host-buffer C entry point."""
        Fh = np.ascontiguousarray(np.asarray(Fh, dtype=np.float64))
        out = np.empty_like(Fh)
        cr = np.ascontiguousarray(self._rows_of(cos_rows, self.nlat))
        self._chk(self.lib.qd_laplacian_host(self.ctx, _ptr(Fh), _ptr(out), _ptr(cr)), "qd_laplacian_host")
        return out

    def op_hyperdiffuse(self, Fh, k4, dt, nsub, cos_rows):
        """jax_compat.hyperdiffuse (jax_compat.py:135-187); k4 scalar or a latitude-only map."""
        Fh = np.ascontiguousarray(np.asarray(Fh, dtype=np.float64))
        out = np.empty_like(Fh)
        cr = np.ascontiguousarray(self._rows_of(cos_rows, self.nlat))
        if np.isscalar(k4):
            kmap, ks = None, float(k4)
        else:
            kmap = np.ascontiguousarray(np.broadcast_to(np.nan_to_num(np.asarray(k4, dtype=np.float64)), self.shape))
            ks = 0.0
        self._chk(self.lib.qd_hyperdiffuse_host(self.ctx, _ptr(Fh), _ptr(out), _ptr(kmap) if kmap is not None else None,
                                                ks, float(dt), int(nsub), _ptr(cr)), "qd_hyperdiffuse_host")
        return out

    def op_advect(self, Fh, u, v, dt, cos_rows):
        """jax_compat.advect_semilag (jax_compat.py:190-216); cos_rows already floored by the caller."""
        Fh, u, v = (np.ascontiguousarray(np.asarray(x, dtype=np.float64)) for x in (Fh, u, v))
        out = np.empty_like(Fh)
        cr = np.ascontiguousarray(self._rows_of(cos_rows, self.nlat))
        self._chk(self.lib.qd_advect_host(self.ctx, _ptr(Fh), _ptr(u), _ptr(v), _ptr(out), float(dt), _ptr(cr)), "qd_advect_host")
        return out

    def op_shapiro(self, Fh, n=2):
        assert self.batch == 1
        t, s = self._stage(5, Fh), self._stage(6)
        self._chk(self.lib.qd_shapiro(self.ctx, _ptr(t), _ptr(s), int(n)), "qd_shapiro")
        self.sync()
        return t.cpu().numpy().copy()

    def op_gaussian(self, Fh, sigma, mode="reflect"):
        assert self.batch == 1
        r, w = gaussian_taps(sigma)
        t, s = self._stage(5, Fh), self._stage(6)
        self._chk(self.lib.qd_gaussian(self.ctx, _ptr(t), _ptr(s), r, int(mode == "wrap"), _ptr(w)), "qd_gaussian")
        self.sync()
        return t.cpu().numpy().copy()

    def op_bandstop(self, Fh, cutoff=0.75, damp=0.5):
        assert self.batch == 1
        t = self._stage(5, Fh)
        self._chk(self.lib.qd_zonal_bandstop(self.ctx, _ptr(t), float(cutoff), float(damp)), "qd_zonal_bandstop")
        self.sync()
        return t.cpu().numpy().copy()

    def op_median_pos(self, Fh, empty=float("nan")):
        assert self.batch == 1
        t = self._stage(5, Fh)
        out = np.zeros(1)
        self._chk(self.lib.qd_median_pos(self.ctx, _ptr(t), float(empty), _ptr(out)), "qd_median_pos")
        return float(out[0])

    def op_wsum(self, Fh):
        assert self.batch == 1
        t = self._stage(5, Fh)
        out = np.zeros(1)
        self._chk(self.lib.qd_wsum(self.ctx, _ptr(t), _ptr(out)), "qd_wsum")
        return float(out[0])
