"""Atmosphere / ocean / loop-body steps of the Qingdai GCM, restated in NumPy (oracle; test-only).

Structure is deliberately different from the reference (free functions over a
``State`` bag and one parameter object ``p`` -- the same attribute names as
``qingdai_b200.params.QDParams``), but every arithmetic expression keeps the
reference's operand order so results are bit-comparable where libm allows.

References (relative to the reference checkout):
  atmos_step  -> pygcm/dynamics.py:260-667 (+ humidity.py, energy.py)
  ocean_step  -> pygcm/ocean.py:265-533
  loop_step   -> scripts/run_simulation.py:1760-2344 (per-step physics around the two cores)
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np

from . import ops

SIGMA = 5.670374e-8            # constants.py:10
PLANET_RADIUS = 6.371e6        # constants.py:31
PLANET_OMEGA = 8.726646259971648e-5   # constants.py:33
EPSILON = 0.622                # humidity.py:34


# =========================================================================== grid
def make_grid(nlat, nlon):
    """grid.py:10-39 + the derived per-row metrics every operator needs."""
    g = SimpleNamespace()
    g.nlat, g.nlon = int(nlat), int(nlon)
    g.lat = np.linspace(-90, 90, nlat)
    g.lon = np.linspace(0, 360, nlon)
    g.dlat = float(np.deg2rad(g.lat[1] - g.lat[0]))
    g.dlon = float(np.deg2rad(g.lon[1] - g.lon[0]))
    g.lat_rad = np.deg2rad(g.lat)
    g.lon_rad = np.deg2rad(g.lon)
    g.cos = np.cos(g.lat_rad)
    g.sin = np.sin(g.lat_rad)
    g.f = 2 * PLANET_OMEGA * np.sin(g.lat_rad)          # grid.py:95
    g.w = np.maximum(g.cos, 0.0)                        # area weights
    g.a = PLANET_RADIUS
    return g


def _col(x):
    return np.asarray(x, dtype=np.float64).reshape(-1, 1)


# =========================================================================== humidity / energy cell physics
def q_sat(T, p0):
    """humidity.py:85-101 (Tetens)."""
    Tc = np.clip(np.asarray(T, dtype=np.float64) - 273.15, -80.0, 60.0)
    es = 610.94 * np.exp(17.625 * Tc / (Tc + 243.04))
    den = np.maximum(p0 - (1.0 - EPSILON) * es, 1.0)
    return np.clip(EPSILON * es / den, 0.0, 0.5)


def evap_factor(land_mask, h_ice, p):
    """humidity.py:116-142."""
    land = (land_mask == 1)
    ocean = ~land
    ice = (h_ice > 1e-6) & ocean
    fac = np.zeros(land_mask.shape, dtype=np.float64)
    fac[ice] = p.ice_evap_scale
    fac[ocean & ~ice] = p.ocean_evap_scale
    fac[land] = p.land_evap_scale
    return fac


def shortwave(I, albedo, cloud, p):
    """energy.py:77-98."""
    alpha = np.clip(albedo, 0.0, 1.0)
    Ic = np.maximum(0.0, I)
    R = Ic * alpha
    A = np.clip(p.sw_a0 + p.sw_kc * np.clip(cloud, 0.0, 1.0), 0.0, 0.95)
    SW_atm = Ic * A
    SW_sfc = np.maximum(0.0, Ic - R - SW_atm)
    return SW_atm, SW_sfc, R


def emissivity_map(land_mask, ice_frac, p):
    """energy.py:141-158."""
    land = (land_mask == 1)
    ocean = ~land
    eps = np.full(np.shape(ice_frac), p.eps_land, dtype=np.float64)
    fi = np.clip(ice_frac, 0.0, 1.0)
    eps[ocean] = ((1.0 - fi) * p.eps_ocean + fi * p.eps_ice)[ocean]
    return ops.nan_to_num(eps)


def longwave_v2(Ts, Ta, cloud_eff, eps_sfc, p):
    """energy.py:161-234."""
    Ts = np.maximum(0.0, Ts)
    Ta = np.maximum(0.0, Ta)
    Ts4 = Ts ** 4
    Ta4 = Ta ** 4
    eps_clear = float(np.clip(p.lw_eps0, 0.0, 1.0))
    ce = np.clip(cloud_eff, 0.0, 1.0)
    eps_cloud = np.clip(1.0 - np.exp(-p.lw_ktau * (p.lw_tau0 * ce)), 0.0, 1.0)
    eps_eff = 1.0 - (1.0 - eps_clear) * (1.0 - eps_cloud)
    if np.isscalar(eps_sfc):
        es = np.full_like(Ts, float(eps_sfc))
    else:
        es = np.clip(ops.nan_to_num(eps_sfc), 0.0, 1.0)
    OLR = eps_eff * SIGMA * Ta4 + (1.0 - eps_eff) * SIGMA * es * Ts4
    DLR = eps_eff * SIGMA * Ta4
    LW_sfc = DLR - SIGMA * es * Ts4
    LW_atm = eps_eff * (SIGMA * es * Ts4 - 2.0 * SIGMA * Ta4)
    if p.gh_lock:
        g = p.gh_factor_lw
        OLR = (1.0 - g) * SIGMA * Ts4
        DLR = g * SIGMA * Ts4
        LW_sfc = DLR - SIGMA * es * Ts4
    return LW_atm, LW_sfc, OLR, DLR, eps_eff


def longwave_v1(Ts, Ta, cloud, p):
    """energy.py:101-137."""
    Ts4 = np.maximum(0.0, Ts) ** 4
    Ta4 = np.maximum(0.0, Ta) ** 4
    eps = np.clip(p.lw_eps0 + p.lw_kc * np.clip(cloud, 0.0, 1.0), 0.0, 1.0)
    OLR = eps * SIGMA * Ta4 + (1.0 - eps) * SIGMA * Ts4
    DLR = eps * SIGMA * Ta4
    LW_sfc = DLR - SIGMA * Ts4
    LW_atm = eps * (SIGMA * Ts4 - 2.0 * SIGMA * Ta4)
    if p.gh_lock:
        g = p.gh_factor_lw
        OLR = (1.0 - g) * SIGMA * Ts4
        DLR = g * SIGMA * Ts4
        LW_sfc = DLR - SIGMA * Ts4
    return LW_atm, LW_sfc, OLR, DLR, eps


def sensible_heat(Ts, Ta, u, v, p):
    """energy.py:423-449 (the Bowen-ratio LH it also returns is discarded by every caller)."""
    V = np.sqrt(u ** 2 + v ** 2)
    return p.rho_a * p.cp_air * p.C_H * V * (Ts - Ta)


def seaice_integrate(Ts, SW_sfc, LW_sfc, SH, LH, dt, land_mask, h_ice, p):
    """energy.py:291-420 (melt first, freeze near t_freeze, residual heating, polar fix, clamps)."""
    Q = SW_sfc - LW_sfc - SH - LH
    land = (land_mask == 1)
    ocean = ~land
    Ts_n = np.array(Ts, dtype=np.float64, copy=True)
    hi = np.array(h_ice, dtype=np.float64, copy=True)
    rl = p.rho_i * p.L_f
    melt = (hi > 0.0) & ocean & (Q > 0.0)
    if np.any(melt):
        dh = np.minimum((Q[melt] * dt) / rl, hi[melt])
        hi[melt] -= dh
        Q[melt] = Q[melt] - (dh * p.rho_i * p.L_f) / dt
    frz = ocean & (Q < 0.0) & (Ts_n <= (p.t_freeze + 0.5))
    if np.any(frz):
        hi[frz] += (-Q[frz] * dt) / rl
        Q[frz] = 0.0
        Ts_n[frz] = np.minimum(Ts_n[frz], p.t_freeze)
    Cs = np.where(land, p.Cs_land, np.where(hi > 0.0, p.Cs_ice, p.Cs_ocean))
    Cs = np.where(np.isfinite(Cs) & (Cs > 1e3), Cs, 1e3)
    Ts_n = Ts_n + (Q / Cs) * dt
    for on, row in ((p.polar_fix_s, 0), (p.polar_fix_n, -1)):
        if on:
            m = ocean[row] & (Q[row] < 0.0) & (Ts_n[row] > p.t_freeze)
            Ts_n[row, m] = p.t_freeze
    Ts_n = np.where((hi > 0.0) & ocean, np.minimum(Ts_n, p.t_freeze), Ts_n)
    Ts_n = np.maximum(p.t_floor, Ts_n)
    return ops.nan_to_num(Ts_n), ops.nan_to_num(hi)


# =========================================================================== atmosphere
def k4_rows(g, p, dt):
    """dynamics.py:557-570: latitude-adaptive K4 rows for u,v,h,q,cloud (env scalars override)."""
    cosr = np.maximum(np.cos(g.lat_rad), 1e-3)
    dx_min = np.minimum(g.a * g.dlat, g.a * g.dlon * cosr)
    base = p.sigma4 * (dx_min ** 4) / max(1e-12, dt)
    out = {}
    for name, scale, ov in (("u", None, p.k4_u), ("v", None, p.k4_v), ("h", 0.5, p.k4_h),
                            ("q", 0.5, p.k4_q), ("c", 0.25, p.k4_c)):
        if ov is not None:
            out[name] = np.full(g.nlat, float(ov))
        else:
            out[name] = base if scale is None else scale * base
    return out


def atmos_step(st, g, p, Teq, dt, albedo=None):
    """SpectralModel.time_step (dynamics.py:260-667).  Mutates and returns ``st``."""
    a, dlat, dlon = g.a, g.dlat, g.dlon
    cos_adv = np.maximum(1e-6, g.cos)           # dynamics.py:104
    cos_lap = np.maximum(g.cos, 0.2)            # dynamics.py:164
    T_a = 288.0 + (p.g / 1004.0) * st.h

    # -- humidity block (dynamics.py:282-297) --------------------------------
    fac = evap_factor(st.land_mask, st.h_ice, p)
    V = np.sqrt(st.u ** 2 + st.v ** 2)
    deficit = np.maximum(0.0, q_sat(st.T_s, p.p0) - st.q)
    E = ops.nan_to_num(p.rho_a * p.C_E * V * deficit * fac)
    LH = p.L_v * E
    M_col = max(1e-6, float(p.rho_a * p.h_mbl))
    q_evap = st.q + (E / M_col) * dt
    excess = np.maximum(0.0, q_evap - q_sat(T_a, p.p0))
    P_cond = (excess / max(1e-6, float(p.tau_cond))) * M_col
    q_next = q_evap - (P_cond / M_col) * dt
    q_next = np.clip(ops.nan_to_num(q_next), 0.0, 0.5)
    P_cond = ops.nan_to_num(P_cond)
    LH_rel = p.L_v * P_cond
    st.q = np.clip(ops.nan_to_num(q_next), 0.0, 0.5)
    st.E_flux, st.P_cond, st.LH, st.LH_release = E, P_cond, LH, LH_rel

    # -- Newtonian surface update (dynamics.py:304-322) ----------------------
    olr_old = SIGMA * st.T_s ** 4
    net_old = SIGMA * Teq ** 4 + p.gh_newton * SIGMA * T_a ** 4 - olr_old
    Ts_newton = st.T_s + (net_old / max(1e-12, p.c_sfc)) * dt

    Ts_energy = None
    h_ice_next = None
    SW_atm = LW_atm = SH = None
    if albedo is not None:
        if p.cloud_couple:
            RH = np.clip(st.q / np.maximum(1e-12, q_sat(T_a, p.p0)), 0.0, 1.5)
            rh_ex = np.maximum(0.0, RH - p.rh0)
            P = st.P_cond
            if p.pcond_ref is not None:
                P_ref = float(p.pcond_ref)
            else:
                P_ref = ops.median_pos(P, empty=1e-6)
            p_term = np.tanh(np.where(P_ref > 0, P / P_ref, 0.0))
            cloud_eff = np.clip(st.cloud + p.k_q * rh_ex + p.k_p * p_term, 0.0, 1.0)
        else:
            cloud_eff = st.cloud
        st.cloud_eff = cloud_eff
        SW_atm, SW_sfc, R = shortwave(st.isr, albedo, cloud_eff, p)
        if p.lw_v2:
            ice_frac = 1.0 - np.exp(-np.maximum(st.h_ice, 0.0) / max(1e-6, p.hice_ref))
            eps_sfc = emissivity_map(st.land_mask, ice_frac, p)
            LW_atm, LW_sfc, OLR, DLR, _ = longwave_v2(st.T_s, T_a, cloud_eff, eps_sfc, p)
        else:
            LW_atm, LW_sfc, OLR, DLR, _ = longwave_v1(st.T_s, T_a, cloud_eff, p)
        SH = sensible_heat(st.T_s, T_a, st.u, st.v, p)
        if p.seaice_enabled:
            Ts_energy, h_ice_next = seaice_integrate(st.T_s, SW_sfc, LW_sfc, SH, LH, dt,
                                                     st.land_mask, st.h_ice, p)
        else:
            net = SW_sfc - LW_sfc - SH - LH
            Cs = np.where(np.isfinite(st.C_s_map) & (st.C_s_map > 1e3), st.C_s_map, 1e3)
            Ts_energy = ops.nan_to_num(np.maximum(p.t_floor, st.T_s + (net / Cs) * dt))
        st.olr = OLR
        st.diag_energy = dict(R=R, OLR=OLR, SW_sfc=SW_sfc, LW_sfc=LW_sfc, SH=SH, LH=LH,
                              SW_atm=SW_atm, LW_atm=LW_atm)
    else:
        st.olr = olr_old

    w = min(1.0, max(0.0, float(p.energy_w)))
    if Ts_energy is None:
        st.T_s = Ts_newton
    else:
        st.T_s = (1.0 - w) * Ts_newton + w * Ts_energy
        if p.seaice_enabled and h_ice_next is not None:
            st.h_ice = h_ice_next

    st.step_counter += 1                                    # dynamics.py:451 (before cadence tests)

    # -- semi-Lagrangian advection of Ts and q (dynamics.py:454-461) ---------
    adv = ops.advect_semilag(st.T_s, st.u, st.v, dt, a, dlat, dlon, cos_adv)
    st.T_s = (1.0 - 0.2) * st.T_s + 0.2 * adv
    advq = ops.advect_semilag(st.q, st.u, st.v, dt, a, dlat, dlon, cos_adv)
    st.q = (1.0 - 0.2) * st.q + 0.2 * advq
    st.q = np.clip(ops.nan_to_num(st.q), 0.0, 0.5)

    # -- radiative relaxation of h (dynamics.py:464-467) ---------------------
    h_eq = (287 / p.g) * Teq
    st.h = st.h + ((h_eq - st.h) / p.tau_rad) * dt

    # -- M3 atmospheric energy coupling (dynamics.py:470-478; energy.py:452-491)
    if (albedo is not None) and (float(p.energy_w) > 0.0):
        F_atm = SW_atm + LW_atm + SH + LH_rel
        denom = max(1e-6, float(p.rho_a)) * max(1.0, float(p.atm_H)) * float(p.g)
        st.h = ops.nan_to_num(st.h + float(p.energy_w) * (F_atm / denom) * dt)

    # -- momentum (dynamics.py:484-530) --------------------------------------
    dh_dlon = ops.grad_cols(st.h, dlon)
    dh_dlat = ops.grad_rows(st.h, dlat)
    cosc = _col(np.maximum(g.cos, 1e-6))
    f = _col(g.f)
    if p.mom_scheme == "primitive":
        u0, v0 = st.u.copy(), st.v.copy()
        PGx = -(p.g / (a * cosc)) * dh_dlon
        PGy = -(p.g / a) * dh_dlat
        st.u = np.clip(u0 + (PGx + f * v0 - st.friction * u0) * dt, -200.0, 200.0)
        st.v = np.clip(v0 + (PGy - f * u0 - st.friction * v0) * dt, -200.0, 200.0)
    else:
        f_min = 2.0 * PLANET_OMEGA * np.sin(np.deg2rad(5.0))
        sgn = np.where(f >= 0.0, 1.0, -1.0)
        f_safe = np.where(np.abs(f) < f_min, sgn * f_min, f)
        u_g = np.clip(-(p.g / (f_safe * a * cosc)) * dh_dlat, -200.0, 200.0)
        v_g = np.clip((p.g / (f_safe * a)) * dh_dlon, -200.0, 200.0)
        st.u = st.u * 0.8 + u_g * 0.2
        st.v = st.v * 0.8 + v_g * 0.2
        st.u = st.u + (-st.friction * st.u) * dt
        st.v = st.v + (-st.friction * st.v) * dt

    # -- hyperdiffusion (dynamics.py:542-594) --------------------------------
    sc = st.step_counter
    if p.diff_enable and p.filter_type in ("hyper4", "combo") and (sc % max(1, p.diff_every) == 0):
        k4 = k4_rows(g, p, dt)
        hd = lambda F, k, n: ops.hyperdiffuse(F, _col(k), dt, n, dlat, dlon, cos_lap, a)
        st.u = hd(st.u, k4["u"], p.k4_nsub)
        st.v = hd(st.v, k4["v"], p.k4_nsub)
        st.h = hd(st.h, k4["h"], p.k4_nsub)
        if np.any(k4["q"] > 0.0) or p.diff_q:
            st.q = hd(st.q, k4["q"], 1)
        if np.any(k4["c"] > 0.0) or p.diff_cloud:
            st.cloud = hd(st.cloud, k4["c"], 1)

    # -- Shapiro / zonal band-stop (dynamics.py:611-637) ---------------------
    if p.filter_type in ("shapiro", "combo", "hyper4") and p.shapiro_every > 0 and (sc % p.shapiro_every == 0):
        st.u = ops.shapiro(st.u, p.shapiro_n)
        st.v = ops.shapiro(st.v, p.shapiro_n)
        st.h = ops.shapiro(st.h, p.shapiro_n)
        if p.diff_q:
            st.q = ops.shapiro(st.q, max(1, p.shapiro_n - 1))
        if p.diff_cloud:
            st.cloud = ops.shapiro(st.cloud, max(1, p.shapiro_n - 1))
    if p.filter_type in ("spectral", "combo") and p.spec_every > 0 and (sc % p.spec_every == 0):
        st.u = ops.zonal_bandstop(st.u, p.spec_cutoff, p.spec_damp)
        st.v = ops.zonal_bandstop(st.v, p.spec_cutoff, p.spec_damp)
        st.h = ops.zonal_bandstop(st.h, p.spec_cutoff, p.spec_damp)

    # -- cloud tail (dynamics.py:641-667) ------------------------------------
    st.cloud = ops.advect_semilag(st.cloud, st.u, st.v, dt, a, dlat, dlon, cos_adv)
    st.cloud = st.cloud * (1 - dt / (2.0 * 24 * 3600))
    for name in ("u", "v", "h", "cloud", "q"):
        setattr(st, name, getattr(st, name) * p.diff_factor)
    for name in ("u", "v", "h", "T_s", "cloud", "q"):
        setattr(st, name, ops.nan_to_num(getattr(st, name)))
    return st


# =========================================================================== ocean
def ocean_nsub(oc, g, p, dt, u_atm, v_atm):
    """ocean.py:285-303: wind stress inputs and the data-dependent sub-step count."""
    u_rel = u_atm - oc.uo
    v_rel = v_atm - oc.vo
    Va = np.sqrt(u_rel ** 2 + v_rel ** 2)
    cosr = np.maximum(g.cos, 0.5)
    dx_min = min(g.a * g.dlat, g.a * g.dlon * max(1e-3, float(np.min(cosr))))
    c = np.sqrt(p.oc_g * p.oc_H)
    uadv = max(float(np.max(np.sqrt(oc.uo ** 2 + oc.vo ** 2))), float(np.max(Va)))
    n = int(np.ceil(max(c, uadv) * (dt / max(1e-12, dx_min)) / max(1e-3, p.oc_cfl)))
    return int(max(1, min(500, n)))


def polar_fill(oc, g):
    """ocean.py:197-262 (scalar ring mean of Ts; tangent-plane vector mean of uo,vo)."""
    ocean = (oc.land_mask == 0)
    lam = np.deg2rad(g.lon)
    for row, pole in ((0, "south"), (-1, "north")):
        m = ocean[row]
        if np.any(m):
            oc.Ts[row, m] = float(np.mean(oc.Ts[row, m]))
    for row, pole in ((0, "south"), (-1, "north")):
        m = ocean[row]
        if not np.any(m):
            continue

        def basis(l):
            ee = np.stack([-np.sin(l), np.cos(l), np.zeros_like(l)], axis=1)
            sgn = -1.0 if pole == "north" else 1.0
            en = np.stack([sgn * np.cos(l), sgn * np.sin(l), np.zeros_like(l)], axis=1)
            return ee, en
        idx = np.where(m)[0]
        ee, en = basis(lam[idx])
        v3 = ee * oc.uo[row, idx][:, None] + en * oc.vo[row, idx][:, None]
        mean3 = np.mean(v3, axis=0)
        ea, na = basis(lam)
        oc.uo[row, m] = (ea @ mean3)[m]
        oc.vo[row, m] = (na @ mean3)[m]


def ocean_step(oc, g, p, dt, u_atm, v_atm, Q_net=None, ice_mask=None):
    """WindDrivenSlabOcean.step (ocean.py:265-533).  Mutates and returns ``oc``."""
    a, dlat, dlon = g.a, g.dlat, g.dlon
    oc.step += 1
    cosr = np.maximum(g.cos, 0.5)
    cosc = _col(cosr)
    f = _col(g.f)
    u_rel = u_atm - oc.uo
    v_rel = v_atm - oc.vo
    Va = np.sqrt(u_rel ** 2 + v_rel ** 2)
    Va_eff = np.minimum(Va, p.oc_vcap)
    tau_x = p.oc_tau_scale * (p.oc_rho_a * p.oc_CD * Va_eff * u_rel)
    tau_y = p.oc_tau_scale * (p.oc_rho_a * p.oc_CD * Va_eff * v_rel)
    n_sub = ocean_nsub(oc, g, p, dt, u_atm, v_atm)
    oc.n_sub_last = n_sub
    sub = dt / n_sub
    on_land = (oc.land_mask == 1)
    ocean = ~on_land
    lat_deg_abs = np.abs(np.rad2deg(g.lat_rad))
    s = np.clip((lat_deg_abs - p.oc_polar_lat0) / max(1e-6, 90.0 - p.oc_polar_lat0), 0.0, 1.0)
    r_extra = _col(p.oc_polar_gain * (s ** 2))
    w_o = _col(g.w) * ocean

    for _ in range(n_sub):
        de_dl = (np.roll(oc.eta, -1, axis=1) - np.roll(oc.eta, 1, axis=1)) / (2.0 * dlon)
        de_dp = (np.roll(oc.eta, -1, axis=0) - np.roll(oc.eta, 1, axis=0)) / (2.0 * dlat)
        gx = de_dl / (a * cosc)
        gy = de_dp / a
        du = (f * oc.vo - p.oc_g * gx + tau_x / (p.oc_rho_w * p.oc_H) - p.oc_r_bot * oc.uo)
        dv = (-f * oc.uo - p.oc_g * gy + tau_y / (p.oc_rho_w * p.oc_H) - p.oc_r_bot * oc.vo)
        oc.uo = oc.uo + sub * du
        oc.vo = oc.vo + sub * dv
        oc.uo[on_land] = 0.0
        oc.vo[on_land] = 0.0
        oc.uo = oc.uo - sub * r_extra * oc.uo
        oc.vo = oc.vo - sub * r_extra * oc.vo

        if (p.oc_diff_every > 0) and (oc.step % p.oc_diff_every == 0):
            dx_min = np.minimum(a * dlat, a * dlon * cosr)
            k4 = p.oc_sigma4 * (dx_min ** 4) / max(1e-12, sub)
            k4u = _col(k4) if p.oc_k4_u is None else float(p.oc_k4_u)
            k4v = _col(k4) if p.oc_k4_v is None else float(p.oc_k4_v)
            k4e = _col(0.5 * k4) if p.oc_k4_eta is None else float(p.oc_k4_eta)
            oc.uo = ops.hyperdiffuse(oc.uo, k4u, sub, p.oc_k4_nsub, dlat, dlon, cosr, a)
            oc.vo = ops.hyperdiffuse(oc.vo, k4v, sub, p.oc_k4_nsub, dlat, dlon, cosr, a)
            oc.eta = ops.hyperdiffuse(oc.eta, k4e, sub, p.oc_k4_nsub, dlat, dlon, cosr, a)
        if (p.oc_shapiro_n > 0) and (p.oc_shapiro_every > 0) and (oc.step % p.oc_shapiro_every == 0):
            oc.uo = ops.shapiro(oc.uo, p.oc_shapiro_n)
            oc.vo = ops.shapiro(oc.vo, p.oc_shapiro_n)
            oc.eta = ops.shapiro(oc.eta, p.oc_shapiro_n)

        div = ops.divergence(oc.uo, oc.vo, g.lat, dlat, dlon, a)
        oc.eta = oc.eta + (-sub * p.oc_H * div)
        oc.eta[on_land] = 0.0
        if np.any(ocean):
            oc.eta = oc.eta - float(np.sum(oc.eta * w_o) / (np.sum(w_o) + 1e-15))

        Ts_adv = ops.advect_semilag(oc.Ts, oc.uo, oc.vo, sub, a, dlat, dlon, cosr)
        oc.Ts = (1.0 - p.oc_adv_alpha) * oc.Ts + p.oc_adv_alpha * Ts_adv
        if p.oc_K_h > 0.0:
            oc.Ts = ops.nan_to_num(oc.Ts)       # ocean.py:112 cleans the caller's array (copy=False)
            oc.Ts = oc.Ts + sub * p.oc_K_h * ops.laplacian(oc.Ts, dlat, dlon, cosr, a)
        if p.oc_use_qnet and (Q_net is not None):
            tend = Q_net / (p.oc_rho_w * p.oc_cp_w * p.oc_H)
            if ice_mask is not None:
                open_m = ocean & (~ice_mask)
                ice_m = ocean & ice_mask
                T = np.where(open_m, oc.Ts + sub * tend, oc.Ts)
                if p.oc_ice_qfac > 0.0:
                    T = np.where(ice_m, T + sub * p.oc_ice_qfac * tend, T)
                oc.Ts = T
            else:
                oc.Ts = np.where(ocean, oc.Ts + sub * tend, oc.Ts)

        oc.uo = ops.nan_to_num(oc.uo)
        oc.vo = ops.nan_to_num(oc.vo)
        speed = np.sqrt(oc.uo ** 2 + oc.vo ** 2)
        cap = float(p.oc_max_u)
        if p.oc_outlier == "mean4":
            um = 0.25 * (np.roll(oc.uo, -1, 0) + np.roll(oc.uo, 1, 0) + np.roll(oc.uo, -1, 1) + np.roll(oc.uo, 1, 1))
            vm = 0.25 * (np.roll(oc.vo, -1, 0) + np.roll(oc.vo, 1, 0) + np.roll(oc.vo, -1, 1) + np.roll(oc.vo, 1, 1))
            fast = speed > cap
            oc.uo = np.where(fast, um, oc.uo)
            oc.vo = np.where(fast, vm, oc.vo)
            sp2 = np.sqrt(oc.uo ** 2 + oc.vo ** 2)
            sc2 = np.where(sp2 > cap, cap / (sp2 + 1e-12), 1.0)
            oc.uo = oc.uo * sc2
            oc.vo = oc.vo * sc2
        else:
            sc1 = np.where(speed > cap, cap / (speed + 1e-12), 1.0)
            oc.uo = oc.uo * sc1
            oc.vo = oc.vo * sc1
        oc.eta = np.clip(ops.nan_to_num(oc.eta), -p.oc_eta_cap, p.oc_eta_cap)
        oc.Ts = ops.nan_to_num(oc.Ts)

    if p.oc_polar_fix:
        polar_fill(oc, g)
    oc.Ts = np.clip(oc.Ts, p.oc_ts_min, p.oc_ts_max)
    return oc


# =========================================================================== forcing
def star_geometry(t):
    """Host scalars of forcing.py:78-136 / orbital.py:15-52 for time t:
    returns [(flux, sin_delta, cos_delta, alpha) for star A, B] and theta."""
    G, M_SUN, L_SUN, AU = 6.67430e-11, 1.989e30, 3.828e26, 1.496e11
    M_A, M_B = 0.914 * M_SUN, 0.8 * M_SUN
    L_A, L_B = 0.7 * L_SUN, 0.410 * L_SUN
    M_T = M_A + M_B
    A_BIN, A_PL = 0.5 * AU, 1.32 * AU
    T_bin = 2 * np.pi * np.sqrt(A_BIN ** 3 / (G * M_T))
    T_pl = 2 * np.pi * np.sqrt(A_PL ** 3 / (G * M_T))
    om_b, om_p = 2 * np.pi / T_bin, 2 * np.pi / T_pl
    r_A, r_B = A_BIN * (M_B / M_T), A_BIN * (M_A / M_T)
    tilt = np.deg2rad(27.0)
    n_hat = np.array([np.sin(tilt), 0.0, np.cos(tilt)])
    x_in = np.array([1.0, 0.0, 0.0])
    x_eq = x_in - np.dot(x_in, n_hat) * n_hat
    x_eq /= np.linalg.norm(x_eq)
    y_eq = np.cross(n_hat, x_eq)
    ang = om_p * t
    x_A, y_A = r_A * np.cos(om_b * t), r_A * np.sin(om_b * t)
    x_B, y_B = -r_B * np.cos(om_b * t), -r_B * np.sin(om_b * t)
    x_p, y_p = A_PL * np.cos(ang), A_PL * np.sin(ang)
    out = []
    for (xs, ys, L) in ((x_A, y_A, L_A), (x_B, y_B, L_B)):
        vec = np.array([xs - x_p, ys - y_p, 0.0])
        dist = np.linalg.norm(vec)
        flux = L / (4 * np.pi * (dist ** 2))
        s_hat = vec / (np.linalg.norm(vec) + 1e-15)
        delta = np.arcsin(np.clip(np.dot(s_hat, n_hat), -1.0, 1.0))
        alpha = np.arctan2(np.dot(s_hat, y_eq), np.dot(s_hat, x_eq))
        out.append((float(flux), float(np.sin(delta)), float(np.cos(delta)), float(alpha)))
    theta = (t * PLANET_OMEGA) % (2 * np.pi)
    return out, float(theta)


def insolation(g, t):
    """forcing.py:105-136 for both stars -> (isr_A, isr_B)."""
    stars, theta = star_geometry(t)
    res = []
    for flux, sd, cd, alpha in stars:
        hang = theta + g.lon_rad[None, :] - alpha
        cz = _col(g.sin) * sd + _col(g.cos) * cd * np.cos(hang)
        res.append(flux * np.maximum(0.0, cz))
    return res[0], res[1]


def teq_field(isr_total, albedo):
    """forcing.py:138-165 (insolation is recomputed there; same value)."""
    num = isr_total * (1 - albedo)
    num = np.where(num < 0, 0.0, num)
    return (num / SIGMA) ** 0.25


# =========================================================================== script loop physics
def orographic_factor(g, elevation, u, v, p):
    """physics.py:116-161."""
    cosl = _col(np.maximum(g.cos, 1e-6))
    dx = g.a * cosl * g.dlon
    dy = g.a * g.dlat
    dHdx = (np.roll(elevation, -1, 1) - np.roll(elevation, 1, 1)) / (2.0 * dx)
    dHdy = (np.roll(elevation, -1, 0) - np.roll(elevation, 1, 0)) / (2.0 * dy)
    dHdy[0] = 0.0
    dHdy[-1] = 0.0
    gn = np.sqrt(dHdx ** 2 + dHdy ** 2)
    nx = np.where(gn > 1e-12, dHdx / (gn + 1e-12), 0.0)
    ny = np.where(gn > 1e-12, dHdy / (gn + 1e-12), 0.0)
    fac = np.clip(1.0 + p.k_orog * np.maximum(0.0, u * nx + v * ny), 1.0, 2.0)
    return ops.gaussian(fac, 1.0)


def precip_hybrid(st, g, p, orog=None):
    """physics.py:253-354 (+ legacy fallback physics.py:12-46 with cloud gating off)."""
    Pq = np.maximum(0.0, st.P_cond)
    div = ops.divergence(st.u, st.v, g.lat, g.dlat, g.dlon, g.a)
    pos = np.maximum(0.0, -(div - p.D_crit))
    if np.any(pos > 0):
        scale = max(ops.median_pos(pos), 1e-12)
        F_div = np.clip(pos / scale, 0.0, 5.0)
    else:
        F_div = np.zeros_like(Pq)
    F_or = 1.0 if orog is None else np.clip(orog, 1.0, 3.0)
    F = (1.0 + p.beta_div * F_div) * F_or
    P_raw = Pq * F
    w = _col(g.w)
    num = float(np.sum(Pq * w))
    den = float(np.sum(P_raw * w)) + 1e-20
    s = num / den if den > 0 else 1.0
    P = ops.gaussian(P_raw * s, 1.0)
    if p.p_fallback:
        wsum = float(np.sum(np.broadcast_to(w, Pq.shape)) + 1e-15)
        if float(np.sum(Pq * w) / wsum) < p.pq_min:
            P_dyn = ops.gaussian(p.k_precip * np.maximum(0.0, -(div - p.D_crit)), 1.0)
            P = (1.0 - p.p_blend) * P + p.p_blend * P_dyn
    return np.clip(P, 0.0, None)


def cloud_source(st, g):
    """physics.py:72-114."""
    src = 0.5 * np.clip(np.tanh((st.T_s - 285.0) / 12.0), 0.0, 1.0)
    vort = ops.vorticity(st.u, st.v, g.lat, g.dlat, g.dlon, g.a)
    rel = vort / (_col(g.f) + 1e-12)
    src = src + 0.4 * np.clip(np.tanh((rel - 0.5) / 2.0), 0.0, 1.0)
    dx = g.dlon * g.a * _col(np.maximum(1e-6, g.cos))
    dy = g.dlat * g.a
    gx = (np.roll(st.T_s, -1, 1) - np.roll(st.T_s, 1, 1)) / (2 * dx)
    gy = (np.roll(st.T_s, -1, 0) - np.roll(st.T_s, 1, 0)) / (2 * dy)
    adv = -(st.u * gx + st.v * gy)
    src = src + 0.3 * np.clip(np.tanh(np.abs(adv) / 2e-5), 0.0, 1.0)
    return np.clip(ops.gaussian(src, 1.0), 0.0, 1.0)


def dynamic_albedo(cloud, base, ice_frac, land_mask, p):
    """physics.py:164-250 as called at run_simulation.py:2144 (ice_frac given, ice only over ocean)."""
    C = np.clip(cloud, 0.0, 1.0)
    fi = np.clip(ice_frac, 0.0, 1.0) * (land_mask == 0)
    surf = base * (1.0 - fi) + float(p.alpha_ice) * fi
    return np.clip(surf * (1.0 - C) + float(p.alpha_cloud) * C, 0.0, 1.0)


def surface_qnet(st, g, p, albedo):
    """run_simulation.py:2201-2239: post-dynamics SW/LW/SH/LH -> Q_net and ice mask."""
    ice_mask = st.h_ice > 0.0
    cloud_eff = st.cloud_eff if getattr(st, "cloud_eff", None) is not None else st.cloud
    _, SW_sfc, _ = shortwave(st.isr, albedo, cloud_eff, p)
    T_a = 288.0 + (9.81 / 1004.0) * st.h
    ice_frac = 1.0 - np.exp(-np.maximum(st.h_ice, 0.0) / max(1e-6, p.hice_ref))
    if p.lw_v2:
        _, LW_sfc, _, _, _ = longwave_v2(st.T_s, T_a, cloud_eff, emissivity_map(st.land_mask, ice_frac, p), p)
    else:
        _, LW_sfc, _, _, _ = longwave_v1(st.T_s, T_a, cloud_eff, p)
    SH = sensible_heat(st.T_s, T_a, st.u, st.v, p)
    return SW_sfc - LW_sfc - SH - st.LH, ice_mask


def snow_phase_step(st, g, p, precip, dt):
    """run_simulation.py:1948-2008 + hydrology.py:100-177 (P019 lapse, sigmoid split, snowpack, glacier)."""
    land = (st.land_mask == 1)
    T_a = 288.0 + (9.81 / 1004.0) * st.h
    Hb = st.elevation if st.elevation is not None else np.zeros_like(st.T_s)
    h_snow = np.where(land, np.maximum(st.S_snow, 0.0) / max(p.rho_snow, 1e-6), 0.0)
    polar = np.abs(_col(g.lat)) >= p.polar_lat_thresh
    h_eff = np.where(polar, np.minimum(h_snow, p.polar_ice_thick_max), h_snow)
    H_eff = np.minimum(Hb + h_eff, p.land_elev_max)
    T_hat = T_a - p.lapse_kpm * (H_eff / 1000.0) if p.lapse_enable else T_a
    f_snow = np.clip(1.0 / (1.0 + np.exp((T_hat - p.snow_thresh) / max(1e-6, p.snow_t_band))), 0.0, 1.0)
    P_snow = ops.nan_to_num(f_snow * precip)
    P_rain = ops.nan_to_num((1.0 - f_snow) * precip)
    if p.swe_enable:
        Ps_land = P_snow * land
        if p.snow_melt_mode == "degree_day":
            melt = (p.snow_ddf / 86400.0) * np.maximum(T_hat - p.snow_melt_tref, 0.0)
        else:
            melt = np.where(T_hat >= p.snow_thresh, p.snow_melt_rate / 86400.0, 0.0)
        amt = np.minimum(np.maximum(st.S_snow, 0.0), melt * dt)
        S_next = st.S_snow + Ps_land * dt - amt
        if p.swe_max is not None and p.swe_max > 0:
            S_next = np.minimum(S_next, p.swe_max)
        S_next = np.maximum(0.0, S_next)
        melt_out = ops.nan_to_num(amt / dt)
        C_snow = np.clip(1.0 - np.exp(-np.maximum(S_next, 0.0) / max(1e-6, p.swe_ref)), 0.0, 1.0)
        S_next = ops.nan_to_num(S_next)
        glacier = land & ((C_snow >= p.glacier_frac) | (S_next >= p.glacier_swe))
        rain_gl = (P_rain * land) * glacier
        if np.any(rain_gl):
            S_next = S_next + rain_gl * dt
    else:
        C_snow = np.zeros_like(st.T_s)
        glacier = land & (C_snow >= p.glacier_frac)
        S_next = st.S_snow.copy()
        melt_out = np.zeros_like(st.T_s)
    return SimpleNamespace(P_rain=P_rain, P_snow=P_snow, S_next=S_next, melt=melt_out,
                           C_snow=C_snow, glacier=glacier, T_hat=T_hat)


def land_bucket(W, P_in, E_land, p, dt):
    """hydrology.py:219-260."""
    tau = max(1.0, float(p.runoff_tau_days) * 86400.0)
    R_base = W / tau
    W_next = np.maximum(0.0, W + (P_in - E_land - R_base) * dt)
    if p.wland_cap is not None and p.wland_cap > 0:
        over = np.maximum(0.0, W_next - float(p.wland_cap))
        W_next = W_next - over
        R_fast = over / dt
    else:
        R_fast = 0.0
    return ops.nan_to_num(W_next), ops.nan_to_num(R_base + R_fast)


def loop_step(st, oc, g, p, t, dt, eco_alpha=None, with_albedo_arg=False):
    """One iteration of the script loop (run_simulation.py:1760-2344), without plotting / daily
    ecology / autosave.  ``eco_alpha`` is the land alpha map returned by the ecology sub-daily
    step for this step (None = ecology off).  Returns a namespace of per-step diagnostics."""
    out = SimpleNamespace()
    land = (st.land_mask == 1)
    orog = None
    if p.orog_enabled and st.elevation is not None:
        orog = orographic_factor(g, st.elevation, st.u, st.v, p)
    precip = precip_hybrid(st, g, p, orog)
    out.precip = precip

    # cloud from precip + source + blend (run_simulation.py:1866-1913)
    if np.any(precip > 0):
        P_ref = float(p.pref) if p.pref is not None else ops.median_pos(precip, empty=1e-6)
    else:
        P_ref = 1e-6
    C_P = np.clip(ops.gaussian(p.cmax * np.tanh(precip / (P_ref + 1e-12)), 1.0), 0.0, 1.0)
    src = cloud_source(st, g)
    tend = src * (dt / (6 * 3600))
    wm, wp, ws = p.w_mem, p.w_p, p.w_src
    wsum = wm + wp + ws
    if wsum <= 0:
        wm, wp, ws, wsum = 0.5, 0.4, 0.1, 1.0
    wm /= wsum
    wp /= wsum
    ws /= wsum
    cl = wm * st.cloud + wp * C_P + ws * np.clip(st.cloud + tend, 0.0, 1.0)
    if p.cloud_floor > 0.0:
        cl = np.maximum(cl, np.clip(p.cloud_floor * C_P, 0.0, 1.0))
    cl = np.clip(cl, 0.0, 1.0)
    if p.cloud_advect:
        adv = ops.advect_semilag(cl, st.u, st.v, dt, g.a, g.dlat, g.dlon, np.maximum(g.cos, 0.5))
        if p.cloud_smooth_sigma > 0.0:
            adv = ops.gaussian(adv, p.cloud_smooth_sigma, mode="wrap")
        cl = np.clip((1.0 - p.cloud_adv_alpha) * cl + p.cloud_adv_alpha * adv, 0.0, 1.0)
    st.cloud = cl

    # insolation (run_simulation.py:1942-1944)
    st.isr_A, st.isr_B = insolation(g, t)
    st.isr = st.isr_A + st.isr_B

    # P019 snow / phase (provisional)
    sn = snow_phase_step(st, g, p, precip, dt)
    out.C_snow, out.glacier = sn.C_snow, sn.glacier

    # albedo synthesis (run_simulation.py:2064-2146)
    ice_frac = 1.0 - np.exp(-np.maximum(st.h_ice, 0.0) / max(1e-6, p.hice_ref))
    cloud_rad = st.cloud_eff if getattr(st, "cloud_eff", None) is not None else st.cloud
    base = st.base_albedo.copy() if p.use_topo_albedo else np.full_like(st.T_s, float(p.alpha_water))
    if eco_alpha is not None:
        m = land & (~sn.glacier) & np.isfinite(eco_alpha)
        base[m] = ((1.0 - p.eco_lai_albedo_weight) * base + p.eco_lai_albedo_weight * eco_alpha)[m]
    if p.swe_enable:
        base[land] = np.clip((1.0 - sn.C_snow) * base + sn.C_snow * p.snow_albedo_fresh, 0.0, 1.0)[land]
    albedo = dynamic_albedo(cloud_rad, base, ice_frac, st.land_mask, p)
    out.albedo = albedo
    Teq = teq_field(st.isr, albedo)
    out.Teq = Teq

    atmos_step(st, g, p, Teq, dt, albedo=albedo if with_albedo_arg else None)

    if oc is not None:
        Q_net, ice_mask = surface_qnet(st, g, p, albedo)
        out.Q_net = Q_net
        ocean_step(oc, g, p, dt, st.u, st.v, Q_net=Q_net, ice_mask=ice_mask)
        st.T_s = np.where((st.land_mask == 0) & (~ice_mask), oc.Ts, st.T_s)

    # hydrology commit (run_simulation.py:2294-2339)
    E = st.E_flux
    st.S_snow = sn.S_next
    non_gl = land & (~sn.glacier)
    P_in = (sn.P_rain * land + sn.melt) * non_gl
    E_land = (E * land) * non_gl
    st.W_land, R_bucket = land_bucket(st.W_land, P_in, E_land, p, dt)
    out.R_land = R_bucket + sn.melt * sn.glacier
    return out


# =========================================================================== routing
def routing_event(acc, flow_order, flow_to, land_flat, lake_is, lake_id, lake_outlet):
    """Sequential topological push of routing.py:261-298.  ``acc`` (kg per cell) is modified
    in place; returns (flow_accum_kg, ocean_inflow_kg, lake_store_kg)."""
    n = acc.size
    flow_acc = np.zeros(n)
    ocean_kg = 0.0
    n_lakes = 0 if lake_outlet is None else len(lake_outlet)
    lake_store = np.zeros(max(n_lakes, int(lake_id.max()) if lake_id is not None else 0))
    has_lakes = lake_is is not None and lake_id is not None and n_lakes > 0
    for idx in flow_order:
        m = acc[idx]
        if m <= 0.0:
            continue
        flow_acc[idx] += m
        if has_lakes and lake_is[idx]:
            lid = int(lake_id[idx])
            if lid > 0 and lid <= n_lakes:
                o = int(lake_outlet[lid - 1])
                if o < 0:
                    ocean_kg += m
                elif o < n and land_flat[o]:
                    acc[o] += m
                else:
                    ocean_kg += m
            elif lid > 0:
                lake_store[lid - 1] += m
            acc[idx] = 0.0
            continue
        dn = int(flow_to[idx])
        if dn < 0 or not land_flat[dn]:
            ocean_kg += m
        else:
            acc[dn] += m
        acc[idx] = 0.0
    return flow_acc, ocean_kg, lake_store


def cell_area_rows(g):
    """routing.py:176-200."""
    dphi = np.deg2rad(abs(g.lat[1] - g.lat[0]))
    dlam = np.deg2rad(abs(g.lon[1] - g.lon[0]))
    pc = np.deg2rad(g.lat)
    band = np.sin(np.clip(pc + 0.5 * dphi, -0.5 * np.pi, 0.5 * np.pi)) - np.sin(np.clip(pc - 0.5 * dphi, -0.5 * np.pi, 0.5 * np.pi))
    return (g.a * g.a) * dlam * band


# =========================================================================== state helpers
def new_atmos_state(g, p, land_mask, friction, base_albedo=None, elevation=None, C_s_map=None):
    """SpectralModel.__init__ initial fields (dynamics.py:56-88) + land reservoirs."""
    st = SimpleNamespace()
    shp = (g.nlat, g.nlon)
    st.u = np.zeros(shp)
    st.v = np.zeros(shp)
    st.h = np.full(shp, float(p.H)) + 300 * (np.sin(_col(g.lat_rad)) ** 2) * np.ones(shp)
    st.T_s = np.full(shp, 288.0)
    st.cloud = np.zeros(shp)
    st.h_ice = np.zeros(shp)
    st.q = float(np.clip(p.q_init_rh, 0.0, 1.0)) * q_sat(st.T_s, p.p0)
    st.isr = np.zeros(shp)
    st.isr_A = np.zeros(shp)
    st.isr_B = np.zeros(shp)
    st.olr = np.zeros(shp)
    st.E_flux = np.zeros(shp)
    st.P_cond = np.zeros(shp)
    st.LH = np.zeros(shp)
    st.LH_release = np.zeros(shp)
    st.cloud_eff = None
    st.step_counter = 0
    st.land_mask = np.asarray(land_mask).astype(np.uint8)
    st.friction = np.asarray(friction, dtype=np.float64)
    st.base_albedo = None if base_albedo is None else np.asarray(base_albedo, dtype=np.float64)
    st.elevation = None if elevation is None else np.asarray(elevation, dtype=np.float64)
    st.C_s_map = None if C_s_map is None else np.asarray(C_s_map, dtype=np.float64)
    st.W_land = np.zeros(shp)
    st.S_snow = np.zeros(shp)
    return st


def new_ocean_state(g, land_mask, init_Ts=None):
    """WindDrivenSlabOcean.__init__ fields (ocean.py:85-97)."""
    oc = SimpleNamespace()
    shp = (g.nlat, g.nlon)
    oc.uo = np.zeros(shp)
    oc.vo = np.zeros(shp)
    oc.eta = np.zeros(shp)
    oc.Ts = np.full(shp, 288.0) if init_Ts is None else np.array(init_Ts, dtype=np.float64, copy=True)
    oc.land_mask = np.asarray(land_mask).astype(int)
    oc.step = 0
    return oc
