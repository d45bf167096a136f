#!/usr/bin/env python3
"""Generate golden vectors by importing and running the REFERENCE (mountain/qingdai).

Runs only where the reference checkout is mounted (``/root/reference``); the produced
``tests/golden/*.npz`` fixtures are committed so that every other machine (the GPU box
included) can check the oracle and the CUDA path against the reference's own outputs.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

Nothing here is copied from the reference: it is imported, called, and its outputs saved.
matplotlib / netCDF4 are absent in this image, so inert stand-ins are placed in
``sys.modules`` before the reference modules that hard-import them are loaded.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types
from unittest import mock

import numpy as np

REF = os.environ.get("QD_REFERENCE_ROOT", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.dirname(os.path.abspath(__file__))

QUIET_ENV = {
    "QD_ENERGY_DIAG": "0", "QD_HUMIDITY_DIAG": "0", "QD_WATER_DIAG": "0", "QD_OCEAN_DIAG": "0",
    "QD_OCEAN_ENERGY_DIAG": "0", "QD_HYDRO_DIAG": "0", "QD_ECO_DIAG": "0", "QD_PHYTO_DIAG": "0",
    "QD_USE_JAX": "0", "QD_AUTOSAVE_ENABLE": "0", "QD_AUTOSAVE_LOAD": "0",
    "QD_PHYTO_ENABLE": "0", "QD_ECO_INDIV_ENABLE": "0", "MPLBACKEND": "Agg",
}


def _install_stubs():
    if "matplotlib" not in sys.modules:
        m = mock.MagicMock()
        m.pyplot.subplots.side_effect = lambda *a, **k: (mock.MagicMock(), mock.MagicMock())
        sys.modules["matplotlib"] = m
        sys.modules["matplotlib.pyplot"] = m.pyplot
        for sub in ("colors", "cm", "gridspec", "ticker", "patches"):
            sys.modules[f"matplotlib.{sub}"] = getattr(m, sub)
    sys.path.insert(0, REF)


class MemDataset:
    """Tiny in-memory stand-in for netCDF4.Dataset (read side only) used to feed
    RiverRouting.__init__ (routing.py:105-154) a network built in this process."""
    store: dict = {}

    def __init__(self, path, mode="r"):
        self.variables = {k: np.asarray(v) for k, v in MemDataset.store[path].items()}

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def set_env(extra=None):
    for k in list(os.environ):
        if k.startswith("QD_"):
            del os.environ[k]
    os.environ.update(QUIET_ENV)
    if extra:
        os.environ.update({k: str(v) for k, v in extra.items()})


# ------------------------------------------------------------------------------------ operators
def gen_ops():
    from pygcm.grid import SphericalGrid
    from pygcm.dynamics import SpectralModel
    from pygcm.ocean import WindDrivenSlabOcean
    from pygcm import physics, energy, humidity, hydrology
    import scripts.run_simulation as rs

    out = {}
    for tag, (nlat, nlon) in {"a": (22, 40), "b": (15, 27)}.items():
        set_env()
        rng = np.random.default_rng(100 + nlat)
        grid = SphericalGrid(nlat, nlon)
        land = (rng.uniform(size=(nlat, nlon)) < 0.3).astype(np.uint8)
        with quiet():
            gcm = SpectralModel(grid, np.full((nlat, nlon), 1e-5), land_mask=land)
            oc = WindDrivenSlabOcean(grid, land, 50.0)
        F = rng.standard_normal((nlat, nlon)) * 10.0 + 280.0
        u = rng.standard_normal((nlat, nlon)) * 60.0
        v = rng.standard_normal((nlat, nlon)) * 40.0
        # a few violent winds so departure points wrap over poles / seam several times
        u[3, 5], v[3, 5], u[0, 0], v[-1, -1] = 5.0e4, -3.0e4, -2.0e5, 9.0e4
        dt = 1800.0
        gcm.u, gcm.v = u.copy(), v.copy()
        out[f"{tag}_F"], out[f"{tag}_u"], out[f"{tag}_v"] = F, u, v
        out[f"{tag}_land"] = land
        out[f"{tag}_dt"] = dt
        out[f"{tag}_adv_atm"] = gcm._advect(F.copy(), dt)
        out[f"{tag}_adv_oc"] = oc._advect_scalar(F.copy(), u * 0.01, v * 0.01, dt)
        out[f"{tag}_adv_cloud"] = rs._advect_scalar_periodic(F.copy(), u, v, dt, grid)
        out[f"{tag}_lap_atm"] = gcm._laplacian_sphere(F.copy())
        out[f"{tag}_lap_oc"] = oc._laplacian_sphere(F.copy())
        k4row = 0.02 * np.minimum(6.371e6 * grid.dlat_rad, 6.371e6 * grid.dlon_rad * np.maximum(np.cos(np.deg2rad(grid.lat_mesh)), 1e-3)) ** 4 / dt
        out[f"{tag}_k4map"] = k4row
        out[f"{tag}_hyp_atm_map"] = gcm._hyperdiffuse(F.copy(), k4row, dt, n_substeps=1)
        out[f"{tag}_hyp_atm_map3"] = gcm._hyperdiffuse(F.copy(), 0.5 * k4row, dt, n_substeps=3)
        out[f"{tag}_hyp_atm_scalar"] = gcm._hyperdiffuse(F.copy(), 1.0e14, dt, n_substeps=2)
        out[f"{tag}_hyp_oc_map"] = oc._hyperdiffuse(F.copy(), dt, k4row, n_substeps=1)
        Fn = F.copy()
        Fn[2, 3], Fn[5, 7], Fn[0, 1] = np.nan, np.inf, -np.inf
        out[f"{tag}_Fnan"] = Fn
        out[f"{tag}_lap_nan"] = gcm._laplacian_sphere(Fn.copy())
        out[f"{tag}_shapiro2"] = gcm._shapiro_filter(F.copy(), n=2)
        out[f"{tag}_shapiro1"] = gcm._shapiro_filter(F.copy(), n=1)
        out[f"{tag}_spec"] = gcm._spectral_zonal_filter(F.copy(), cutoff=0.75, damp=0.5)
        out[f"{tag}_spec2"] = gcm._spectral_zonal_filter(F.copy(), cutoff=0.3, damp=1.0)
        out[f"{tag}_div"] = grid.divergence(u, v)
        out[f"{tag}_vort"] = grid.vorticity(u, v)
        from scipy.ndimage import gaussian_filter
        out[f"{tag}_gauss1"] = gaussian_filter(F, sigma=1.0)
        out[f"{tag}_gauss02w"] = gaussian_filter(F, sigma=0.2, mode="wrap")
        pos = np.maximum(0.0, F - 280.0)
        out[f"{tag}_median_pos"] = np.array(float(np.median(pos[pos > 0])))
        # cell physics
        hp = humidity.get_humidity_params_from_env()
        ep = energy.get_energy_params_from_env()
        Ts = 240.0 + 70.0 * rng.uniform(size=(nlat, nlon))
        Ta = 230.0 + 60.0 * rng.uniform(size=(nlat, nlon))
        q = 0.02 * rng.uniform(size=(nlat, nlon))
        cloud = rng.uniform(size=(nlat, nlon))
        hice = np.where(rng.uniform(size=(nlat, nlon)) < 0.3, rng.uniform(size=(nlat, nlon)) * 2.0, 0.0)
        isr = 900.0 * rng.uniform(size=(nlat, nlon))
        alb = rng.uniform(size=(nlat, nlon))
        for k, val in dict(Ts=Ts, Ta=Ta, q=q, cloud=cloud, hice=hice, isr=isr, alb=alb).items():
            out[f"{tag}_{k}"] = val
        out[f"{tag}_qsat"] = humidity.q_sat(Ts)
        fac = humidity.surface_evaporation_factor(land, hice, hp)
        out[f"{tag}_evapfac"] = fac
        out[f"{tag}_E"] = humidity.evaporation_flux(Ts, q, u, v, fac, hp)
        Pc, qn = humidity.condensation(q * 3.0, Ta, dt, hp)
        out[f"{tag}_Pcond"], out[f"{tag}_qnext"] = Pc, qn
        swa, sws, R = energy.shortwave_radiation(isr, alb, cloud, ep)
        out[f"{tag}_sw_atm"], out[f"{tag}_sw_sfc"], out[f"{tag}_sw_R"] = swa, sws, R
        ice_frac = 1.0 - np.exp(-np.maximum(hice, 0.0) / 0.5)
        eps = energy.surface_emissivity_map(land, ice_frac)
        out[f"{tag}_eps_sfc"] = eps
        lwa, lws, olr, dlr, ee = energy.longwave_radiation_v2(Ts, Ta, cloud, eps, ep)
        out[f"{tag}_lw2_atm"], out[f"{tag}_lw2_sfc"], out[f"{tag}_lw2_olr"] = lwa, lws, olr
        lwa, lws, olr, dlr, ee = energy.longwave_radiation(Ts, Ta, cloud, ep)
        out[f"{tag}_lw1_atm"], out[f"{tag}_lw1_sfc"], out[f"{tag}_lw1_olr"] = lwa, lws, olr
        SH, _ = energy.boundary_layer_fluxes(Ts, Ta, u, v, land)
        out[f"{tag}_SH"] = SH
        LH = 2.5e6 * out[f"{tag}_E"]
        Tn, hn = energy.integrate_surface_energy_with_seaice(
            Ts, sws, lws, SH, LH, dt, land, hice, 2.1e8, 3e6, 5e6)
        out[f"{tag}_seaice_Ts"], out[f"{tag}_seaice_h"] = Tn, hn
        out[f"{tag}_alb_dyn"] = physics.calculate_dynamic_albedo(cloud, Ts, alb * 0.5, 0.6, 0.5, land_mask=land, ice_frac=ice_frac)
        # hydrology
        hyp = hydrology.get_hydrology_params_from_env()
        P = 1e-4 * rng.uniform(size=(nlat, nlon))
        S = 60.0 * rng.uniform(size=(nlat, nlon)) * land
        W = 30.0 * rng.uniform(size=(nlat, nlon)) * land
        out[f"{tag}_P"], out[f"{tag}_S"], out[f"{tag}_W"] = P, S, W
        pr, ps, fs = hydrology.partition_precip_phase_smooth(P, Ta + 35.0, hyp.snow_thresh_K, hyp.snow_t_band_K)
        out[f"{tag}_Prain"], out[f"{tag}_Psnow"] = pr, ps
        Sn, melt, Cs, _ = hydrology.snowpack_step(S, ps * land, Ta + 35.0, hyp, dt)
        out[f"{tag}_Snext"], out[f"{tag}_melt"], out[f"{tag}_Csnow"] = Sn, melt, Cs
        Wn, Rf = hydrology.update_land_bucket(W, pr * land, out[f"{tag}_E"] * land, hyp, dt)
        out[f"{tag}_Wnext"], out[f"{tag}_Rflux"] = Wn, Rf
        # composite physics
        gcm.T_s = Ts.copy()
        gcm.P_cond_flux_last = Pc.copy()
        gcm.cloud_cover = cloud.copy()
        out[f"{tag}_cloud_src"] = physics.parameterize_cloud_cover(gcm, grid, land)
        elev = rng.standard_normal((nlat, nlon)) * 1500.0
        out[f"{tag}_elev"] = elev
        orog = physics.compute_orographic_factor(grid, elev, u, v, k_orog=7e-4)
        out[f"{tag}_orog"] = orog
        out[f"{tag}_precip_hyb"] = physics.diagnose_precipitation_hybrid(gcm, grid, D_crit=-1e-7, k_precip=1e5, orog_factor=None, smooth_sigma=1.0, beta_div=0.4, renorm=True)
        out[f"{tag}_precip_hyb_orog"] = physics.diagnose_precipitation_hybrid(gcm, grid, D_crit=-1e-7, k_precip=1e5, orog_factor=orog, smooth_sigma=1.0, beta_div=0.4, renorm=True)
        gcm.P_cond_flux_last = Pc * 1e-9          # -> triggers the weak-humidity fallback blend
        out[f"{tag}_precip_hyb_fb"] = physics.diagnose_precipitation_hybrid(gcm, grid, D_crit=-1e-7, k_precip=1e5, orog_factor=None, smooth_sigma=1.0, beta_div=0.4, renorm=True)
        out[f"{tag}_cloud_from_p"] = physics.cloud_from_precip(out[f"{tag}_precip_hyb"], C_max=0.95, P_ref=float(np.median(out[f"{tag}_precip_hyb"][out[f"{tag}_precip_hyb"] > 0])), smooth_sigma=1.0)
    np.savez_compressed(os.path.join(OUT, "ops_golden.npz"), **out)
    print("ops_golden.npz:", len(out), "arrays")


# ------------------------------------------------------------------------------------ cores
ATM_FIELDS = ("u", "v", "h", "T_s", "q", "cloud_cover", "h_ice")
OC_FIELDS = ("uo", "vo", "eta", "Ts")


def snap_gcm(gcm):
    d = {k: np.array(getattr(gcm, k), dtype=np.float64, copy=True) for k in ATM_FIELDS}
    for k in ("olr", "E_flux_last", "P_cond_flux_last", "LH_last", "LH_release_last", "isr"):
        d[k] = np.array(getattr(gcm, k), dtype=np.float64, copy=True)
    ce = getattr(gcm, "cloud_eff_last", None)
    if ce is not None:
        d["cloud_eff_last"] = np.array(ce, dtype=np.float64, copy=True)
    return d


def snap_oc(oc):
    return {k: np.array(getattr(oc, k), dtype=np.float64, copy=True) for k in OC_FIELDS}


def build_world(nlat, nlon, banded=True):
    from pygcm.grid import SphericalGrid
    from pygcm.orbital import OrbitalSystem
    from pygcm.forcing import ThermalForcing
    from pygcm.dynamics import SpectralModel
    from pygcm.ocean import WindDrivenSlabOcean
    from pygcm.topography import create_land_sea_mask, generate_base_properties
    grid = SphericalGrid(nlat, nlon)
    with quiet():
        land = create_land_sea_mask(grid)
        base_alb, fric = generate_base_properties(land)
        Cs_map = np.where(land == 1, 3e6, 2.1e8).astype(float)
        forcing = ThermalForcing(grid, OrbitalSystem())
        gcm = SpectralModel(grid, fric, H=8000, tau_rad=10 * 24 * 3600,
                            greenhouse_factor=float(os.getenv("QD_GH_FACTOR", "0.40")),
                            C_s_map=Cs_map, land_mask=land, Cs_ocean=2.1e8, Cs_land=3e6, Cs_ice=5e6)
        if banded:
            phi = np.deg2rad(grid.lat_mesh)
            gcm.T_s = 258.0 + (297.0 - 258.0) * np.cos(phi) ** 2
        oc = WindDrivenSlabOcean(grid, land, 50.0, init_Ts=np.where(land == 0, gcm.T_s, 288.0))
    return grid, land, base_alb, fric, Cs_map, forcing, gcm, oc


def gen_cores():
    """Teacher-forced single steps of SpectralModel.time_step (both call shapes) and
    WindDrivenSlabOcean.step, driven the way scripts/benchmark_jax.py:122-156 drives them."""
    from pygcm import energy
    out = {}
    nlat, nlon, dt = 31, 60, 600.0
    for tag, env in {"w1": {"QD_ENERGY_W": "1"}, "w0": {}, "w05k": {"QD_ENERGY_W": "0.5", "QD_K4_NSUB": "2", "QD_SPEC_EVERY": "3", "QD_MOM_SCHEME": "primitive"}}.items():
        set_env(env)
        grid, land, base_alb, fric, Cs_map, forcing, gcm, oc = build_world(nlat, nlon)
        eparams = energy.get_energy_params_from_env()
        out[f"{tag}_land"], out[f"{tag}_base_alb"], out[f"{tag}_fric"] = land, base_alb, fric
        nsteps = 14
        for i in range(nsteps):
            t = i * dt + 3.0e5
            with quiet():
                insA, insB = forcing.calculate_insolation_components(t)
            gcm.isr_A, gcm.isr_B = insA, insB
            gcm.isr = insA + insB
            albedo = np.where(land == 0, 0.08, base_alb)
            Teq = forcing.calculate_equilibrium_temp(t, albedo)
            pre = snap_gcm(gcm)
            use_alb = (tag != "w0") or (i % 2 == 1)
            with quiet():
                gcm.time_step(Teq, dt, albedo=albedo if use_alb else None)
            post = snap_gcm(gcm)
            # ocean driven like benchmark_jax.py:134-156
            cloud_eff = getattr(gcm, "cloud_eff_last", gcm.cloud_cover)
            _, SW_sfc, _ = energy.shortwave_radiation(gcm.isr, albedo, cloud_eff, eparams)
            T_a = 288.0 + (9.81 / 1004.0) * gcm.h
            ice_frac = 1.0 - np.exp(-np.maximum(gcm.h_ice, 0.0) / 0.5)
            eps = energy.surface_emissivity_map(land, ice_frac)
            _, LW_sfc, _, _, _ = energy.longwave_radiation_v2(gcm.T_s, T_a, cloud_eff, eps, eparams)
            SH, _ = energy.boundary_layer_fluxes(gcm.T_s, T_a, gcm.u, gcm.v, land)
            Q_net = SW_sfc - LW_sfc - SH - gcm.LH_last
            ice_mask = gcm.h_ice > 0.0
            opre = snap_oc(oc)
            with quiet():
                oc.step(dt, gcm.u, gcm.v, Q_net=Q_net, ice_mask=ice_mask)
            opost = snap_oc(oc)
            gcm.T_s = np.where((land == 0) & (~ice_mask), oc.Ts, gcm.T_s)
            keep = {"w1": (0, 5, 12, 13), "w0": (0, 5), "w05k": (1, 2)}[tag]   # i=5 -> counter 6 (Shapiro); i=2 -> counter 3 (band-stop)
            if i in keep:
                for k, v in pre.items():
                    out[f"{tag}_s{i}_pre_{k}"] = v
                for k, v in post.items():
                    out[f"{tag}_s{i}_post_{k}"] = v
                out[f"{tag}_s{i}_Teq"], out[f"{tag}_s{i}_albedo"] = Teq, albedo
                out[f"{tag}_s{i}_use_alb"] = np.array(int(use_alb))
                out[f"{tag}_s{i}_counter_pre"] = np.array(i)
                out[f"{tag}_s{i}_t"] = np.array(t)
                out[f"{tag}_s{i}_Qnet"], out[f"{tag}_s{i}_ice_mask"] = Q_net, ice_mask
                for k, v in opre.items():
                    out[f"{tag}_s{i}_opre_{k}"] = v
                for k, v in opost.items():
                    out[f"{tag}_s{i}_opost_{k}"] = v
                out[f"{tag}_s{i}_Ts_injected"] = gcm.T_s.copy()
    out["nlat"], out["nlon"], out["dt"] = np.array(nlat), np.array(nlon), np.array(dt)
    # an ocean step that needs several CFL sub-steps (violent winds)
    set_env()
    grid, land, base_alb, fric, Cs_map, forcing, gcm, oc = build_world(nlat, nlon)
    rng = np.random.default_rng(7)
    ua = rng.standard_normal((nlat, nlon)) * 300.0
    va = rng.standard_normal((nlat, nlon)) * 300.0
    oc.uo = rng.standard_normal((nlat, nlon)) * 2.0 * (land == 0)
    oc.vo = rng.standard_normal((nlat, nlon)) * 2.0 * (land == 0)
    oc.eta = rng.standard_normal((nlat, nlon)) * 0.5 * (land == 0)
    Qn = rng.standard_normal((nlat, nlon)) * 200.0
    im = rng.uniform(size=(nlat, nlon)) < 0.2
    opre = snap_oc(oc)
    with quiet():
        oc.step(dt, ua, va, Q_net=Qn, ice_mask=im)
    out["storm_land"], out["storm_ua"], out["storm_va"], out["storm_Q"], out["storm_ice"] = land, ua, va, Qn, im
    for k, v in opre.items():
        out[f"storm_opre_{k}"] = v
    for k, v in snap_oc(oc).items():
        out[f"storm_opost_{k}"] = v
    np.savez_compressed(os.path.join(OUT, "cores_golden.npz"), **out)
    print("cores_golden.npz:", len(out), "arrays")


# ------------------------------------------------------------------------------------ full loop
def gen_loop():
    """Run the UNMODIFIED scripts.run_simulation.main() on a small grid and record the state
    at the end of every step through the plotting hook (plot_state is called at
    run_simulation.py:2428 with precip/albedo/ocean) and the hydrology helpers."""
    import scripts.run_simulation as rs
    import tempfile
    out = {}
    cases = {
        "base": {"QD_ECO_ENABLE": "0", "QD_HYDRO_ENABLE": "0"},
        "banded": {"QD_ECO_ENABLE": "0", "QD_HYDRO_ENABLE": "0", "QD_INIT_BANDED": "1",
                   "QD_INIT_T_POLE": "255.0", "QD_DT_SECONDS": "900", "QD_OROG": "1"},
    }
    nlat, nlon = 31, 60
    for tag, env in cases.items():
        set_env(env)
        os.environ["QD_PLOT_EVERY_DAYS"] = "1e-9"          # plot_interval_steps = 1 -> hook every step
        dt = int(os.environ.get("QD_DT_SECONDS", "300"))
        nsteps = 16
        os.environ["QD_SIM_DAYS"] = repr((nsteps - 0.5) * dt / (2 * np.pi / 8.726646259971648e-5))
        rec = []
        cur = {}

        def hook_plot_state(grid, gcm, land_mask, precip, cloud_cover, albedo, t_days, output_dir, ocean=None, routing=None):
            d = snap_gcm(gcm)
            d.update(snap_oc(ocean))
            d["precip"] = np.array(precip, copy=True)
            d["albedo"] = np.array(albedo, copy=True)
            d["land_mask"] = np.array(land_mask)
            d["friction"] = np.array(gcm.friction_map)
            d["C_snow"] = np.array(gcm.C_snow_map_last, dtype=np.float64)
            d["glacier"] = np.array(gcm.glacier_mask_last)
            d.update(cur)
            rec.append(d)

        real_bucket = rs.update_land_bucket
        real_snow = rs.snowpack_step

        def hook_bucket(W, P_in, E, params, dt_):
            Wn, R = real_bucket(W, P_in, E, params, dt_)
            cur["W_land"], cur["R_bucket"] = np.array(Wn, copy=True), np.array(R, copy=True)
            return Wn, R

        def hook_snow(S_snow, P_snow_land, T_hat_a, params, dt):
            cur["S_snow_in"] = np.array(S_snow, dtype=np.float64, copy=True)
            return real_snow(S_snow, P_snow_land, T_hat_a, params, dt)

        real_gen = rs.generate_base_properties
        statics = {}

        def hook_gen(mask, *a, **k):
            alb, fr = real_gen(mask, *a, **k)
            statics["base_albedo"], statics["friction"] = alb.copy(), fr.copy()
            return alb, fr

        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as tmp, \
                mock.patch.object(rs, "SphericalGrid", lambda n_lat, n_lon: __import__("pygcm.grid", fromlist=["SphericalGrid"]).SphericalGrid(nlat, nlon)), \
                mock.patch.object(rs, "plot_state", hook_plot_state), \
                mock.patch.object(rs, "plot_true_color", lambda *a, **k: None), \
                mock.patch.object(rs, "plot_ecology", lambda *a, **k: None), \
                mock.patch.object(rs, "update_land_bucket", hook_bucket), \
                mock.patch.object(rs, "snowpack_step", hook_snow), \
                mock.patch.object(rs, "generate_base_properties", hook_gen):
            os.chdir(tmp)
            try:
                with quiet():
                    rs.main()
            finally:
                os.chdir(cwd)
        assert len(rec) == nsteps, (len(rec), nsteps)
        for k, v in statics.items():
            out[f"{tag}_{k}"] = v
        out[f"{tag}_land_mask"] = rec[0]["land_mask"]
        out[f"{tag}_dt"] = np.array(dt)
        out[f"{tag}_nsteps"] = np.array(nsteps)
        for i, d in enumerate(rec):
            if i not in (0, 1, 2, 5, 6, 14, 15):      # consecutive pairs for teacher-forced checks
                continue
            for k, v in d.items():
                if k in ("land_mask", "friction", "isr", "olr", "LH_release_last"):
                    continue
                out[f"{tag}_s{i}_{k}"] = v
        print(tag, "steps", len(rec), "max|u|", float(np.abs(rec[-1]["u"]).max()))
    out["nlat"], out["nlon"] = np.array(nlat), np.array(nlon)
    np.savez_compressed(os.path.join(OUT, "loop_golden.npz"), **out)
    print("loop_golden.npz:", len(out), "arrays")


# ------------------------------------------------------------------------------------ ecology sub-daily
def gen_eco():
    """EcologyAdapter.step_subdaily / PopulationManager canopy-cache policy / get_surface_albedo_bands
    (adapter.py:140-186, population.py:252-286,831-915) driven directly, plus the unmodified main()
    loop with QD_ECO_ENABLE=1 (run_simulation.py:1716-1723,2064-2106)."""
    from pygcm.grid import SphericalGrid
    from pygcm.ecology import EcologyAdapter
    import scripts.run_simulation as rs
    import tempfile
    out = {}
    unit_cases = {
        "u1": ({}, (19, 36), 1800.0),
        "u2": ({"QD_ECO_NS": "3", "QD_ECO_SPECIES_WEIGHTS": "0.5,0.3,0.2", "QD_ECO_COHORT_K": "2", "QD_ECO_RAND_SEED": "7",
                "QD_ECO_LIGHT_UPDATE_EVERY_HOURS": "2", "QD_ECO_TOA_TO_SURF_MODE": "rayleigh", "QD_ECO_SOIL_REFLECT": "0.17",
                "QD_ECO_LAI_K": "0.65", "QD_ECO_SPECIES_1_PEAKS": "500:35:0.7, 640:25:0.5", "QD_ECO_SUBSTEP_EVERY_NPHYS": "2"},
               (15, 27), 2700.0),
    }
    for tag, (env, (nlat, nlon), dt) in unit_cases.items():
        set_env(env)
        rng = np.random.default_rng(300 + nlat)
        grid = SphericalGrid(nlat, nlon)
        land = (rng.uniform(size=(nlat, nlon)) < 0.4).astype(np.uint8)
        with quiet():
            eco = EcologyAdapter(grid, land)
        pop = eco.pop
        out[f"{tag}_land"] = land
        out[f"{tag}_dt"] = np.array(dt)
        out[f"{tag}_env"] = np.array(repr(env))
        out[f"{tag}_alpha_leaf_scalar"] = np.array(eco.alpha_leaf_scalar)
        out[f"{tag}_w_b"] = np.array(eco.w_b)
        out[f"{tag}_R_leaf"] = np.array(eco.R_leaf)
        out[f"{tag}_R_species"] = np.array(pop._species_R_leaf)
        out[f"{tag}_species_weights"] = np.array(pop.species_weights)
        out[f"{tag}_lai0"] = np.array(pop.LAI_layers_SK)
        ncalls = 14
        out[f"{tag}_ncalls"] = np.array(ncalls)
        for n in range(ncalls):
            isr = np.maximum(0.0, rng.standard_normal((nlat, nlon)) * 300.0 + 200.0)
            if n == 3:
                isr[1, 2], isr[4, 5] = np.nan, np.inf
            if n == 5:
                pop.LAI_layers_SK = pop.LAI_layers_SK * (1.0 + 0.4 * rng.uniform(size=pop.LAI_layers_SK.shape))
            if n == 9:
                pop.LAI_layers_SK = pop.LAI_layers_SK * 1.01
            out[f"{tag}_c{n}_isr"] = isr
            out[f"{tag}_c{n}_lai"] = np.array(pop.LAI_layers_SK)
            with quiet():
                a = eco.step_subdaily(isr, 0.3, dt)
            out[f"{tag}_c{n}_alpha_is_none"] = np.array(a is None)
            if a is not None:
                out[f"{tag}_c{n}_alpha"] = np.array(a)
            out[f"{tag}_c{n}_E_day"] = np.array(pop.E_day)
            out[f"{tag}_c{n}_f"] = np.array(pop._canopy_f_cached)
            out[f"{tag}_c{n}_snap"] = np.array(pop._lai_snapshot)
            out[f"{tag}_c{n}_clock"] = np.array([pop._hours_accum, pop._next_recompute_hours])
        A, w = eco.get_surface_albedo_bands()
        out[f"{tag}_bands_A"], out[f"{tag}_bands_w"] = np.array(A), np.array(w)

    # ---- the unmodified main() with the ecology sub-daily coupling on
    nlat, nlon = 31, 60
    tag = "loop"
    set_env({"QD_ECO_ENABLE": "1", "QD_HYDRO_ENABLE": "0", "QD_ECO_LIGHT_UPDATE_EVERY_HOURS": "0.25"})
    os.environ["QD_PLOT_EVERY_DAYS"] = "1e-9"
    dt = 300
    nsteps = 8
    os.environ["QD_SIM_DAYS"] = repr((nsteps - 0.5) * dt / (2 * np.pi / 8.726646259971648e-5))
    rec, holder, statics = [], {}, {}
    real_adapter = rs.EcologyAdapter

    def make_adapter(grid, land_mask):
        holder["eco"] = real_adapter(grid, land_mask)
        return holder["eco"]

    def hook_plot_state(grid, gcm, land_mask, precip, cloud_cover, albedo, t_days, output_dir, ocean=None, routing=None):
        d = snap_gcm(gcm)
        d.update(snap_oc(ocean))
        d["precip"], d["albedo"] = np.array(precip, copy=True), np.array(albedo, copy=True)
        d["land_mask"] = np.array(land_mask)
        d["C_snow"] = np.array(gcm.C_snow_map_last, dtype=np.float64)
        d["glacier"] = np.array(gcm.glacier_mask_last)
        pop = holder["eco"].pop
        d["E_day"], d["f_canopy"] = np.array(pop.E_day), np.array(pop._canopy_f_cached)
        d["eco_clock"] = np.array([pop._hours_accum, pop._next_recompute_hours])
        rec.append(d)

    real_gen = rs.generate_base_properties

    def hook_gen(mask, *a, **k):
        alb, fr = real_gen(mask, *a, **k)
        statics["base_albedo"], statics["friction"] = alb.copy(), fr.copy()
        return alb, fr

    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp, \
            mock.patch.object(rs, "SphericalGrid", lambda n_lat, n_lon: __import__("pygcm.grid", fromlist=["SphericalGrid"]).SphericalGrid(nlat, nlon)), \
            mock.patch.object(rs, "plot_state", hook_plot_state), \
            mock.patch.object(rs, "plot_true_color", lambda *a, **k: None), \
            mock.patch.object(rs, "plot_ecology", lambda *a, **k: None), \
            mock.patch.object(rs, "EcologyAdapter", make_adapter), \
            mock.patch.object(rs, "generate_base_properties", hook_gen):
        os.chdir(tmp)
        try:
            with quiet():
                rs.main()
        finally:
            os.chdir(cwd)
    assert len(rec) == nsteps, (len(rec), nsteps)
    eco = holder["eco"]
    for k, v in statics.items():
        out[f"{tag}_{k}"] = v
    out[f"{tag}_land_mask"] = rec[0]["land_mask"]
    out[f"{tag}_dt"], out[f"{tag}_nsteps"] = np.array(dt), np.array(nsteps)
    out[f"{tag}_alpha_leaf_scalar"] = np.array(eco.alpha_leaf_scalar)
    out[f"{tag}_lai"] = np.array(eco.pop.LAI_layers_SK)
    for i, d in enumerate(rec):
        for k, v in d.items():
            if k in ("land_mask", "isr", "olr", "LH_release_last"):
                continue
            out[f"{tag}_s{i}_{k}"] = v
    out["loop_nlat"], out["loop_nlon"] = np.array(nlat), np.array(nlon)
    np.savez_compressed(os.path.join(OUT, "eco_golden.npz"), **out)
    print("eco_golden.npz:", len(out), "arrays")


# ------------------------------------------------------------------------------------ phytoplankton transport
def gen_phyto():
    """PhytoManager.advect_diffuse (pygcm/ecology/phyto.py:496-547): per-species semi-Lagrangian advection by the
    ocean currents + explicit lateral diffusion + polar ring means, called every physics step by the script
    (run_simulation.py:2256-2258).  SURVEY 8f row 1."""
    from pygcm.grid import SphericalGrid
    from pygcm.ecology.phyto import PhytoManager
    out = {}
    for tag, (nlat, nlon, dt, env) in {"p1": (25, 48, 300.0, {}),
                                       "p2": (19, 36, 1800.0, {"QD_PHYTO_ADV_ALPHA": "0.45", "QD_PHYTO_KH": "2.0e4"})}.items():
        set_env(env)
        rng = np.random.default_rng(500 + nlat)
        grid = SphericalGrid(nlat, nlon)
        land = (rng.uniform(size=(nlat, nlon)) < 0.3).astype(np.uint8)
        if tag == "p2":
            land[0, :] = 1                                           # no ocean on the south pole row
        with quiet():
            ph = PhytoManager(grid, land, H_mld_m=50.0, diag=False)
        S = int(ph.S)
        C = np.abs(rng.standard_normal((S, nlat, nlon))) * 0.3 * (land == 0)
        C[0, 3, 4] = np.nan
        ph.C_phyto_s = C.copy()
        out[f"{tag}_land"], out[f"{tag}_dt"], out[f"{tag}_C0"] = land, np.array(dt), C
        out[f"{tag}_kh"] = np.array(ph.K_h)
        out[f"{tag}_alpha"] = np.array(float(env.get("QD_PHYTO_ADV_ALPHA", "0.7")))
        ncalls = 4
        out[f"{tag}_ncalls"] = np.array(ncalls)
        for n in range(ncalls):
            uo = rng.standard_normal((nlat, nlon)) * 0.8 * (land == 0)
            vo = rng.standard_normal((nlat, nlon)) * 0.5 * (land == 0)
            if n == 2:
                uo[5, 6], vo[7, 8] = 40.0, -25.0                     # departure points several cells away
            with quiet(), np.errstate(all="ignore"):
                ph.advect_diffuse(uo, vo, dt)
            out[f"{tag}_c{n}_uo"], out[f"{tag}_c{n}_vo"] = uo, vo
            out[f"{tag}_c{n}_C"] = np.array(ph.C_phyto_s, copy=True)
    np.savez_compressed(os.path.join(OUT, "phyto_golden.npz"), **out)
    print("phyto_golden.npz:", len(out), "arrays")


# ------------------------------------------------------------------------------------ individual pool sub-steps
def gen_indiv():
    """IndividualPool.try_substep (pygcm/ecology/individuals.py:142-191) with the NB=16 band split of the dual-star
    insolation (spectral.dual_star_insolation_to_bands, spectral.py:304-426).  SURVEY 8f row 2."""
    from pygcm.grid import SphericalGrid
    from pygcm.ecology import EcologyAdapter
    from pygcm.ecology.individuals import IndividualPool
    from pygcm.ecology.spectral import dual_star_insolation_to_bands
    out = {}
    cases = {"i1": ({}, (19, 36)),
             "i2": ({"QD_ECO_NS": "4", "QD_ECO_SPECIES_WEIGHTS": "0.4,0.3,0.2,0.1", "QD_ECO_RAND_SEED": "3", "QD_ECO_TOA_TO_SURF_MODE": "rayleigh",
                     "QD_ECO_INDIV_SAMPLE_FRAC": "0.2", "QD_ECO_INDIV_PER_CELL": "7", "QD_ECO_INDIV_SUBSTEPS_PER_DAY": "24",
                     "QD_ECO_SPECIES_2_PEAKS": "520:30:0.9", "QD_ECO_SPECIES_1_DROUGHT_TOL": "0.7"}, (15, 27))}
    for tag, (env, (nlat, nlon)) in cases.items():
        set_env(env)
        rng = np.random.default_rng(700 + nlat)
        grid = SphericalGrid(nlat, nlon)
        land = (rng.uniform(size=(nlat, nlon)) < 0.45).astype(np.uint8)
        with quiet():
            eco = EcologyAdapter(grid, land)
            pool = IndividualPool(grid, land, eco, diag=False)
        out[f"{tag}_env"] = np.array(repr(env))
        out[f"{tag}_land"] = land
        for k in ("sample_j", "sample_i", "indiv_cell_index", "indiv_species_id", "indiv_Ab", "indiv_tol", "sp_weights"):
            out[f"{tag}_{k}"] = np.array(getattr(pool, k))
        out[f"{tag}_cfg"] = np.array([pool.cfg.sample_frac, pool.cfg.per_cell, pool.cfg.substeps_per_day, pool.nb], dtype=float)
        dt, day = 1800.0, 72000.0
        out[f"{tag}_dt"], out[f"{tag}_day"] = np.array(dt), np.array(day)
        ncalls = 12
        out[f"{tag}_ncalls"] = np.array(ncalls)
        for n in range(ncalls):
            isrA = np.maximum(0.0, rng.standard_normal((nlat, nlon)) * 250.0 + 150.0)
            isrB = np.maximum(0.0, rng.standard_normal((nlat, nlon)) * 120.0 + 20.0)
            if n == 4:
                isrA[:, : nlon // 2] = 0.0
                isrB[:, : nlon // 2] = 0.0                             # night side
            soil = rng.uniform(0.0, 1.0, (nlat, nlon))
            out[f"{tag}_c{n}_isrA"], out[f"{tag}_c{n}_isrB"], out[f"{tag}_c{n}_soil"] = isrA, isrB, soil
            e0 = pool.indiv_E_day.copy()
            pool.try_substep(isrA, isrB, eco, soil, dt, day)
            out[f"{tag}_c{n}_fired"] = np.array(not np.array_equal(e0, pool.indiv_E_day) or n == 4)
            out[f"{tag}_c{n}_E"] = pool.indiv_E_day.copy()
            out[f"{tag}_c{n}_stress"] = pool.indiv_water_stress_days.copy()
            out[f"{tag}_c{n}_accum"] = np.array(pool._substep_accum)
        Ib = dual_star_insolation_to_bands(isrA, isrB, eco.bands)
        out[f"{tag}_Ib_last"] = Ib
    np.savez_compressed(os.path.join(OUT, "indiv_golden.npz"), **out)
    print("indiv_golden.npz:", len(out), "arrays")


def gen_routing():
    """Network from the reference's own builder (scripts/generate_hydrology_maps.py:85-273) on a small
    procedural elevation, then pygcm.routing.RiverRouting (unmodified, fed through an in-memory stand-in
    for netCDF4.Dataset) driven for several events with random runoff / precip / evaporation."""
    import scripts.generate_hydrology_maps as hm
    import pygcm.routing as routing
    from pygcm.grid import SphericalGrid
    from pygcm.topography import generate_elevation_map, create_land_sea_mask_from_elevation
    out = {}
    set_env()
    for tag, (nlat, nlon) in {"r1": (25, 48), "r2": (31, 60)}.items():
        grid = SphericalGrid(nlat, nlon)
        with quiet():
            elev = generate_elevation_map(grid, seed=11 + nlat)
            land, sea = create_land_sea_mask_from_elevation(elev, grid, target_land_frac=0.45)
            elev = elev - sea
            filled = hm.pit_fill(elev.copy(), land.astype(np.uint8), max_iters=200, eps=1e-3)
            flow_to = hm.compute_flow_to_index(grid, filled, land)
            lake_mask, lake_id, n_lakes = hm.identify_lakes(flow_to, land)
            outlet = hm.compute_lake_outlets(grid, filled, lake_mask, lake_id, land) if n_lakes > 0 else None
            order = hm.topo_sort_flow_order(flow_to, land)
        net = {"land_mask": land.astype(np.uint8), "flow_to_index": flow_to.astype(np.int32),
               "flow_order": order.astype(np.int32), "lake_mask": lake_mask.astype(np.uint8), "lake_id": lake_id.astype(np.int32)}
        if n_lakes > 0:
            net["lake_outlet_index"] = outlet.astype(np.int32)
        path = f"/virtual/{tag}.nc"
        MemDataset.store[path] = net
        for k, v in net.items():
            out[f"{tag}_{k}"] = v
        out[f"{tag}_n_lakes"] = np.array(n_lakes)
        out[f"{tag}_elev_filled"] = filled
        out[f"{tag}_elev_in"] = elev
        rng = np.random.default_rng(5)
        dt, nsub = 1800.0, 4                                  # dt_hydro = 2 h -> event every 4 steps
        with mock.patch.object(routing, "Dataset", MemDataset), mock.patch.object(routing.os.path, "exists", lambda p: True), quiet():
            rr = routing.RiverRouting(grid, path, dt_hydro_hours=2.0, diag=False)
            ev = 0
            for step in range(12):
                R = np.where(land == 1, rng.uniform(0, 2e-5, (nlat, nlon)), 0.0)
                P = rng.uniform(0, 5e-5, (nlat, nlon))
                E = rng.uniform(0, 4e-5, (nlat, nlon))
                out[f"{tag}_R{step}"], out[f"{tag}_P{step}"], out[f"{tag}_E{step}"] = R, P, E
                rr.step(R, dt, precip_flux=P, evap_flux=E)
                if (step + 1) % nsub == 0:
                    d = rr.diagnostics()
                    out[f"{tag}_ev{ev}_flow"] = d["flow_accum_kgps"].copy()
                    out[f"{tag}_ev{ev}_ocean"] = np.array(d["ocean_inflow_kgps"])
                    out[f"{tag}_ev{ev}_err"] = np.array(d["mass_closure_error_kg"])
                    if d["lake_volume_kg"] is not None:
                        out[f"{tag}_ev{ev}_lake"] = d["lake_volume_kg"].copy()
                    ev += 1
        out[f"{tag}_dt"] = np.array(dt)
        print(tag, "land", int(land.sum()), "lakes", n_lakes, "ocean inflow", float(out[f"{tag}_ev2_ocean"]))
    np.savez_compressed(os.path.join(OUT, "routing_golden.npz"), **out)
    print("routing_golden.npz:", len(out), "arrays")


def gen_restart():
    """The reference's OWN file helpers -- run_simulation.save_restart / load_restart / save_topography / save_ocean /
    load_ocean (scripts/run_simulation.py:63-246) and topography.load_topography_from_netcdf (pygcm/topography.py:
    428-575) -- run unmodified with qingdai_b200.ncio installed as ``netCDF4`` (this container has no netCDF4).
    Records what they return for seeded inputs, including the regrid of a coarse topography onto a finer grid, so that
    qingdai_b200.restart can be compared without the reference."""
    import tempfile
    sys.path.insert(0, ROOT)
    from qingdai_b200 import ncio
    assert ncio.install_netcdf4_shim(), "a real netCDF4 is installed: record with it instead"
    set_env()
    from pygcm.grid import SphericalGrid
    from pygcm import topography as topo_ref
    import scripts.run_simulation as rs
    rng = np.random.default_rng(2024)
    g = SphericalGrid(13, 24)
    shape = (13, 24)
    ns = lambda: types.SimpleNamespace()          # noqa: E731
    gcm, oc = ns(), ns()
    for k in ("u", "v", "h", "T_s", "cloud_cover", "q", "h_ice"):
        setattr(gcm, k, rng.standard_normal(shape) * 10 + 100)
    for k in ("uo", "vo", "eta", "Ts"):
        setattr(oc, k, rng.standard_normal(shape))
    land = (rng.uniform(size=shape) < 0.4).astype(np.uint8)
    W, S_ = rng.uniform(size=shape), rng.uniform(size=shape)
    out = {"in_land": land, "in_W": W, "in_S": S_}
    for k in ("u", "v", "h", "T_s", "cloud_cover", "q", "h_ice"):
        out["in_gcm_" + k] = getattr(gcm, k)
    for k in ("uo", "vo", "eta", "Ts"):
        out["in_oc_" + k] = getattr(oc, k)
    with tempfile.TemporaryDirectory() as td, quiet():
        rp = os.path.join(td, "restart.nc")
        rs.save_restart(rp, g, gcm, oc, land, W_land=W, S_snow=S_, C_snow=None, t_seconds=123456.789)
        r = rs.load_restart(rp)
        for k, v in r.items():
            if v is not None:
                out["restart_" + k] = np.asarray(v)
        out["restart_none"] = np.array([k for k, v in r.items() if v is None])
        op = os.path.join(td, "ocean.nc")
        rs.save_ocean(op, g, oc, day_value=12.5)
        o = rs.load_ocean(op)
        for k, v in o.items():
            out["ocean_" + k] = np.asarray(v)
        # topography: coarse source (seam column included, as grid.lon has it) -> same grid and a finer grid
        elev = rng.uniform(-500, 3000, shape)
        alb = rng.uniform(0.05, 0.5, shape)
        fric = rng.uniform(1e-6, 1e-4, shape)
        tp = os.path.join(td, "topography.nc")
        rs.save_topography(tp, g, land, alb, fric, elevation=elev)
        out.update(in_elev=elev, in_alb=alb, in_fric=fric)
        for tag, gt in (("same", g), ("fine", SphericalGrid(19, 40))):
            e2, m2, a2, f2 = topo_ref.load_topography_from_netcdf(tp, gt)
            out.update({f"topo_{tag}_elev": e2, f"topo_{tag}_mask": m2, f"topo_{tag}_alb": a2, f"topo_{tag}_fric": f2})
        # a foreign layout: latitude descending, longitude in [-180, 180) without a seam column, coarser than the target
        fl_lat = np.linspace(90.0, -90.0, 10)
        fl_lon = np.linspace(-180.0, 180.0, 18, endpoint=False)
        f_elev = rng.uniform(-500, 3000, (10, 18))
        f_mask = (rng.uniform(size=(10, 18)) < 0.5).astype(np.uint8)
        f_alb = rng.uniform(0.05, 0.5, (10, 18))
        f_fric = rng.uniform(1e-6, 1e-4, (10, 18))
        fp = os.path.join(td, "foreign.nc")
        with ncio.Dataset(fp, "w") as ds:
            ds.createDimension("lat", 10)
            ds.createDimension("lon", 18)
            v = ds.createVariable("lat", "f8", ("lat",)); v[:] = fl_lat
            v = ds.createVariable("lon", "f8", ("lon",)); v[:] = fl_lon
            for name, dt_, arr in (("elevation", "f8", f_elev), ("land_mask", "u1", f_mask), ("base_albedo", "f8", f_alb), ("friction", "f8", f_fric)):
                v = ds.createVariable(name, dt_, ("lat", "lon")); v[:] = arr
        out.update(foreign_lat=fl_lat, foreign_lon=fl_lon, foreign_elev=f_elev, foreign_mask=f_mask, foreign_alb=f_alb, foreign_fric=f_fric)
        e2, m2, a2, f2 = topo_ref.load_topography_from_netcdf(fp, SphericalGrid(13, 24))
        out.update(topo_foreign_elev=e2, topo_foreign_mask=m2, topo_foreign_alb=a2, topo_foreign_fric=f2)
    np.savez_compressed(os.path.join(OUT, "restart_golden.npz"), **out)
    print("restart_golden.npz:", len(out), "arrays")


# ------------------------------------------------------------------------------------ reference-generated inputs
def gen_topo():
    """The bench / parity inputs SURVEY 8(d) prescribes, produced by the reference's own generators:
    * ``topography_qingdai_181x360_seed42.nc``: scripts/generate_topography.py defaults (:60-61 -- 181x360, seed 42,
      land 0.40; elevation parameters :64-75) through pygcm.topography.generate_elevation_map /
      create_land_sea_mask_from_elevation / generate_base_properties / export_topography_to_netcdf, written through the
      NetCDF-3 shim (float32 variables, exactly what the reference's exporter stores and what QD_TOPO_NC reads back);
    * ``default_mask_181x360.npz``: the built-in path of scripts/run_simulation.py:1212-1213 for BASELINE configs[0]
      (create_land_sea_mask(grid) = land 0.29, seed 42, and generate_base_properties(mask))."""
    sys.path.insert(0, ROOT)
    from qingdai_b200 import ncio
    assert ncio.install_netcdf4_shim(), "a real netCDF4 is installed: record with it instead"
    set_env()
    from pygcm.grid import SphericalGrid
    from pygcm import topography as T
    grid = SphericalGrid(n_lat=181, n_lon=360)
    params = {"N_CONTINENTS": 3, "CONTINENT_SIGMA_DEG": 30.0, "CONTINENT_SHAPE_P": 2.0, "CONT_MIN_DIST_DEG": 40.0, "W_VLF": 0.35,
              "FBM_OCTAVES": 5, "HURST_H": 0.8, "W1": 1.0, "W3": 0.6, "SCALE_M": 4500.0}
    with quiet():
        elevation = T.generate_elevation_map(grid, seed=42, params=params)
        land_mask, sea_level = T.create_land_sea_mask_from_elevation(elevation, grid, target_land_frac=0.40)
        base_albedo, friction = T.generate_base_properties(land_mask, elevation=elevation, grid=grid)
        path = os.path.join(OUT, "topography_qingdai_181x360_seed42.nc")
        T.export_topography_to_netcdf(grid=grid, elevation=elevation, land_mask=land_mask, base_albedo=base_albedo,
                                      friction=friction, sea_level_m=sea_level, out_path=path)
        # what the reference's loader returns for it on the same grid: pins qingdai_b200.restart.load_topography_from_netcdf
        e2, m2, a2, f2 = T.load_topography_from_netcdf(path, grid)
        mask0 = T.create_land_sea_mask(grid)
        alb0, fric0 = T.generate_base_properties(mask0)
    w = np.cos(np.deg2rad(grid.lat_mesh))
    print("topography_qingdai_181x360_seed42.nc: land fraction", float((w * (land_mask == 1)).sum() / w.sum()), "sea level", float(sea_level),
          "bytes", os.path.getsize(path))
    np.savez_compressed(os.path.join(OUT, "default_mask_181x360.npz"), land_mask=mask0.astype(np.uint8), base_albedo=alb0, friction=fric0,
                        topo_loaded_elev_sum=np.array(float(np.sum(e2))), topo_loaded_mask_sum=np.array(int(np.sum(m2))),
                        topo_loaded_alb_sum=np.array(float(np.sum(a2))), topo_loaded_fric_sum=np.array(float(np.sum(f2))))
    print("default_mask_181x360.npz: land fraction", float((w * (mask0 == 1)).sum() / w.sum()))


def main():
    _install_stubs()
    which = sys.argv[1:] or ["ops", "cores", "loop"]
    if "ops" in which:
        gen_ops()
    if "cores" in which:
        gen_cores()
    if "loop" in which:
        gen_loop()
    if "routing" in which:
        gen_routing()
    if "eco" in which:
        gen_eco()
    if "phyto" in which:
        gen_phyto()
    if "indiv" in which:
        gen_indiv()
    if "restart" in which:
        gen_restart()
    if "topo" in which:
        gen_topo()


if __name__ == "__main__":
    main()
