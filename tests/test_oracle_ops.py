"""Pin the oracle's operators to the reference outputs (tests/golden/ops_golden.npz, produced by
tests/golden/make_golden.py from the imported reference) and, where SciPy is present, to SciPy."""
import numpy as np
import pytest

from conftest import relerr
from oracle import ops, model
from qingdai_b200.params import QDParams

TOL = 1e-13


@pytest.fixture(scope="module")
def G(golden):
    return golden("ops_golden.npz")


@pytest.mark.parametrize("tag,shape", [("a", (22, 40)), ("b", (15, 27))])
def test_ops_vs_reference(G, tag, shape):
    g = model.make_grid(*shape)
    p = QDParams()
    F, u, v, dt = G[f"{tag}_F"], G[f"{tag}_u"], G[f"{tag}_v"], float(G[f"{tag}_dt"])
    a = g.a
    c_atm, c_oc, c_lap = np.maximum(1e-6, g.cos), np.maximum(g.cos, 0.5), np.maximum(g.cos, 0.2)
    # gathers: bit-exact
    assert np.array_equal(ops.advect_semilag(F, u, v, dt, a, g.dlat, g.dlon, c_atm), G[f"{tag}_adv_atm"])
    assert np.array_equal(ops.advect_semilag(F, u * 0.01, v * 0.01, dt, a, g.dlat, g.dlon, c_oc), G[f"{tag}_adv_oc"])
    assert np.array_equal(ops.advect_semilag(F, u, v, dt, a, g.dlat, g.dlon, c_oc), G[f"{tag}_adv_cloud"])
    # stencils: bit-exact (same operand order)
    assert np.array_equal(ops.laplacian(F, g.dlat, g.dlon, c_lap, a), G[f"{tag}_lap_atm"])
    assert np.array_equal(ops.laplacian(F, g.dlat, g.dlon, c_oc, a), G[f"{tag}_lap_oc"])
    k4 = G[f"{tag}_k4map"]
    assert np.array_equal(ops.hyperdiffuse(F, k4, dt, 1, g.dlat, g.dlon, c_lap, a), G[f"{tag}_hyp_atm_map"])
    assert np.array_equal(ops.hyperdiffuse(F, 0.5 * k4, dt, 3, g.dlat, g.dlon, c_lap, a), G[f"{tag}_hyp_atm_map3"])
    assert np.array_equal(ops.hyperdiffuse(F, 1.0e14, dt, 2, g.dlat, g.dlon, c_lap, a), G[f"{tag}_hyp_atm_scalar"])
    assert np.array_equal(ops.hyperdiffuse(F, k4, dt, 1, g.dlat, g.dlon, c_oc, a), G[f"{tag}_hyp_oc_map"])
    with np.errstate(all="ignore"):
        assert np.array_equal(ops.laplacian(G[f"{tag}_Fnan"], g.dlat, g.dlon, c_lap, a), G[f"{tag}_lap_nan"], equal_nan=True)
    assert np.array_equal(ops.shapiro(F, 2), G[f"{tag}_shapiro2"])
    assert np.array_equal(ops.shapiro(F, 1), G[f"{tag}_shapiro1"])
    assert np.array_equal(ops.zonal_bandstop(F, 0.75, 0.5), G[f"{tag}_spec"])
    assert np.array_equal(ops.zonal_bandstop(F, 0.3, 1.0), G[f"{tag}_spec2"])
    assert np.array_equal(ops.divergence(u, v, g.lat, g.dlat, g.dlon, a), G[f"{tag}_div"])
    assert np.array_equal(ops.vorticity(u, v, g.lat, g.dlat, g.dlon, a), G[f"{tag}_vort"])
    assert np.array_equal(ops.gaussian(F, 1.0), G[f"{tag}_gauss1"])
    assert np.array_equal(ops.gaussian(F, 0.2, "wrap"), G[f"{tag}_gauss02w"])
    assert ops.median_pos(np.maximum(0.0, F - 280.0)) == float(G[f"{tag}_median_pos"])


@pytest.mark.parametrize("tag,shape", [("a", (22, 40)), ("b", (15, 27))])
def test_cell_physics_vs_reference(G, tag, shape):
    g = model.make_grid(*shape)
    p = QDParams()
    X = lambda k: G[f"{tag}_{k}"]
    land, dt = X("land"), float(X("dt"))
    Ts, Ta, q, cloud, hice, isr, alb, u, v = (X(k) for k in ("Ts", "Ta", "q", "cloud", "hice", "isr", "alb", "u", "v"))
    assert np.array_equal(model.q_sat(Ts, p.p0), X("qsat"))
    fac = model.evap_factor(land, hice, p)
    assert np.array_equal(fac, X("evapfac"))
    sa, ss, R = model.shortwave(isr, alb, cloud, p)
    assert np.array_equal(sa, X("sw_atm")) and np.array_equal(ss, X("sw_sfc")) and np.array_equal(R, X("sw_R"))
    ice_frac = 1.0 - np.exp(-np.maximum(hice, 0.0) / 0.5)
    eps = model.emissivity_map(land, ice_frac, p)
    assert np.array_equal(eps, X("eps_sfc"))
    la, ls, olr, _, _ = model.longwave_v2(Ts, Ta, cloud, eps, p)
    assert np.array_equal(la, X("lw2_atm")) and np.array_equal(ls, X("lw2_sfc")) and np.array_equal(olr, X("lw2_olr"))
    la, ls, olr, _, _ = model.longwave_v1(Ts, Ta, cloud, p)
    assert np.array_equal(la, X("lw1_atm")) and np.array_equal(ls, X("lw1_sfc")) and np.array_equal(olr, X("lw1_olr"))
    SH = model.sensible_heat(Ts, Ta, u, v, p)
    assert np.array_equal(SH, X("SH"))
    Tn, hn = model.seaice_integrate(Ts, ss, X("lw1_sfc"), SH, 2.5e6 * X("E"), dt, land, hice, p)
    assert np.array_equal(Tn, X("seaice_Ts")) and np.array_equal(hn, X("seaice_h"))
    assert np.array_equal(model.dynamic_albedo(cloud, alb * 0.5, ice_frac, land, p), X("alb_dyn"))
    # hydrology
    st = type("S", (), {})()
    Wn, Rf = model.land_bucket(X("W"), X("Prain") * land, X("E") * land, p, dt)
    assert np.array_equal(Wn, X("Wnext")) and np.array_equal(Rf, X("Rflux"))


@pytest.mark.parametrize("tag,shape", [("a", (22, 40)), ("b", (15, 27))])
def test_composite_physics_vs_reference(G, tag, shape):
    g = model.make_grid(*shape)
    p = QDParams()
    X = lambda k: G[f"{tag}_{k}"]
    st = type("S", (), {})()
    st.T_s, st.u, st.v, st.cloud = X("Ts"), X("u"), X("v"), X("cloud")
    st.P_cond = X("Pcond")
    assert relerr(model.cloud_source(st, g), X("cloud_src")) == 0.0
    assert relerr(model.orographic_factor(g, X("elev"), st.u, st.v, p), X("orog")) == 0.0
    assert relerr(model.precip_hybrid(st, g, p, None), X("precip_hyb")) < TOL
    assert relerr(model.precip_hybrid(st, g, p, X("orog")), X("precip_hyb_orog")) < TOL
    st.P_cond = X("Pcond") * 1e-9
    assert relerr(model.precip_hybrid(st, g, p, None), X("precip_hyb_fb")) < TOL


def test_against_scipy_if_present():
    ndi = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(3)
    for shape in [(9, 16), (31, 40), (5, 7)]:
        F = rng.standard_normal(shape) * 50
        J = rng.uniform(-4 * shape[0], 5 * shape[0], shape)
        I = rng.uniform(-4 * shape[1], 5 * shape[1], shape)
        J[1] *= rng.uniform(0, 1e-3, shape[1])      # sub-2^-53 fractions: exposes w1 = 1-(1-t)
        I[2] *= rng.uniform(0, 1e-3, shape[1])
        J.flat[:4] = [0.0, shape[0] - 1.0, -(shape[0] - 1.0), 2.0 * (shape[0] - 1)]
        I.flat[:4] = [shape[1] - 1.0, 0.0, 3.0 * (shape[1] - 1), -0.0]
        assert np.array_equal(ops.bilinear_wrap(F, J, I), ndi.map_coordinates(F, [J, I], order=1, mode="wrap", prefilter=False))
        for sigma, mode in [(1.0, "reflect"), (0.5, "reflect"), (0.2, "wrap"), (1.0, "wrap")]:
            assert np.array_equal(ops.gaussian(F, sigma, mode), ndi.gaussian_filter(F, sigma=sigma, mode=mode))
        k1 = np.array([0.25, 0.5, 0.25])
        ref = ndi.convolve(ndi.convolve(F, k1[None, :], mode="wrap"), k1[:, None], mode="nearest")
        assert np.array_equal(ops.shapiro(F, 1), ref)
