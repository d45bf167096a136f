"""Pin oracle.model.atmos_step / ocean_step / loop_step to trajectories recorded from the reference
(tests/golden/cores_golden.npz, loop_golden.npz; generator: tests/golden/make_golden.py)."""
import numpy as np
import pytest

from conftest import relerr
from oracle import model
from qingdai_b200.params import QDParams

TOL = 1e-12
ATM = {"u": "u", "v": "v", "h": "h", "T_s": "T_s", "q": "q", "cloud_cover": "cloud", "h_ice": "h_ice"}
DIAG = {"olr": "olr", "E_flux_last": "E_flux", "P_cond_flux_last": "P_cond", "LH_last": "LH",
        "LH_release_last": "LH_release"}
CASES = {"w1": dict(energy_w=1.0), "w0": dict(),
         "w05k": dict(energy_w=0.5, k4_nsub=2, spec_every=3, mom_scheme="primitive")}
KEEP = {"w1": (0, 5, 12, 13), "w0": (0, 5), "w05k": (1, 2)}


@pytest.fixture(scope="module")
def C(golden):
    return golden("cores_golden.npz")


def load_atm(C, tag, i, g, p):
    st = model.new_atmos_state(g, p, C[f"{tag}_land"], C[f"{tag}_fric"], base_albedo=C[f"{tag}_base_alb"],
                               C_s_map=np.where(C[f"{tag}_land"] == 1, 3e6, 2.1e8).astype(float))
    for k, mine in {**ATM, **DIAG}.items():
        setattr(st, mine, C[f"{tag}_s{i}_pre_{k}"].copy())
    st.isr = C[f"{tag}_s{i}_post_isr"].copy()
    key = f"{tag}_s{i}_pre_cloud_eff_last"
    st.cloud_eff = C[key].copy() if key in C.files else None
    st.step_counter = int(C[f"{tag}_s{i}_counter_pre"])
    return st


@pytest.mark.parametrize("tag", list(CASES))
def test_atmos_step_vs_reference(C, tag):
    g = model.make_grid(int(C["nlat"]), int(C["nlon"]))
    p = QDParams(**CASES[tag])
    dt = float(C["dt"])
    for i in KEEP[tag]:
        st = load_atm(C, tag, i, g, p)
        alb = C[f"{tag}_s{i}_albedo"] if int(C[f"{tag}_s{i}_use_alb"]) else None
        model.atmos_step(st, g, p, C[f"{tag}_s{i}_Teq"], dt, albedo=alb)
        for k, mine in {**ATM, **DIAG}.items():
            e = relerr(getattr(st, mine), C[f"{tag}_s{i}_post_{k}"])
            assert e < TOL, (tag, i, k, e)
        if alb is not None:
            assert relerr(st.cloud_eff, C[f"{tag}_s{i}_post_cloud_eff_last"]) < TOL


@pytest.mark.parametrize("tag", list(CASES))
def test_ocean_step_vs_reference(C, tag):
    g = model.make_grid(int(C["nlat"]), int(C["nlon"]))
    p = QDParams(**CASES[tag])
    dt = float(C["dt"])
    for i in KEEP[tag]:
        oc = model.new_ocean_state(g, C[f"{tag}_land"])
        for k in ("uo", "vo", "eta", "Ts"):
            setattr(oc, k, C[f"{tag}_s{i}_opre_{k}"].copy())
        oc.step = i
        model.ocean_step(oc, g, p, dt, C[f"{tag}_s{i}_post_u"], C[f"{tag}_s{i}_post_v"],
                         Q_net=C[f"{tag}_s{i}_Qnet"], ice_mask=C[f"{tag}_s{i}_ice_mask"])
        for k in ("uo", "vo", "eta", "Ts"):
            e = relerr(getattr(oc, k), C[f"{tag}_s{i}_opost_{k}"])
            assert e < TOL, (tag, i, k, e)


def test_ocean_substeps_vs_reference(C):
    """Violent winds -> several CFL sub-steps (ocean.py:293-303) and the mean4 outlier path."""
    g = model.make_grid(int(C["nlat"]), int(C["nlon"]))
    p = QDParams()
    oc = model.new_ocean_state(g, C["storm_land"])
    for k in ("uo", "vo", "eta", "Ts"):
        setattr(oc, k, C[f"storm_opre_{k}"].copy())
    model.ocean_step(oc, g, p, float(C["dt"]), C["storm_ua"], C["storm_va"], Q_net=C["storm_Q"], ice_mask=C["storm_ice"])
    assert oc.n_sub_last > 1
    for k in ("uo", "vo", "eta", "Ts"):
        assert relerr(getattr(oc, k), C[f"storm_opost_{k}"]) < TOL, k


def test_qnet_vs_reference(C):
    """run_simulation.py:2207-2239 / benchmark_jax.py:135-152 surface heat flux."""
    g = model.make_grid(int(C["nlat"]), int(C["nlon"]))
    p = QDParams(energy_w=1.0)
    for i in KEEP["w1"]:
        st = load_atm(C, "w1", i, g, p)
        for k, mine in {**ATM, **DIAG}.items():
            setattr(st, mine, C[f"w1_s{i}_post_{k}"].copy())
        st.cloud_eff = C[f"w1_s{i}_post_cloud_eff_last"]
        Q, ice = model.surface_qnet(st, g, p, C[f"w1_s{i}_albedo"])
        assert relerr(Q, C[f"w1_s{i}_Qnet"]) < TOL
        assert np.array_equal(ice, C[f"w1_s{i}_ice_mask"])


# ------------------------------------------------------------------------------------ full loop
LOOP_ENV = {"base": {}, "banded": {"QD_OROG": "1"}}


def _loop_state(L, tag, i, g, p):
    land = L[f"{tag}_land_mask"]
    st = model.new_atmos_state(g, p, land, L[f"{tag}_friction"], base_albedo=L[f"{tag}_base_albedo"])
    oc = model.new_ocean_state(g, land)
    if i >= 0:
        X = lambda k: L[f"{tag}_s{i}_{k}"].copy()
        for k, mine in ATM.items():
            setattr(st, mine, X(k))
        st.E_flux, st.P_cond, st.LH = X("E_flux_last"), X("P_cond_flux_last"), X("LH_last")
        st.W_land = X("W_land")
        st.S_snow = L[f"{tag}_s{i + 1}_S_snow_in"].copy()
        for k in ("uo", "vo", "eta", "Ts"):
            setattr(oc, k, X(k))
        st.step_counter = i + 1
        oc.step = i + 1
    return st, oc


@pytest.mark.parametrize("tag", ["base", "banded"])
def test_loop_teacher_forced(golden, tag):
    """state(end of step i) from the reference -> one oracle loop_step -> reference state(end of i+1)."""
    L = golden("loop_golden.npz")
    g = model.make_grid(int(L["nlat"]), int(L["nlon"]))
    p = QDParams.from_env(LOOP_ENV[tag])
    dt = float(L[f"{tag}_dt"])
    for i in (0, 1, 5, 14):
        st, oc = _loop_state(L, tag, i, g, p)
        out = model.loop_step(st, oc, g, p, t=(i + 1) * dt, dt=dt)
        X = lambda k: L[f"{tag}_s{i + 1}_{k}"]
        assert relerr(out.precip, X("precip")) < TOL, (i, "precip")
        assert relerr(out.albedo, X("albedo")) < TOL, (i, "albedo")
        for k, mine in ATM.items():
            assert relerr(getattr(st, mine), X(k)) < TOL, (i, k)
        for k in ("uo", "vo", "eta", "Ts"):
            assert relerr(getattr(oc, k), X(k)) < TOL, (i, k)
        assert relerr(st.W_land, X("W_land")) < TOL
        assert relerr(out.C_snow, X("C_snow")) < TOL
        assert np.array_equal(out.glacier, X("glacier"))


@pytest.mark.parametrize("tag", ["base", "banded"])
def test_loop_free_running(golden, tag):
    """From the reference's own initial state, 16 free-running oracle steps track main() to 1e-9."""
    L = golden("loop_golden.npz")
    g = model.make_grid(int(L["nlat"]), int(L["nlon"]))
    p = QDParams.from_env(LOOP_ENV[tag])
    dt = float(L[f"{tag}_dt"])
    st, oc = _loop_state(L, tag, -1, g, p)
    if tag == "banded":     # apply_banded_initial_ts, run_simulation.py:310-328
        Ts0 = 255.0 + (295.0 - 255.0) * (np.cos(np.deg2rad(g.lat)) ** 2)[:, None] * np.ones((g.nlat, g.nlon))
        st.T_s = Ts0.copy()
        oc.Ts = np.where(st.land_mask == 0, Ts0, oc.Ts)
    else:
        oc.Ts = np.where(st.land_mask == 0, st.T_s, 288.0)
    for i in range(int(L[f"{tag}_nsteps"])):
        model.loop_step(st, oc, g, p, t=i * dt, dt=dt)
        if i in (0, 1, 2, 5, 6, 14, 15):
            for k, mine in ATM.items():
                assert relerr(getattr(st, mine), L[f"{tag}_s{i}_{k}"]) < 1e-9, (i, k)
            for k in ("uo", "vo", "eta", "Ts"):
                assert relerr(getattr(oc, k), L[f"{tag}_s{i}_{k}"]) < 1e-9, (i, k)
