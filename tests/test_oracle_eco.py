"""Pin oracle.ecology (sub-daily ecology: E_day, canopy-cache policy, land alpha, band albedo) and the
oracle loop with the ecology coupling to vectors recorded from the reference
(tests/golden/eco_golden.npz; generator: tests/golden/make_golden.py eco)."""
import ast

import numpy as np
import pytest

from conftest import relerr
from oracle import ecology, model
from qingdai_b200.params import QDParams

ATM = {"u": "u", "v": "v", "h": "h", "T_s": "T_s", "q": "q", "cloud_cover": "cloud", "h_ice": "h_ice"}


@pytest.fixture(scope="module")
def E(golden):
    return golden("eco_golden.npz")


def oracle_constants(env):
    """Host constants of EcologyAdapter.__init__ (adapter.py:55-112) from the QD_ECO_* environment."""
    nb = int(env.get("QD_ECO_SPECTRAL_BANDS", "16"))
    _, centers, _ = ecology.make_bands(nb)
    w_b = ecology.band_weights(centers, env.get("QD_ECO_TOA_TO_SURF_MODE", "simple"))
    R_leaf = ecology.default_leaf_reflectance(centers)
    return centers, w_b, R_leaf, float(np.sum(R_leaf * w_b))


def parse_peaks(s):
    return tuple(tuple(float(x) for x in p.strip().split(":")) for p in s.split(","))


def make_state(E, tag):
    env = ast.literal_eval(str(E[f"{tag}_env"]))
    centers, w_b, R_leaf, leaf_s = oracle_constants(env)
    st = ecology.EcoState(E[f"{tag}_land"], E[f"{tag}_lai0"], leaf_s,
                          k_canopy=float(env.get("QD_ECO_LAI_K", "0.5")),
                          update_every_hours=float(env.get("QD_ECO_LIGHT_UPDATE_EVERY_HOURS", "6")),
                          lai_delta=float(env.get("QD_ECO_LIGHT_RECOMPUTE_LAI_DELTA", "0.05")),
                          soil_ref=float(env.get("QD_ECO_SOIL_REFLECT", "0.20")),
                          substep_every=int(env.get("QD_ECO_SUBSTEP_EVERY_NPHYS", "1")))
    return env, centers, w_b, R_leaf, leaf_s, st


@pytest.mark.parametrize("tag", ["u1", "u2"])
def test_adapter_constants(E, tag):
    env, centers, w_b, R_leaf, leaf_s, _ = make_state(E, tag)
    assert np.array_equal(w_b, E[f"{tag}_w_b"])
    assert np.array_equal(R_leaf, E[f"{tag}_R_leaf"])
    assert leaf_s == float(E[f"{tag}_alpha_leaf_scalar"])
    ns = E[f"{tag}_R_species"].shape[0]
    R = np.stack([np.clip(1.0 - ecology.gene_absorbance(
        centers, parse_peaks(env[f"QD_ECO_SPECIES_{i}_PEAKS"]) if f"QD_ECO_SPECIES_{i}_PEAKS" in env else ecology.DEFAULT_PEAKS), 0.0, 1.0)
        for i in range(ns)])
    assert np.array_equal(R, E[f"{tag}_R_species"])


@pytest.mark.parametrize("tag", ["u1", "u2"])
def test_subdaily_sequence(E, tag):
    """E_day, the canopy cache (time- and LAI-change-triggered recomputes), clocks and alpha: bit-exact."""
    env, centers, w_b, R_leaf, leaf_s, st = make_state(E, tag)
    dt = float(E[f"{tag}_dt"])
    for n in range(int(E[f"{tag}_ncalls"])):
        st.lai_layers = E[f"{tag}_c{n}_lai"].copy()
        with np.errstate(over="ignore"):
            a = st.step_subdaily(E[f"{tag}_c{n}_isr"], dt)
        assert (a is None) == bool(E[f"{tag}_c{n}_alpha_is_none"]), n
        if a is not None:
            assert np.array_equal(a, E[f"{tag}_c{n}_alpha"], equal_nan=True), n
        assert np.array_equal(st.E_day, E[f"{tag}_c{n}_E_day"]), n
        assert np.array_equal(st.f_cached, E[f"{tag}_c{n}_f"]), n
        assert np.array_equal(st.snapshot, E[f"{tag}_c{n}_snap"]), n
        assert np.array_equal(np.array([st.hours, st.next_hours]), E[f"{tag}_c{n}_clock"]), n
    A = st.surface_albedo_bands(E[f"{tag}_R_species"], E[f"{tag}_species_weights"])
    assert np.array_equal(A, E[f"{tag}_bands_A"], equal_nan=True)
    assert np.array_equal(w_b, E[f"{tag}_bands_w"])


def test_loop_with_ecology_vs_main(E):
    """Free-running oracle loop + ecology coupling tracks the unmodified main() (QD_ECO_ENABLE=1)."""
    nlat, nlon = int(E["loop_nlat"]), int(E["loop_nlon"])
    g = model.make_grid(nlat, nlon)
    p = QDParams.from_env({})
    dt = float(E["loop_dt"])
    land = E["loop_land_mask"]
    st = model.new_atmos_state(g, p, land, E["loop_friction"], base_albedo=E["loop_base_albedo"])
    oc = model.new_ocean_state(g, land)
    oc.Ts = np.where(land == 0, st.T_s, 288.0)
    eco = ecology.EcoState(land, E["loop_lai"], float(E["loop_alpha_leaf_scalar"]), update_every_hours=0.25)
    ia, ib = model.insolation(g, 0.0)
    eco.step_subdaily(ia + ib, dt)                               # bootstrap call, run_simulation.py:1716-1723
    for i in range(int(E["loop_nsteps"])):
        ia, ib = model.insolation(g, i * dt)
        alpha = eco.step_subdaily(ia + ib, dt)
        out = model.loop_step(st, oc, g, p, t=i * dt, dt=dt, eco_alpha=alpha)
        X = lambda k: E[f"loop_s{i}_{k}"]
        assert np.array_equal(eco.E_day, X("E_day")), i
        assert np.array_equal(eco.f_cached, X("f_canopy")), i
        assert np.array_equal(np.array([eco.hours, eco.next_hours]), X("eco_clock")), i
        assert relerr(out.albedo, X("albedo")) < 1e-9, i
        for k, mine in ATM.items():
            assert relerr(getattr(st, mine), X(k)) < 1e-9, (i, k)
        for k in ("uo", "vo", "eta", "Ts"):
            assert relerr(getattr(oc, k), X(k)) < 1e-9, (i, k)
