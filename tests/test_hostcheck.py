"""CPU-only CI: run the parity checks of tests/qdcheck.py against the host-compiled check build of the
kernel sources (tests/hostcheck).  This validates kernel LOGIC (indexing, operand order, boundary
semantics, step orchestration through the C ABI); the same checks run on the B200 in test_gpu.py."""
import numpy as np
import pytest

import qdcheck


@pytest.fixture(scope="module")
def lib():
    from hostcheck import library
    return library()


@pytest.fixture(scope="module")
def G(golden):
    return golden("ops_golden.npz")


@pytest.fixture(scope="module")
def C(golden):
    return golden("cores_golden.npz")


@pytest.fixture(scope="module")
def L(golden):
    return golden("loop_golden.npz")


@pytest.mark.parametrize("tag,shape", [("a", (22, 40)), ("b", (15, 27))])
def test_ops_vs_golden(lib, G, tag, shape):
    qdcheck.check_ops_vs_golden(lib, G, tag, shape)


def test_median_edge_cases(lib):
    qdcheck.check_median_edge_cases(lib)


def test_median_large_duplicates_and_batch(lib):
    qdcheck.check_median_large(lib)


def test_ops_random(lib):
    qdcheck.check_ops_random(lib)


@pytest.mark.parametrize("tag", list(qdcheck.CASES))
def test_atmos_step(lib, C, tag):
    qdcheck.check_atmos_step(lib, C, tag)


@pytest.mark.parametrize("tag", list(qdcheck.CASES))
def test_ocean_step(lib, C, tag):
    qdcheck.check_ocean_step(lib, C, tag)


def test_ocean_storm(lib, C):
    qdcheck.check_ocean_storm(lib, C)


@pytest.mark.parametrize("tag", ["base", "banded"])
def test_loop_teacher_forced(lib, L, tag):
    qdcheck.check_loop_teacher_forced(lib, L, tag)


@pytest.mark.parametrize("tag", ["base", "banded"])
def test_loop_free_running(lib, L, tag):
    qdcheck.check_loop_free_running(lib, L, tag)


def test_loop_energy_branch(lib):
    qdcheck.check_loop_energy_branch(lib)


def test_dropin_classes(lib, C):
    qdcheck.check_dropin_classes(lib, C)


def test_jax_compat_seam(lib, G):
    qdcheck.check_jax_compat_seam(lib, G)


@pytest.mark.parametrize("tag", ["r1", "r2"])
def test_routing(lib, golden, tag):
    qdcheck.check_routing(lib, golden("routing_golden.npz"), tag)


def test_graph_levels_are_noops_in_host_build(lib, L):
    qdcheck.check_graph_levels_agree(lib, L, nsteps=7)


@pytest.mark.parametrize("tag", ["u1", "u2"])
def test_ecology_adapter(lib, golden, tag):
    qdcheck.check_eco_unit(lib, golden("eco_golden.npz"), tag)


def test_loop_with_ecology(lib, golden):
    qdcheck.check_eco_loop(lib, golden("eco_golden.npz"))


def test_hyper4_streaming_kernel(lib):
    qdcheck.check_hyper4_stream(lib)


@pytest.mark.parametrize("tag", ["p1", "p2"])
def test_phyto_transport(lib, golden, tag):
    qdcheck.check_phyto(lib, golden("phyto_golden.npz"), tag)


def test_global_diagnostics(lib):
    qdcheck.check_diag(lib)


def test_config3_routing_and_ecology_in_one_loop(lib, golden):
    qdcheck.check_config3(lib, golden("routing_golden.npz"))


def test_multiday_global_diagnostics(lib):
    qdcheck.check_multiday(lib)


def test_bandstop_long_rows(lib):
    qdcheck.check_bandstop_large(lib)


@pytest.mark.parametrize("tag", ["i1", "i2"])
def test_individual_pool_substeps(lib, golden, tag):
    qdcheck.check_indiv(lib, golden("indiv_golden.npz"), tag)


def test_gaussian_fused_tile_kernel(lib):
    qdcheck.check_gauss2d_large(lib)


def test_tiny_and_ragged_grids(lib):
    qdcheck.check_tiny_grids(lib)


def test_checkpoint_resume_is_bit_exact(lib):
    qdcheck.check_checkpoint_resume(lib)


def test_checkpoint_resume_with_routing_and_ecology(lib, golden):
    qdcheck.check_checkpoint_resume_config3(lib, golden("routing_golden.npz"))


def test_reference_format_restart(lib):
    qdcheck.check_reference_format_restart(lib)


# ---------------------------------------------------------------------------- round 2 checks at CPU-affordable sizes
def test_full_loop_teacher_forced_from_device_state(lib):
    qdcheck.check_loop_step_vs_oracle(lib, (31, 60), 600.0, spin=3, nsteps=2, cold=True)
    qdcheck.check_loop_step_vs_oracle(lib, (31, 60), 300.0, spin=3, nsteps=2, with_albedo=False, p=qdcheck.QDParams())


def test_config3_teacher_forced(lib):
    qdcheck.check_config3_teacher_forced(lib, (31, 60), nsteps=2, spin=3, dt=600.0, dt_hydro_hours=0.5)


def test_batch_members_equal_standalone_runs_bitwise(lib):
    qdcheck.check_batch_equivalence(lib, (19, 36), B=3, nsteps=3)


def test_batch_with_mismatched_switches_raises(lib):
    qdcheck.check_switch_mismatch_raises(lib)


def test_long_call_and_parameter_change(lib):
    qdcheck.check_long_call_and_param_change(lib)


def test_env_driven_entry_point(lib, tmp_path, golden, monkeypatch):
    """qingdai_b200.run_simulation.main: QD_N_LAT / QD_N_LON / QD_SIM_DAYS / QD_TOPO_NC / QD_RESTART_OUT -> the same state as
    driving Simulation directly; QD_RESTART_IN resumes from the file it wrote."""
    import os
    from qingdai_b200 import _binding, run_simulation
    from qingdai_b200.params import QDParams
    from qingdai_b200.simulation import Simulation, DAY_SECONDS
    from qingdai_b200.synthetic import load_reference_topography
    monkeypatch.setattr(_binding, "_default", lib)
    topo_nc = os.path.join(qdcheck.GOLDEN_DIR, "topography_qingdai_181x360_seed42.nc")
    nlat, nlon, dt, nsteps = 19, 36, 600, 6
    r_out = str(tmp_path / "restart.nc")
    env = {"QD_N_LAT": str(nlat), "QD_N_LON": str(nlon), "QD_DT_SECONDS": str(dt), "QD_SIM_DAYS": repr((nsteps - 0.5) * dt / DAY_SECONDS),
           "QD_TOPO_NC": topo_nc, "QD_OROG": "1", "QD_ENERGY_W": "1", "QD_LOOP_WITH_ALBEDO": "1", "QD_RESTART_OUT": r_out, "QD_INIT_BANDED": "1"}
    logs = []
    sim = run_simulation.main(env, log=logs.append)
    assert sim.step_index == nsteps and any("<Ts>" in s for s in logs) and os.path.exists(r_out)
    ref = Simulation(nlat, nlon, load_reference_topography(topo_nc, nlat, nlon), QDParams.from_env(env), dt=dt, lib=lib, loop_with_albedo=True)
    ts0 = 265.0 + (295.0 - 265.0) * (np.cos(np.deg2rad(ref.grid.lat_mesh)) ** 2)
    ref.engine.set("ts", ts0)
    ref.engine.set("sst", np.where(ref.engine.get_mask("land") == 0, ts0, ref.engine.get("sst")))
    ref.step(nsteps)
    for k in ("u", "v", "h", "ts", "q", "cloud", "uo", "vo", "eta", "sst", "wland"):
        assert np.array_equal(sim.engine.get(k), ref.engine.get(k)), k
    env2 = dict(env, QD_RESTART_IN=r_out, QD_SIM_DAYS=repr(1.5 * dt / DAY_SECONDS))
    env2.pop("QD_RESTART_OUT"); env2.pop("QD_INIT_BANDED")
    sim2 = run_simulation.main(env2, log=logs.append)
    assert sim2.t == (nsteps + 2) * dt and np.all(np.isfinite(sim2.engine.get("ts")))
