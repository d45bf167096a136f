"""Pin oracle.individuals (band split of the dual-star insolation + IndividualPool.try_substep) to vectors recorded
from the reference (tests/golden/indiv_golden.npz; generator: tests/golden/make_golden.py indiv)."""
import ast

import numpy as np
import pytest

from oracle import ecology, individuals
from qingdai_b200 import constants as const


def spectra(env):
    _, centers, widths = ecology.make_bands(int(env.get("QD_ECO_SPECTRAL_BANDS", "16")))
    TA = individuals.teff(const.L_A / individuals.L_SUN, const.M_A / individuals.M_SUN)
    TB = individuals.teff(const.L_B / individuals.L_SUN, const.M_B / individuals.M_SUN)
    if env.get("QD_ECO_TOA_TO_SURF_MODE", "simple") == "rayleigh":
        T_ray = np.clip(0.9 * (np.maximum(1e-6, centers) / 550.0) ** 4.0, 0.0, None)
    else:
        T_ray = np.ones_like(centers)
    return individuals.blackbody_band_weights(TA, centers, widths), individuals.blackbody_band_weights(TB, centers, widths), T_ray


@pytest.mark.parametrize("tag", ["i1", "i2"])
def test_try_substep_vs_reference(golden, tag):
    G = golden("indiv_golden.npz")
    env = ast.literal_eval(str(G[f"{tag}_env"]))
    spec = spectra(env)
    pool = individuals.Pool(G[f"{tag}_sample_j"], G[f"{tag}_sample_i"], G[f"{tag}_indiv_cell_index"], G[f"{tag}_indiv_Ab"],
                            G[f"{tag}_indiv_tol"], int(G[f"{tag}_cfg"][2]))
    dt, day = float(G[f"{tag}_dt"]), float(G[f"{tag}_day"])
    for n in range(int(G[f"{tag}_ncalls"])):
        pool.try_substep(G[f"{tag}_c{n}_isrA"], G[f"{tag}_c{n}_isrB"], spec, G[f"{tag}_c{n}_soil"], dt, day)
        assert np.array_equal(pool.E_day, G[f"{tag}_c{n}_E"]), n
        assert np.array_equal(pool.stress, G[f"{tag}_c{n}_stress"]), n
        assert pool.accum == float(G[f"{tag}_c{n}_accum"]), n
    n = int(G[f"{tag}_ncalls"]) - 1
    assert np.array_equal(individuals.insolation_to_bands(G[f"{tag}_c{n}_isrA"], G[f"{tag}_c{n}_isrB"], *spec), G[f"{tag}_Ib_last"])
