"""Pin oracle.phyto.advect_diffuse to call sequences recorded from the reference's PhytoManager
(tests/golden/phyto_golden.npz; generator: tests/golden/make_golden.py phyto)."""
import numpy as np
import pytest

from conftest import relerr
from oracle import model, phyto


@pytest.mark.parametrize("tag", ["p1", "p2"])
def test_advect_diffuse_vs_reference(golden, tag):
    G = golden("phyto_golden.npz")
    land = G[f"{tag}_land"]
    g = model.make_grid(*land.shape)
    C = G[f"{tag}_C0"]
    for n in range(int(G[f"{tag}_ncalls"])):
        with np.errstate(all="ignore"):
            C = phyto.advect_diffuse(C, G[f"{tag}_c{n}_uo"], G[f"{tag}_c{n}_vo"], land, float(G[f"{tag}_dt"]), g.a, g.dlat, g.dlon,
                                     g.lat_rad, adv_alpha=float(G[f"{tag}_alpha"]), K_h=float(G[f"{tag}_kh"]))
        assert relerr(C, G[f"{tag}_c{n}_C"]) < 1e-14, n
