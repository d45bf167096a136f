"""The host C++ routing-network builder (csrc/qd_netbuild.h) against networks the REFERENCE's own builder produced
(scripts/generate_hydrology_maps.py, recorded in tests/golden/routing_golden.npz by make_golden.py routing): the
pit-filled elevation, D8 flow directions, topological order, lake labels and outlets must be bit-identical.  Host code
only -- runs in the CPU suite through the host check build of the same sources."""
import numpy as np
import pytest


@pytest.mark.parametrize("tag", ["r1", "r2"])
def test_network_builder_bit_exact(golden, tag):
    from hostcheck import library
    from qingdai_b200.grid import SphericalGrid
    from qingdai_b200.hydrology_network import build_network
    G = golden("routing_golden.npz")
    land = G[f"{tag}_land_mask"]
    net = build_network(SphericalGrid(*land.shape), G[f"{tag}_elev_in"], land, lib=library())
    assert np.array_equal(net["elevation_filled"], G[f"{tag}_elev_filled"])
    assert np.array_equal(net["flow_to_index"], G[f"{tag}_flow_to_index"])
    assert np.array_equal(net["flow_order"], G[f"{tag}_flow_order"])
    assert np.array_equal(net["lake_mask"], G[f"{tag}_lake_mask"])
    assert np.array_equal(net["lake_id"], G[f"{tag}_lake_id"])
    assert net["n_lakes"] == int(G[f"{tag}_n_lakes"])
    if net["n_lakes"] > 0:
        assert np.array_equal(net["lake_outlet_index"], G[f"{tag}_lake_outlet_index"])
