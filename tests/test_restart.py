"""Restart / topography / ocean files without netCDF4: the NetCDF-3 shim, parity of
qingdai_b200.restart with what the reference's own helpers returned for the same inputs (recorded by
tests/golden/make_golden.py restart, which runs them unmodified on the shim), and bit-exact checkpoint resume."""
import os
import sys
import types

import numpy as np
import pytest

from qingdai_b200 import ncio, restart
from qingdai_b200.grid import SphericalGrid


@pytest.fixture(scope="module")
def R(golden):
    return golden("restart_golden.npz")


def test_shim_roundtrip_types_and_attributes(tmp_path):
    p = str(tmp_path / "a.nc")
    with ncio.Dataset(p, "w") as ds:
        ds.createDimension("y", 3)
        ds.createDimension("x", 4)
        f = ds.createVariable("f", "f4", ("y", "x")); f[:] = np.arange(12.0).reshape(3, 4) + 0.1
        d = ds.createVariable("d", "f8", ("y", "x")); d[:] = np.pi * np.arange(12.0).reshape(3, 4)
        m = ds.createVariable("mask", "u1", ("y", "x")); m[:] = (np.arange(12).reshape(3, 4) % 3 == 0) * 200
        i = ds.createVariable("idx", "i8", ("x",)); i[:] = np.array([-1, 0, 7, 2 ** 31 - 1])
        s = ds.createVariable("t", "f8"); s[...] = 12.5
        d.units = "m"
        d.setncattr("long_name", "depth")
        ds.setncattr("title", "t")
        ds.setncattr("day", 3.25)
        ds.setncattr("n", 7)
    with ncio.Dataset(p, "r") as ds:
        assert ds.dimensions["y"].size == 3 and len(ds.dimensions["x"]) == 4
        assert set(ds.variables) == {"f", "d", "mask", "idx", "t"} and "mask" in ds.variables
        assert ds.variables["f"][:].data.dtype == np.float32
        assert np.array_equal(ds["f"][:], (np.arange(12.0).reshape(3, 4) + 0.1).astype(np.float32))
        assert np.array_equal(ds["d"][:].data, np.pi * np.arange(12.0).reshape(3, 4))        # float64 survives exactly
        assert ds["mask"][:].dtype == np.uint8 and ds["mask"][:].max() == 200               # unsigned restored
        assert ds["idx"][:].dtype == np.int64 and list(ds["idx"][:]) == [-1, 0, 7, 2 ** 31 - 1]
        assert float(ds.variables["t"][...]) == 12.5
        assert ds["d"].units == "m" and ds["d"].getncattr("long_name") == "depth" and ds["d"].dimensions == ("y", "x")
        assert ds.getncattr("title") == "t" and ds.getncattr("day") == 3.25 and ds.getncattr("n") == 7
        assert np.array_equal(ds["d"][1:, ::2], (np.pi * np.arange(12.0).reshape(3, 4))[1:, ::2])
        with pytest.raises(AttributeError):
            ds.getncattr("missing")


def test_shim_refuses_what_netcdf3_cannot_hold(tmp_path):
    p = str(tmp_path / "b.nc")
    with ncio.Dataset(p, "w") as ds:
        ds.createDimension("x", 2)
        v = ds.createVariable("big", "i8", ("x",))
        with pytest.raises(OverflowError):
            v[:] = np.array([0, 2 ** 40])
    h5 = tmp_path / "c.nc"
    h5.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    with pytest.raises(OSError, match="NetCDF-4/HDF5"):
        ncio.Dataset(str(h5), "r")


def test_install_as_netcdf4(tmp_path, monkeypatch):
    monkeypatch.delitem(sys.modules, "netCDF4", raising=False)
    try:
        import netCDF4  # noqa: F401
        pytest.skip("a real netCDF4 is installed")
    except ImportError:
        pass
    assert ncio.install_netcdf4_shim()
    from netCDF4 import Dataset
    assert Dataset is ncio.Dataset
    monkeypatch.delitem(sys.modules, "netCDF4", raising=False)


def _objects(R):
    gcm, oc = types.SimpleNamespace(), types.SimpleNamespace()
    for k in ("u", "v", "h", "T_s", "cloud_cover", "q", "h_ice"):
        setattr(gcm, k, R["in_gcm_" + k])
    for k in ("uo", "vo", "eta", "Ts"):
        setattr(oc, k, R["in_oc_" + k])
    return gcm, oc


def test_restart_matches_reference_helpers(R, tmp_path):
    g = SphericalGrid(13, 24)
    gcm, oc = _objects(R)
    p = str(tmp_path / "data" / "restart.nc")                 # the directory is created, as in the reference
    restart.save_restart(p, g, gcm, oc, R["in_land"], W_land=R["in_W"], S_snow=R["in_S"], C_snow=None, t_seconds=123456.789)
    out = restart.load_restart(p)
    none_keys = set(str(k) for k in R["restart_none"])
    assert {k for k, v in out.items() if v is None} == none_keys == {"C_snow"}
    for k, v in out.items():
        if v is None:
            continue
        want = R["restart_" + k]
        assert np.asarray(v).dtype == want.dtype and np.array_equal(np.asarray(v), want), k
    assert out["t_seconds"] == 123456.789 and out["u"].dtype == np.float32
    # float64 option: exact
    restart.save_restart(p, g, gcm, oc, R["in_land"], dtype="f8")
    out = restart.load_restart(p)
    assert np.array_equal(out["T_s"], R["in_gcm_T_s"]) and np.array_equal(out["eta"], R["in_oc_eta"])
    assert out["W_land"] is None and out["t_seconds"] == 0.0
    # no ocean object: its variables are simply absent
    restart.save_restart(p, g, gcm, None, R["in_land"])
    assert restart.load_restart(p)["uo"] is None


def test_ocean_file_matches_reference_helpers(R, tmp_path):
    g = SphericalGrid(13, 24)
    _, oc = _objects(R)
    p = str(tmp_path / "ocean.nc")
    assert restart.save_ocean(p, g, oc, day_value=12.5) is True
    out = restart.load_ocean(p)
    for k in ("uo", "vo", "eta", "Ts"):
        assert np.array_equal(out[k], R["ocean_" + k]), k
    assert out["day"] == float(R["ocean_day"]) == 12.5
    missing = restart.load_ocean(str(tmp_path / "nope.nc"))          # reference behaviour: report, return Nones
    assert all(v is None for v in missing.values())


@pytest.mark.parametrize("tag,shape", [("same", (13, 24)), ("fine", (19, 40))])
def test_topography_load_and_regrid_match_reference(R, tmp_path, tag, shape):
    src = SphericalGrid(13, 24)
    p = str(tmp_path / "topography.nc")
    restart.save_topography(p, src, R["in_land"], R["in_alb"], R["in_fric"], elevation=R["in_elev"])
    e, m, a, f = restart.load_topography_from_netcdf(p, SphericalGrid(*shape))
    assert m.dtype == np.uint8 and e.shape == shape
    for got, key in ((e, "elev"), (m, "mask"), (a, "alb"), (f, "fric")):
        assert np.array_equal(got, R[f"topo_{tag}_{key}"]), (tag, key)
    if tag == "fine":
        with pytest.raises(ValueError, match="regrid='never'"):
            restart.load_topography_from_netcdf(p, SphericalGrid(*shape), regrid="never")


def test_network_file_roundtrip_feeds_the_routing_loader(tmp_path):
    """hydrology_network.save_network writes generate_hydrology_maps.py's layout; routing.load_network reads it back
    (NetCDF-3 reader when netCDF4 is absent) with the arrays the dict path would have handed to RiverRouting."""
    from qingdai_b200.hydrology_network import save_network
    from qingdai_b200.routing import load_network
    g = SphericalGrid(9, 16)
    rng = np.random.default_rng(1)
    land = (rng.uniform(size=(9, 16)) < 0.5).astype(np.uint8)
    n_land = int(land.sum())
    net = {"land_mask": land, "elevation_filled": rng.uniform(0, 100, (9, 16)),
           "flow_to_index": rng.integers(-1, 144, (9, 16)).astype(np.int64), "flow_order": rng.permutation(144)[:n_land].astype(np.int64),
           "lake_mask": (rng.uniform(size=(9, 16)) < 0.1).astype(np.uint8), "lake_id": rng.integers(0, 3, (9, 16)).astype(np.int32),
           "n_lakes": 2, "lake_outlet_index": np.array([-1, 17], dtype=np.int32)}
    p = str(tmp_path / "data" / "hydrology_network.nc")
    save_network(p, g, net)
    back = load_network(p)
    for k in ("land_mask", "flow_to_index", "flow_order", "lake_mask", "lake_id", "lake_outlet_index"):
        assert np.array_equal(back[k], net[k]), k
    assert back["land_mask"].dtype == np.uint8


def test_topography_foreign_layout_matches_reference(R, tmp_path):
    """Latitude descending, longitudes in [-180, 180) without a seam column, coarser than the target grid: the loader
    flips, renormalises, sorts and regrids exactly like topography.load_topography_from_netcdf."""
    p = str(tmp_path / "foreign.nc")
    with ncio.Dataset(p, "w") as ds:
        ds.createDimension("lat", 10)
        ds.createDimension("lon", 18)
        v = ds.createVariable("lat", "f8", ("lat",)); v[:] = R["foreign_lat"]
        v = ds.createVariable("lon", "f8", ("lon",)); v[:] = R["foreign_lon"]
        for name, dt_, key in (("elevation", "f8", "elev"), ("land_mask", "u1", "mask"), ("base_albedo", "f8", "alb"), ("friction", "f8", "fric")):
            v = ds.createVariable(name, dt_, ("lat", "lon")); v[:] = R["foreign_" + key]
    e, m, a, f = restart.load_topography_from_netcdf(p, SphericalGrid(13, 24))
    for got, key in ((e, "elev"), (m, "mask"), (a, "alb"), (f, "fric")):
        assert np.array_equal(got, R["topo_foreign_" + key]), key
