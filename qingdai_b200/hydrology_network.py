"""D8 routing-network builder behind the reference's function names (scripts/generate_hydrology_maps.py:65-273).

``build_network(grid, elevation, land_mask)`` returns the arrays the reference writes to ``data/hydrology_network.nc``
(and ``qingdai_b200.routing.RiverRouting`` accepts as a dict): ``land_mask, flow_to_index, flow_order, lake_mask,
lake_id, lake_outlet_index`` plus ``elevation_filled``.  The order-dependent sequential algorithms run in host C++
inside libqd_b200 (csrc/qd_netbuild.h) with bit-identical results; the Python loops of the reference need seconds at
61x120 and far longer at the benchmark grids."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import constants
from ._binding import default_library


def neighbour_distances(grid):
    """dist[cls, j, dj+1, di+1] = spherical_distance(grid, i, j, (i+di) % n_lon, j+dj) (:65-82) for the three column
    classes cls = 0: i = 0, 1: interior, 2: i = n_lon-1.  The grid carries both 0 and 360 degrees, so across the seam
    the reference's wrapped longitude difference is 0, not dlon."""
    R = float(constants.PLANET_RADIUS)
    lat = np.asarray(grid.lat, dtype=float)
    lon = np.asarray(grid.lon, dtype=float)
    nlat, nlon = lat.size, lon.size
    out = np.zeros((3, nlat, 3, 3))
    for cls, i in enumerate((0, 1, nlon - 1)):
        for j in range(nlat):
            for dj in (-1, 0, 1):
                jj = j + dj
                if jj < 0 or jj >= nlat:
                    continue
                for di in (-1, 0, 1):
                    if di == 0 and dj == 0:
                        continue
                    ii = (i + di) % nlon
                    lat1, lon1 = np.deg2rad(lat[j]), np.deg2rad(lon[i])
                    lat2, lon2 = np.deg2rad(lat[jj]), np.deg2rad(lon[ii])
                    dlat, dlon = lat2 - lat1, lon2 - lon1
                    if dlon > np.pi:
                        dlon -= 2 * np.pi
                    elif dlon < -np.pi:
                        dlon += 2 * np.pi
                    x = dlon * np.cos(0.5 * (lat1 + lat2))
                    out[cls, j, dj + 1, di + 1] = R * np.sqrt(x * x + dlat * dlat)
    return out


def build_network(grid, elevation, land_mask, pit_iters=200, pit_eps=1e-3, lib=None):
    lib = lib or default_library()
    land = np.ascontiguousarray((np.asarray(land_mask) == 1).astype(np.uint8))
    nlat, nlon = land.shape
    elev = np.array(elevation, dtype=np.float64, order="C")
    dist = np.ascontiguousarray(neighbour_distances(grid))
    n = nlat * nlon
    flow_to = np.empty(n, dtype=np.int64)
    order = np.empty(n, dtype=np.int64)
    lake_mask = np.empty(n, dtype=np.uint8)
    lake_id = np.empty(n, dtype=np.int32)
    outlet = np.full(n, -1, dtype=np.int32)
    n_order, n_lakes, sweeps = C.c_int64(0), C.c_int(0), C.c_int(0)
    P = lambda a: C.c_void_p(a.ctypes.data)
    rc = lib.qd_net_build(nlat, nlon, P(elev), P(land), P(dist), int(pit_iters), float(pit_eps), P(flow_to), P(order),
                          C.cast(C.byref(n_order), C.c_void_p), P(lake_mask), P(lake_id), P(outlet), C.byref(n_lakes), C.byref(sweeps))
    if rc != 0:
        raise RuntimeError(f"qd_net_build failed with status {rc}")
    net = {"land_mask": land, "elevation_filled": elev, "flow_to_index": flow_to.reshape(nlat, nlon),
           "flow_order": order[: n_order.value].copy(), "lake_mask": lake_mask.reshape(nlat, nlon),
           "lake_id": lake_id.reshape(nlat, nlon), "n_lakes": int(n_lakes.value), "pit_sweeps": int(sweeps.value)}
    if n_lakes.value > 0:
        net["lake_outlet_index"] = outlet[: n_lakes.value].copy()
    return net


def save_network(path, grid, net):
    """Write ``build_network``'s result with the layout of scripts/generate_hydrology_maps.py:329-362 (variables,
    types, dimensions and attributes), as NetCDF-3 through qingdai_b200.ncio: ``pygcm.routing.RiverRouting`` and
    ``qingdai_b200.routing.RiverRouting`` both read it, with or without netCDF4 installed."""
    import os
    from .ncio import Dataset
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    land = np.asarray(net["land_mask"]).astype(np.uint8)
    with Dataset(path, "w") as ds:
        ds.createDimension("lat", grid.n_lat)
        ds.createDimension("lon", grid.n_lon)
        ds.createDimension("n_land", int(np.asarray(net["flow_order"]).size))
        n_lakes = int(net.get("n_lakes", 0))
        if n_lakes > 0:
            ds.createDimension("n_lakes", n_lakes)
        vlat = ds.createVariable("lat", "f4", ("lat",)); vlat[:] = np.asarray(grid.lat, dtype=np.float32)
        vlon = ds.createVariable("lon", "f4", ("lon",)); vlon[:] = np.asarray(grid.lon, dtype=np.float32)
        for name, dtype, dims in (("land_mask", "u1", ("lat", "lon")), ("elevation_filled", "f4", ("lat", "lon")),
                                  ("flow_to_index", "i4", ("lat", "lon")), ("flow_order", "i4", ("n_land",)),
                                  ("lake_mask", "u1", ("lat", "lon")), ("lake_id", "i4", ("lat", "lon"))):
            v = ds.createVariable(name, dtype, dims)
            v[:] = land if name == "land_mask" else np.asarray(net[name])
        if n_lakes > 0:
            v = ds.createVariable("lake_outlet_index", "i4", ("n_lakes",))
            v[:] = np.asarray(net["lake_outlet_index"])
        ds.setncattr("title", "Qingdai Hydrology Network")
        ds.setncattr("indexing", "row-major (i=lon index, j=lat index), idx=j*n_lon+i")
        ds.setncattr("projection", "latlon")
        ds.setncattr("created_by", "qingdai_b200.hydrology_network.save_network")
