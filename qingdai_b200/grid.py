"""Regular lat-lon grid with duplicated seam meridian and both poles (pygcm/grid.py:10-39).

``divergence`` / ``vorticity`` keep the reference's NumPy-in / NumPy-out signatures
(grid.py:41-88) but run on the GPU through the engine attached to the grid.
"""
from __future__ import annotations

import numpy as np

from . import constants


class SphericalGrid:
    def __init__(self, n_lat, n_lon):
        self.n_lat = int(n_lat)
        self.n_lon = int(n_lon)
        self.lat = np.linspace(-90, 90, self.n_lat)
        self.lon = np.linspace(0, 360, self.n_lon)
        self.lon_mesh, self.lat_mesh = np.meshgrid(self.lon, self.lat)
        self.coriolis_param = 2 * constants.PLANET_OMEGA * np.sin(np.deg2rad(self.lat_mesh))
        self.dlat_rad = np.deg2rad(self.lat[1] - self.lat[0])
        self.dlon_rad = np.deg2rad(self.lon[1] - self.lon[0])

    def _engine(self):
        from .engine import engine_for_grid
        return engine_for_grid(self)

    def divergence(self, u, v):
        return self._engine().op_divvort(u, v, vort=False)

    def vorticity(self, u, v):
        return self._engine().op_divvort(u, v, vort=True)
