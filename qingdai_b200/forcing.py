"""Orbital mechanics and thermal forcing: host side.

Mirrors the reference interfaces ``pygcm.orbital.OrbitalSystem`` (orbital.py:10-77) and
``pygcm.forcing.ThermalForcing`` (forcing.py:11-165).  Only ~10 scalars per step depend on time
(star fluxes, declinations, right ascensions, rotation angle); they are evaluated here with NumPy
exactly as the reference does and shipped to the device as a ``qd_forcing_t``; the per-cell
insolation field itself is computed on the GPU inside the fused column kernel.
"""
from __future__ import annotations

import numpy as np

from . import constants as const


class OrbitalSystem:
    """Two-star + planet circular orbits (orbital.py:15-77)."""

    def __init__(self):
        self.T_binary = 2 * np.pi * np.sqrt(const.A_BINARY ** 3 / (const.G * const.M_TOTAL_STARS))
        self.T_planet = 2 * np.pi * np.sqrt(const.A_PLANET ** 3 / (const.G * const.M_TOTAL_STARS))
        self.omega_binary = 2 * np.pi / self.T_binary
        self.omega_planet = 2 * np.pi / self.T_planet
        self.r_A = const.A_BINARY * (const.M_B / const.M_TOTAL_STARS)
        self.r_B = const.A_BINARY * (const.M_A / const.M_TOTAL_STARS)

    def calculate_stellar_positions(self, t):
        ang = self.omega_binary * t
        return (self.r_A * np.cos(ang), self.r_A * np.sin(ang), -self.r_B * np.cos(ang), -self.r_B * np.sin(ang))

    def calculate_total_flux(self, t):
        x_A, y_A, x_B, y_B = self.calculate_stellar_positions(t)
        x_p = const.A_PLANET * np.cos(self.omega_planet * t)
        y_p = const.A_PLANET * np.sin(self.omega_planet * t)
        d_A = np.sqrt((x_p - x_A) ** 2 + (y_p - y_A) ** 2)
        d_B = np.sqrt((x_p - x_B) ** 2 + (y_p - y_B) ** 2)
        return const.L_A / (4 * np.pi * d_A ** 2) + const.L_B / (4 * np.pi * d_B ** 2)


class ThermalForcing:
    """Dual-star insolation and radiative-equilibrium temperature (forcing.py:16-165)."""

    def __init__(self, grid, orbital_system: OrbitalSystem):
        self.grid = grid
        self.orbital_system = orbital_system
        self.planet_params = {"axial_tilt": const.PLANET_AXIAL_TILT, "omega": const.PLANET_OMEGA,
                              "T_planet": orbital_system.T_planet}
        tilt = np.deg2rad(const.PLANET_AXIAL_TILT)
        self.n_hat = np.array([np.sin(tilt), 0.0, np.cos(tilt)])
        x_in = np.array([1.0, 0.0, 0.0])
        self.x_eq = x_in - np.dot(x_in, self.n_hat) * self.n_hat
        self.x_eq /= np.linalg.norm(self.x_eq)
        self.y_eq = np.cross(self.n_hat, self.x_eq)

    # ---- scalars for the device -------------------------------------------------------------
    def star_geometry(self, t):
        """[(flux, sin_delta, cos_delta, alpha)_A, (...)_B], theta for time t (forcing.py:78-131)."""
        o = self.orbital_system
        ang = o.omega_planet * t
        x_A, y_A, x_B, y_B = o.calculate_stellar_positions(t)
        x_p = const.A_PLANET * np.cos(ang)
        y_p = const.A_PLANET * np.sin(ang)
        out = []
        for xs, ys, L in ((x_A, y_A, const.L_A), (x_B, y_B, const.L_B)):
            vec = np.array([xs - x_p, ys - y_p, 0.0])
            dist = np.linalg.norm(vec)
            flux = L / (4 * np.pi * (dist ** 2))
            s_hat = vec / (np.linalg.norm(vec) + 1e-15)
            delta = np.arcsin(np.clip(np.dot(s_hat, self.n_hat), -1.0, 1.0))
            alpha = np.arctan2(np.dot(s_hat, self.y_eq), np.dot(s_hat, self.x_eq))
            out.append((float(flux), float(np.sin(delta)), float(np.cos(delta)), float(alpha)))
        theta = (t * self.planet_params["omega"]) % (2 * np.pi)
        return out, float(theta)

    # ---- NumPy-in / NumPy-out reference API (host evaluation; small and off the hot path) ------
    def calculate_insolation_components(self, t):
        stars, theta = self.star_geometry(t)
        lon = np.deg2rad(self.grid.lon_mesh)
        lat = np.deg2rad(self.grid.lat_mesh)
        res = []
        for flux, sd, cd, alpha in stars:
            cz = np.sin(lat) * sd + np.cos(lat) * cd * np.cos(theta + lon - alpha)
            res.append(flux * np.maximum(0.0, cz))
        return res[0], res[1]

    def calculate_insolation(self, t):
        a, b = self.calculate_insolation_components(t)
        return a + b

    def calculate_equilibrium_temp(self, t, albedo):
        num = self.calculate_insolation(t) * (1 - albedo)
        num[num < 0] = 0
        return (num / const.SIGMA) ** 0.25
