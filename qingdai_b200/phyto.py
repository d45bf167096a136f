"""Phytoplankton tracer transport behind the reference's interface (pygcm/ecology/phyto.py:96-126,453-547).

``PhytoTransport`` carries the per-physics-step part of ``PhytoManager``: ``C_phyto_s[S, lat, lon]`` lives in HBM and
``advect_diffuse(uo, vo, dt_seconds)`` runs three batched kernels (csrc/qd_phyto.cuh) for all species.  The daily
growth / optics step of the reference (``step_daily``) stays host Python and reads / assigns ``C_phyto_s``."""
from __future__ import annotations

import os

import numpy as np
import torch

from .engine import Engine, _ptr


class PhytoTransport:
    def __init__(self, grid, land_mask, engine: Engine | None = None, n_species=10, env=None, lib=None, device=None):
        env = os.environ if env is None else env
        self.grid = grid
        self.land_mask = np.asarray(land_mask).astype(int)
        nlat, nlon = self.land_mask.shape
        if engine is None:
            engine = Engine(nlat, nlon, batch=1, device=device, lib=lib)
            engine.set_mask("land", self.land_mask.astype(np.uint8))
        self.engine = engine
        self.K_h = float(env.get("QD_PHYTO_KH", env.get("QD_KH_OCEAN", "5.0e3")))       # phyto.py:122-125
        self.adv_alpha = float(env.get("QD_PHYTO_ADV_ALPHA", "0.7"))                    # phyto.py:517
        self.S = int(n_species)
        self._C = torch.zeros((self.S, nlat, nlon), dtype=torch.float64, device=engine.device)
        self._uv = torch.zeros((2, nlat, nlon), dtype=torch.float64, device=engine.device)

    @property
    def C_phyto_s(self):
        self.engine.sync()
        return self._C.cpu().numpy()

    @C_phyto_s.setter
    def C_phyto_s(self, value):
        v = np.array(value, dtype=np.float64, order="C")
        if v.shape != tuple(self._C.shape):
            self.S = v.shape[0]
            self._C = torch.zeros(v.shape, dtype=torch.float64, device=self.engine.device)
        self._C.copy_(torch.from_numpy(v))

    def advect_diffuse(self, uo=None, vo=None, dt_seconds=0.0):
        """``uo``/``vo``: host arrays like the reference passes (``ocean.uo``, ``ocean.vo``), or None to use the
        engine's own ocean currents without leaving the device."""
        e = self.engine
        pu = pv = None
        if uo is not None:
            self._uv[0].copy_(torch.from_numpy(np.array(uo, dtype=np.float64, order="C")))
            self._uv[1].copy_(torch.from_numpy(np.array(vo, dtype=np.float64, order="C")))
            pu, pv = _ptr(self._uv[0]), _ptr(self._uv[1])
        e._chk(e.lib.qd_phyto_advect_diffuse(e.ctx, _ptr(self._C), self.S, pu, pv, float(dt_seconds), self.adv_alpha, self.K_h),
               "qd_phyto_advect_diffuse")
