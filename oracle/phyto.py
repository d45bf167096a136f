"""oracle/phyto.py -- NumPy restatement of PhytoManager.advect_diffuse (test infrastructure only; never imported
by the product).  Reference: pygcm/ecology/phyto.py:453-547 (Laplacian :453-468, gather :470-493, step :496-547);
called once per physics step from scripts/run_simulation.py:2256-2258.  Pinned against
tests/golden/phyto_golden.npz (recorded from the reference by tests/golden/make_golden.py phyto)."""
import numpy as np

from . import ops


def advect_diffuse(C, uo, vo, land_mask, dt, a, dlat, dlon, lat_rad_rows, adv_alpha=0.7, K_h=5.0e3):
    """C: [S, nlat, nlon] chlorophyll per species (returned updated)."""
    if dt <= 0.0:
        return C
    cosr = np.maximum(np.cos(lat_rad_rows), 0.5)                       # phyto.py:119
    ocean = (np.asarray(land_mask) == 0)
    C = np.array(C, dtype=float, copy=True)
    for s in range(C.shape[0]):
        Cs = C[s]
        adv = ops.advect_semilag(Cs, uo, vo, float(dt), a, dlat, dlon, cosr)
        new = (1.0 - adv_alpha) * Cs + adv_alpha * adv
        if K_h > 0.0:
            new = ops.nan_to_num(new)
            new = new + float(dt) * K_h * ops.laplacian(new, dlat, dlon, cosr, a)
        new = np.clip(new, 0.0, np.inf)
        new[~ocean] = 0.0
        C[s] = new
    for j in (0, -1):                                                  # polar ring scalar means, phyto.py:531-546
        row_ocean = ocean[j, :]
        if np.any(row_ocean):
            for s in range(C.shape[0]):
                C[s, j, row_ocean] = float(np.mean(C[s, j, :][row_ocean]))
    return C
