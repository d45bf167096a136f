"""The reference's OWN script loop on the drop-in classes (north_star: "drop-in behind the existing pygcm entry points:
scripts.run_simulation ...").  ``scripts.run_simulation.main()`` is imported from the reference checkout and run
UNMODIFIED for 16 steps on a 31x60 grid, with the four classes it builds -- SphericalGrid, SpectralModel,
WindDrivenSlabOcean, RiverRouting (run_simulation.py:19-36) -- replaced by the qingdai_b200 classes, exactly the import
swap INTEGRATION.md describes.  Everything else of the script (orbital forcing, precipitation / cloud diagnosis, albedo,
hydrology, the attribute reads and writes ``gcm.T_s = ...`` :2253, ``gcm.cloud_cover = ...`` :1900, ``gcm.q``, ``gcm.h_ice``)
stays the reference's NumPy code talking to the device state through the property protocol.  The states it plots must
follow the recording of the unmodified reference (tests/golden/loop_golden.npz) to the free-running bar of DESIGN
section 2 (1e-7: chaotic growth of last-bit differences; single steps hold 1e-12 in the teacher-forced tests).

CPU box only: needs /root/reference (skipped elsewhere); the library is the host check build of the kernel sources."""
import contextlib
import io
import os
import sys
import tempfile
from unittest import mock

import numpy as np
import pytest

REF = os.environ.get("QD_REFERENCE_ROOT", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "scripts")), reason="reference checkout not present")

QUIET_ENV = {
    "QD_ENERGY_DIAG": "0", "QD_HUMIDITY_DIAG": "0", "QD_WATER_DIAG": "0", "QD_OCEAN_DIAG": "0", "QD_OCEAN_ENERGY_DIAG": "0",
    "QD_HYDRO_DIAG": "0", "QD_ECO_DIAG": "0", "QD_PHYTO_DIAG": "0", "QD_USE_JAX": "0", "QD_AUTOSAVE_ENABLE": "0", "QD_AUTOSAVE_LOAD": "0",
    "QD_PHYTO_ENABLE": "0", "QD_ECO_INDIV_ENABLE": "0", "MPLBACKEND": "Agg", "QD_ECO_ENABLE": "0", "QD_HYDRO_ENABLE": "0",
}
CASES = {"base": {}, "banded": {"QD_INIT_BANDED": "1", "QD_INIT_T_POLE": "255.0", "QD_DT_SECONDS": "900", "QD_OROG": "1"}}


def _import_reference_script():
    if "matplotlib" not in sys.modules:
        m = mock.MagicMock()
        m.pyplot.subplots.side_effect = lambda *a, **k: (mock.MagicMock(), mock.MagicMock())
        sys.modules["matplotlib"] = m
        sys.modules["matplotlib.pyplot"] = m.pyplot
        for sub in ("colors", "cm", "gridspec", "ticker", "patches"):
            sys.modules[f"matplotlib.{sub}"] = getattr(m, sub)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import scripts.run_simulation as rs
    return rs


@pytest.mark.parametrize("tag", ["base", "banded"])
def test_reference_main_runs_on_the_dropin_classes(tag, golden, monkeypatch):
    from hostcheck import library
    from qingdai_b200 import _binding
    from qingdai_b200.dynamics import SpectralModel
    from qingdai_b200.grid import SphericalGrid
    from qingdai_b200.ocean import WindDrivenSlabOcean
    from qingdai_b200.routing import RiverRouting
    L = golden("loop_golden.npz")
    nlat, nlon, nsteps = int(L["nlat"]), int(L["nlon"]), int(L[f"{tag}_nsteps"])
    for k in list(os.environ):
        if k.startswith("QD_"):
            monkeypatch.delenv(k)
    for k, v in {**QUIET_ENV, **CASES[tag]}.items():
        monkeypatch.setenv(k, v)
    dt = int(os.environ.get("QD_DT_SECONDS", "300"))
    monkeypatch.setenv("QD_PLOT_EVERY_DAYS", "1e-9")
    monkeypatch.setenv("QD_SIM_DAYS", repr((nsteps - 0.5) * dt / (2 * np.pi / 8.726646259971648e-5)))
    monkeypatch.setattr(_binding, "_default", library())          # the drop-in classes bind to the default library
    rs = _import_reference_script()
    rec = []

    def hook_plot_state(grid, gcm, land_mask, precip, cloud_cover, albedo, t_days, output_dir, ocean=None, routing=None):
        d = {k: np.array(getattr(gcm, k), copy=True) for k in ("u", "v", "h", "T_s", "q", "cloud_cover", "h_ice")}
        d.update({k: np.array(getattr(ocean, k), copy=True) for k in ("uo", "vo", "eta", "Ts")})
        d["precip"], d["albedo"] = np.array(precip, copy=True), np.array(albedo, copy=True)
        rec.append(d)

    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp, contextlib.redirect_stdout(io.StringIO()), \
            mock.patch.object(rs, "SphericalGrid", lambda n_lat, n_lon: SphericalGrid(nlat, nlon)), \
            mock.patch.object(rs, "SpectralModel", SpectralModel), \
            mock.patch.object(rs, "WindDrivenSlabOcean", WindDrivenSlabOcean), \
            mock.patch.object(rs, "RiverRouting", RiverRouting), \
            mock.patch.object(rs, "plot_state", hook_plot_state), \
            mock.patch.object(rs, "plot_true_color", lambda *a, **k: None), \
            mock.patch.object(rs, "plot_ecology", lambda *a, **k: None):
        os.chdir(tmp)
        try:
            rs.main()
        finally:
            os.chdir(cwd)
    assert len(rec) == nsteps
    worst = 0.0
    for i in (0, 1, 2, 5, 6, 14, 15):
        for k, v in rec[i].items():
            ref = L[f"{tag}_s{i}_{k}"]
            scale = max(float(np.max(np.abs(ref))), 1e-300)
            err = float(np.max(np.abs(v - ref))) / scale
            worst = max(worst, err)
            assert err < 1e-7, (tag, i, k, err)
    print(f"reference main() on the drop-in classes, {tag}: worst relative deviation over 16 steps {worst:.2e}")
