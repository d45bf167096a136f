"""ctypes binding of libqd_b200 (include/qd_b200.h).

Enum values (field / row / parameter / scalar slots) are parsed from the header itself so the
Python side can never drift from the C side.  The library is built in-tree by
``qingdai_b200.build`` (nvcc, sm_100a) -- there is no CPU fallback: loading fails loudly when
the shared object is missing, and ``qd_create`` fails when there is no CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "qd_b200.h")
DEFAULT_LIB = os.path.join(_HERE, "_lib", "libqd_b200.so")


def _parse_enums(path):
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for body in re.findall(r"typedef\s+enum\s*\{(.*?)\}\s*\w+\s*;", text, flags=re.S):
        val = -1
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                name, v = [s.strip() for s in item.split("=")]
                val = int(v, 0)
            else:
                name, val = item, val + 1
            out[name] = val
    return out


ENUM = _parse_enums(HEADER)
NF, NM, NR, NC, NP, NS = (ENUM[k] for k in ("QD_F_COUNT", "QD_M_COUNT", "QD_R_COUNT", "QD_C_COUNT", "QD_P_COUNT", "QD_S_COUNT"))


class Forcing(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("t", "flux_a", "sin_delta_a", "cos_delta_a", "alpha_a",
                                          "flux_b", "sin_delta_b", "cos_delta_b", "alpha_b", "theta")]


class StepCfg(C.Structure):
    _fields_ = [("dt", C.c_double), ("has_albedo", C.c_int), ("diff_enable", C.c_int),
                ("diff_every", C.c_int), ("k4_nsub", C.c_int), ("apply_q", C.c_int), ("apply_cloud", C.c_int),
                ("shapiro_every", C.c_int), ("shapiro_n", C.c_int), ("shapiro_q", C.c_int), ("shapiro_cloud", C.c_int),
                ("spec_every", C.c_int), ("spec_cutoff", C.c_double), ("spec_damp", C.c_double),
                ("oc_diff_every", C.c_int), ("oc_k4_nsub", C.c_int), ("oc_shapiro_n", C.c_int), ("oc_shapiro_every", C.c_int),
                ("oc_has_q", C.c_int), ("oc_has_ice", C.c_int),
                ("with_ocean", C.c_int), ("with_hydrology", C.c_int), ("with_routing", C.c_int), ("with_eco", C.c_int),
                ("loop_with_albedo", C.c_int), ("store_isr_ab", C.c_int)]


_P = C.c_void_p
_D = C.c_double
_I = C.c_int

# name -> (restype, argtypes); every symbol declared in include/qd_b200.h must appear here
PROTOTYPES = {
    "qd_create": (_I, [_I, _I, _I, _I, _D, _D, _D, _D, _D, _P, _P, _P, C.POINTER(_P)]),
    "qd_destroy": (_I, [_P]),
    "qd_last_error": (C.c_char_p, [_P]),
    "qd_version": (_I, []),
    "qd_set_stream": (_I, [_P, _P]),
    "qd_synchronize": (_I, [_P]),
    "qd_bind": (_I, [_P, _P, _P]),
    "qd_set_params": (_I, [_P, _P]),
    "qd_set_rows": (_I, [_P, _P]),
    "qd_set_rows_member": (_I, [_P, _I, _P]),
    "qd_get_scalars": (_I, [_P, _P]),
    "qd_upload_field": (_I, [_P, _I, _I, _P]),
    "qd_download_field": (_I, [_P, _I, _I, _P]),
    "qd_upload_mask": (_I, [_P, _I, _I, _P]),
    "qd_download_mask": (_I, [_P, _I, _I, _P]),
    "qd_laplacian": (_I, [_P, _P, _P, _P]),
    "qd_hyperdiffuse": (_I, [_P, _P, _P, _P, _D, _D, _I, _P]),
    "qd_advect": (_I, [_P, _P, _P, _P, _P, _D, _P]),
    "qd_shapiro": (_I, [_P, _P, _P, _I]),
    "qd_gaussian": (_I, [_P, _P, _P, _I, _I, _P]),
    "qd_set_gauss": (_I, [_P, _I, _I, _I, _P]),
    "qd_zonal_bandstop": (_I, [_P, _P, _D, _D]),
    "qd_divergence": (_I, [_P, _P, _P, _P]),
    "qd_vorticity": (_I, [_P, _P, _P, _P]),
    "qd_median_pos": (_I, [_P, _P, _D, _P]),
    "qd_wsum": (_I, [_P, _P, _P]),
    "qd_median_stats": (_I, [_P, _P]),
    "qd_minmax": (_I, [_P, _P, _P]),
    "qd_math_check": (_I, [_P, _P, C.c_longlong, _P, _I]),
    "qd_band_exchange_bench": (_I, [_P, _I, _I, _P]),
    "qd_row_dev": (_P, [_P, _I]),
    "qd_user_row": (_P, [_P, _I, _P]),
    "qd_user_row_member": (_I, [_P, _I, _I, _P]),
    "qd_laplacian_host": (_I, [_P, _P, _P, _P]),
    "qd_hyperdiffuse_host": (_I, [_P, _P, _P, _P, _D, _D, _I, _P]),
    "qd_advect_host": (_I, [_P, _P, _P, _P, _P, _D, _P]),
    "qd_atmos_step": (_I, [_P, C.POINTER(StepCfg)]),
    "qd_ocean_step": (_I, [_P, C.POINTER(StepCfg)]),
    "qd_ocean_step_winds": (_I, [_P, C.POINTER(StepCfg), _P, _P]),
    "qd_loop_step": (_I, [_P, C.POINTER(StepCfg), C.POINTER(Forcing), _I]),
    "qd_last_nsub": (_I, [_P, _P]),
    "qd_use_graphs": (_I, [_P, _I]),
    "qd_graph_status": (_I, [_P, C.POINTER(_I), C.POINTER(_I)]),
    "qd_set_counters": (_I, [_P, _I, _I, _I]),
    "qd_get_counters": (_I, [_P, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "qd_set_gauss2d": (_I, [_P, _I]),
    "qd_set_h4_stream": (_I, [_P, _I]),
    "qd_set_ocean_fused": (_I, [_P, _I]),
    "qd_launch_count": (_I, [_P, C.POINTER(C.c_longlong)]),
    "qd_profile": (_I, [_P, _I]),
    "qd_profile_report": (_I, [_P, C.c_char_p, _I]),
    "qd_band_init": (_I, [_P, _I, _I, _I]),
    "qd_band_export": (_I, [_P, _P]),
    "qd_band_connect": (_I, [_P, _P]),
    "qd_band_info": (_I, [_P, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "qd_eco_bind": (_I, [_P, _P, _I, _D, _D, _D, _I]),
    "qd_eco_reset": (_I, [_P, _D, _D, _I, _I]),
    "qd_eco_state": (_I, [_P, _P, _P, _I]),
    "qd_eco_subdaily": (_I, [_P, _P, _D, _P, C.POINTER(_I)]),
    "qd_eco_bands": (_I, [_P, _I, _P, _D, _P]),
    "qd_indiv_setup": (_I, [_P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "qd_indiv_substep": (_I, [_P, _P, _D, _D, _D]),
    "qd_indiv_state": (_I, [_P, _P, _P, _I]),
    "qd_net_build": (_I, [_I, _I, _P, _P, _P, _I, _D, _P, _P, _P, _P, _P, _P, C.POINTER(_I), C.POINTER(_I)]),
    "qd_diag_count": (_I, []),
    "qd_diag": (_I, [_P, _P]),
    "qd_phyto_advect_diffuse": (_I, [_P, _P, _I, _P, _P, _D, _D, _D]),
    "qd_route_setup": (_I, [_P, _I, _P, _P, _P, _P, _P, _I, _P]),
    "qd_route_levels": (_I, [_P]),
    "qd_route_accumulate": (_I, [_P, _D]),
    "qd_route_event": (_I, [_P, _I, _P, _P, _P, _P, _P]),
    "qd_route_buffer": (_I, [_P, _I, _P, _I]),
}


def header_symbols(path=HEADER):
    """Every function name declared in the public header."""
    text = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(qd_[a-z0-9_]+)\s*\(", text)))


class Library:
    """A loaded libqd_b200 with typed entry points.  ``host_emulation`` is only ever true for the
    test scaffolding under tests/hostcheck (device pointers are then host pointers)."""

    def __init__(self, path=None, host_emulation=False):
        path = path or os.environ.get("QD_B200_LIB") or DEFAULT_LIB          # QD_B200_LIB: a tuning variant built by build.py --out
        if not os.path.exists(path):
            raise RuntimeError(
                f"libqd_b200 not found at {path}: build it with `python -m qingdai_b200.build` "
                "(nvcc, sm_100a).  There is no CPU fallback.")
        self.path = path
        self.host_emulation = bool(host_emulation)
        self.dll = C.CDLL(path)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(self.dll, name)
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)

    def check(self, ctx, rc, what=""):
        if rc != 0:
            msg = self.qd_last_error(ctx)
            raise RuntimeError(f"libqd_b200 {what} failed (status {rc}): {msg.decode() if msg else ''}")


_default = None


def default_library():
    global _default
    if _default is None:
        _default = Library()
    return _default
