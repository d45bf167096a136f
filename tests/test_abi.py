"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/qd_b200.h declares
(no compute calls: there is no GPU in CPU-only CI)."""
import ctypes
import os

import pytest

from qingdai_b200 import _binding


@pytest.fixture(scope="module")
def so_path():
    from qingdai_b200.build import build
    return build()


def test_header_and_prototypes_agree():
    declared = set(_binding.header_symbols())
    assert declared == set(_binding.PROTOTYPES), declared ^ set(_binding.PROTOTYPES)


def test_library_exports_every_declared_symbol(so_path):
    dll = ctypes.CDLL(so_path)
    for name in _binding.header_symbols():
        assert hasattr(dll, name), name
    dll.qd_version.restype = ctypes.c_int
    assert dll.qd_version() >= 100


def test_enum_tables_are_consistent():
    E = _binding.ENUM
    assert E["QD_F_COUNT"] == max(v for k, v in E.items() if k.startswith("QD_F_") and k != "QD_F_COUNT") + 1
    assert E["QD_P_COUNT"] == max(v for k, v in E.items() if k.startswith("QD_P_") and k != "QD_P_COUNT") + 1
    assert ctypes.sizeof(_binding.Forcing) == 80


def test_product_refuses_to_run_without_cuda(so_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from qingdai_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(10, 20)


def test_missing_library_is_loud(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _binding.Library(str(tmp_path / "nope.so"))
