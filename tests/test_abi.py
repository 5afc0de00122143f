"""CPU: the C-ABI shared library loads and exports every symbol include/omnigs_b200.h declares.
No compute entry point is called here (no GPU)."""
import ctypes
import os
import re

import _harness as h

HEADER = os.path.join(h.ROOT, "include", "omnigs_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"OGS_API\s+[\w\s\*]+?\b(ogs_\w+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for must in ["ogs_geom_bytes", "ogs_img_bytes", "ogs_binning_bytes", "ogs_lonlat_forward_stage1",
                 "ogs_lonlat_forward_stage2", "ogs_lonlat_backward", "ogs_mark_all_visible"]:
        assert must in syms


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(h.pkg.library_path())
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    # and the Python binding table covers exactly the header
    from importlib import import_module
    binding = import_module("omnigs-fork_b200._lib")
    assert sorted(binding.SYMBOLS) == declared_symbols()


def test_workspace_queries():
    lib = h.pkg.load_library()
    assert lib.ogs_abi_version() == 2
    assert lib.ogs_geom_bytes(0) > 0
    g1, g2 = lib.ogs_geom_bytes(1000), lib.ogs_geom_bytes(2000)
    assert g2 > g1 > 1000 * 100
    assert lib.ogs_img_bytes(2048, 1024) >= 2048 * 1024 * 8
    assert lib.ogs_img_bytes(0, 10) == 0
    b0, b1 = lib.ogs_binning_bytes(0, 256, 128), lib.ogs_binning_bytes(10_000_000, 2048, 1024)
    assert b1 >= 16 * 10_000_000 > b0
    # 16-byte keys+values per instance vs the reference's 24 + CUB temp
    assert b1 < 20 * 10_000_000


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from importlib import import_module
    binding = import_module("omnigs-fork_b200._lib")
    monkeypatch.setattr(binding, "_lib", None)
    monkeypatch.setattr(binding, "_HERE", str(tmp_path))
    import pytest
    with pytest.raises(binding.OgsError):
        binding.load_library()
