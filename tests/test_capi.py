"""CPU-side checks of the C-ABI boundary: the shared library builds/loads and exports every symbol
include/hgt_b200.h declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest

from tests.conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hgt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hgt_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = _declared_symbols()
    for must in ("hgt_last_error", "hgt_plan_nodes", "hgt_plan_edges_sort", "hgt_plan_edges_fill", "hgt_plan_tiles",
                 "hgt_fold_weights", "hgt_typed_linear", "hgt_edge_forward", "hgt_update_epilogue"):
        assert must in syms


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as ge
    path = ge.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in _declared_symbols():
        assert hasattr(lib, name), "libhgt_b200.so does not export %s" % name
    lib.hgt_abi_version.restype = ctypes.c_int
    assert lib.hgt_abi_version() >= 1


def test_ctypes_signatures_cover_the_header():
    from pyhgt_b200 import _lib
    declared = set(_declared_symbols()) - {"hgt_last_error", "hgt_kernel_launches", "hgt_conv_args_size",
                                           "hgt_sampler_budget_update", "hgt_sampler_add_budget"}      # non-status return types, bound in load()
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))


def test_error_channel_reports_bad_arguments():
    from pyhgt_b200 import _lib
    lib = _lib.load()
    out = ctypes.c_size_t()
    rc = lib.hgt_plan_workspace_bytes(-1, 0, ctypes.byref(out))
    assert rc != 0
    assert b"int32" in lib.hgt_last_error()
    with pytest.raises(_lib.HgtError):
        _lib.call("hgt_plan_workspace_bytes", 2 ** 40, 0, ctypes.byref(out))


def test_no_cpu_fallback():
    import torch
    import pyhgt_b200
    from pyhgt_b200 import _lib
    m = pyhgt_b200.HGTConv(16, 16, 2, 1, 2).eval()
    with torch.no_grad(), pytest.raises(_lib.HgtError):
        m(torch.randn(3, 16), torch.zeros(3, dtype=torch.long), torch.zeros(2, 1, dtype=torch.long),
          torch.zeros(1, dtype=torch.long), torch.zeros(1, dtype=torch.long))


def test_conv_args_struct_layout_matches_c():
    """The ctypes mirror of hgt_conv_args has the C struct's size, and its field list follows the header's order."""
    from pyhgt_b200 import _lib
    lib = _lib.load()
    assert ctypes.sizeof(_lib.ConvArgs) == lib.hgt_conv_args_size()
    text = open(os.path.join(ROOT, "include", "hgt_b200.h")).read()
    body = text[text.index("typedef struct {", text.index("Whole layer in one call")):text.index("} hgt_conv_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S).replace("typedef struct {", "")
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        parts = [p_.strip() for p_ in decl.split(",")]
        names.append(parts[0].split()[-1].lstrip("*"))
        names += [p_.lstrip("*").strip() for p_ in parts[1:]]
    assert names == [f[0] for f in _lib.ConvArgs._fields_], names
