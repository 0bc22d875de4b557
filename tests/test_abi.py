"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/hyprec_b200.h
declares.  No compute here (no GPU); on a box without a device, handle creation must fail loudly
instead of falling back to the CPU."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from hyprec_b200 import _native, build


def declared_functions():
    text = open(os.path.join(ROOT, "include", "hyprec_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hyprec_[a-z_0-9]+)\s*\(", text)) - {"hyprec_epoch_cb"})


def test_library_builds_and_exports_the_header():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = declared_functions()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(_native.EXPORTS) == names
    lib.hyprec_abi_version.restype = ctypes.c_int
    assert lib.hyprec_abi_version() == 2


def test_no_cpu_fallback():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present; the failure path is for boxes without one")
    with pytest.raises(_native.HyprecError) as err:
        _native.Handle(0, _native.F64)
    assert "no CPU fallback" in str(err.value)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "hyprec_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
