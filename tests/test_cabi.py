"""The C-ABI library loads, exports every symbol include/prtc.h declares, keeps the reference-compatible struct
layouts, and FAILS LOUDLY (no CPU fallback) when there is no device."""
import ctypes
import os
import re

import numpy as np
import pytest

from prototype2_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "prtc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(prtc_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    lib = capi.load_library()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libprtc.so does not export %s" % n
    assert sorted(capi.SYMBOLS) == names, "capi.SYMBOLS out of date with include/prtc.h"


def test_struct_layouts_match_the_reference_types():
    # Transform = vec3 + quat + vec3 (component.h:14-17) = 40 bytes, rotation stored x,y,z,w at offset 12
    assert capi.TRANSFORM_DTYPE.itemsize == 40
    assert capi.TRANSFORM_DTYPE.fields["rotation"][1] == 12 and capi.TRANSFORM_DTYPE.fields["scale"][1] == 28
    assert capi.PHYSICS_DTYPE.itemsize == 44
    assert ctypes.sizeof(capi.Config) == 40 and ctypes.sizeof(capi.Counters) == 56


def test_default_config_holds_the_reference_constants():
    cfg = capi.Config()
    capi.lib().prtc_default_config(ctypes.byref(cfg))
    assert (cfg.gravity, cfg.substeps, cfg.friction) == (1.0, 4, 10.0)
    assert abs(cfg.leaf_margin - 0.05) < 1e-9 and abs(cfg.ground_dot - 0.3) < 1e-7 and abs(cfg.stick - 0.05) < 1e-9


def test_no_device_means_error_not_fallback():
    if capi.device_count() > 0:
        pytest.skip("a B200 is present")
    with pytest.raises(capi.PrtcError) as e:
        capi.PhysicsContext()
    assert e.value.code == capi.PRTC_ERR_CUDA and "no CPU path" in str(e.value)


def test_missing_library_raises(monkeypatch):
    monkeypatch.setattr(capi, "LIB_PATH", os.path.join(ROOT, "prototype2_b200", "does_not_exist.so"))
    with pytest.raises(ImportError):
        capi.load_library()


def test_product_does_not_reference_the_oracle():
    """the product path must never import, link or execute anything under oracle/"""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "prototype2_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"oraclelib|oracle/|l1_oracle|libprt_l[01]", text):
                    bad.append(os.path.join(base, f))
    assert bad == []
    for inc in ("include/prtc.h",):
        assert "oracle" not in open(os.path.join(ROOT, inc)).read().lower()


def test_invalid_arguments_are_rejected_without_a_device():
    lib = capi.lib()
    assert lib.prtc_create(None, None) == capi.PRTC_ERR_INVALID
    n = ctypes.c_uint32()
    assert lib.prtc_host_build_tree(None, ctypes.c_uint32(3), ctypes.c_float(0.05), None, None, ctypes.c_uint32(0), ctypes.byref(n)) == capi.PRTC_ERR_INVALID
    assert b"null" in lib.prtc_last_error()
