"""CPU: the C-ABI library loads, exports every symbol include/snb_b200.h declares, and its host-only entry points
(crop plan, argument validation) behave like the reference -- no compute call is made without a GPU."""
import ctypes
import os
import re

import pytest

import snb_b200  # noqa: F401
from snb_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "snb_b200.h")) as fh:
        src = fh.read()
    return sorted(set(re.findall(r"SNB_API[^;(]*?\b(snb_\w+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    syms = declared_symbols()
    assert len(syms) >= 20
    handle = ctypes.CDLL(N.LIB_PATH)
    for s in syms:
        assert hasattr(handle, s), "libsnb_b200.so does not export %s" % s
    assert sorted(N.SIGNATURES) == syms          # the ctypes table binds exactly the header's surface
    assert N.lib().snb_version() == 100


def test_slicer_plan_matches_reference_kats(kats):
    lib = N.lib()
    for c in kats["slicer"]["cases"]:
        h = ctypes.c_void_p()
        N.check(lib.snb_slicer_create(c["shape"][0], c["shape"][1], c["tile"], c["step"], c["margin"], ctypes.byref(h)))
        info = (ctypes.c_int64 * 8)()
        N.check(lib.snb_slicer_info(h, info))
        assert list(info[:4]) == c["margins"]
        assert info[4] == c["n_crops"] and info[5] * info[6] == c["n_crops"] and info[7] == c["tile"]
        xy = (ctypes.c_int64 * (2 * c["n_crops"]))()
        N.check(lib.snb_slicer_crops(h, xy))
        crops = [[xy[2 * i], xy[2 * i + 1], c["tile"], c["tile"]] for i in range(c["n_crops"])]
        assert crops[:3] == c["crops_head"] and crops[-3:] == c["crops_tail"]
        assert [sum(v[0] for v in crops), sum(v[1] for v in crops)] == c["crops_sum"]
        lib.snb_slicer_destroy(h)


def test_slicer_errors_map_to_value_error(kats):
    lib = N.lib()
    for c in kats["slicer"]["errors"]:
        h = ctypes.c_void_p()
        rc = lib.snb_slicer_create(c["shape"][0], c["shape"][1], c["tile"], c["step"], c["margin"], ctypes.byref(h))
        if c["error"] is None:
            assert rc == 0
            lib.snb_slicer_destroy(h)
        else:
            assert rc == N.SNB_E_INVALID and h.value is None
            with pytest.raises(ValueError):
                N.check(rc)
            assert N.last_error()


def test_error_code_mapping():
    for rc, exc in [(N.SNB_E_INVALID, ValueError), (N.SNB_E_SHAPE, AssertionError), (N.SNB_E_CUDA, RuntimeError),
                    (N.SNB_E_UNSUPPORTED, NotImplementedError)]:
        with pytest.raises(exc):
            N.check(rc)
    N.check(0)


def test_argument_validation_without_gpu():
    lib = N.lib()
    h = ctypes.c_void_p()
    N.check(lib.snb_slicer_create(64, 64, 32, 16, 0, ctypes.byref(h)))
    # null device pointers and bad ranges are rejected before anything is launched
    assert lib.snb_split_hwc(h, None, 3, 1, 0, None, None, 0, 1, None) == N.SNB_E_INVALID
    assert lib.snb_merge(h, None, N.DT_F32, 1, 1, None, None, N.DT_F32, None, 0.5, None) == N.SNB_E_INVALID
    assert lib.snb_loss_iou_reduce(None, None, N.DT_I64, 4, -1.0, None, None, None, None, None) == N.SNB_E_INVALID
    assert lib.snb_reduce_workspace_bytes() >= 64 * 2048
    d = N.ConvDesc()
    out = ctypes.c_void_p()
    assert lib.snb_conv_create(ctypes.byref(d), ctypes.byref(out)) == N.SNB_E_INVALID
    lib.snb_slicer_destroy(h)
