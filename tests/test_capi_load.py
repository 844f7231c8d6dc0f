"""The C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol that
include/a3d.h declares.  No compute calls here (no GPU in the build container)."""
import os
import re

from conftest import ROOT


def test_library_exports_every_declared_symbol(built):
    from mav_active_3d_planning_b200 import capi
    lib = capi.load_library()
    hdr = open(os.path.join(ROOT, "include", "a3d.h")).read()
    declared = sorted(set(re.findall(r"\b(a3d_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no prototypes found in include/a3d.h"
    assert sorted(capi.EXPORTS) == declared
    for sym in declared:
        assert hasattr(lib, sym), sym


def test_param_struct_layouts_match_header(built):
    # sizes implied by include/a3d.h (all fields 4- or 8-byte, no implicit padding)
    import ctypes as C
    from mav_active_3d_planning_b200 import capi
    assert C.sizeof(capi.CasterParams) == 4 * 4 + 8 * 4 + 8 * 3 + 8 * 4 + 8 * 2
    assert C.sizeof(capi.EvalParams) == 4 * 4 + 8 + 8 * 6 + 4 * 2 + 8 * 3 + 8 * 3 + 8 * 4 + 8 * 5
    assert C.sizeof(capi.CasterInfo) == 4 * 4 + 8 * 3 + 8 * 33 + 4 * 33 + 4


def test_no_device_is_a_loud_error(built):
    # In the CPU container context creation must fail with A3D_ERR_CUDA — never fall back.
    import torch
    from mav_active_3d_planning_b200 import capi
    if torch.cuda.is_available():
        return
    try:
        capi.Context(0)
    except capi.A3DError as e:
        assert e.code == -2
    else:
        raise AssertionError("context creation succeeded without a GPU")


def test_sm100a_sass_present(built):
    import subprocess
    from mav_active_3d_planning_b200 import capi
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
