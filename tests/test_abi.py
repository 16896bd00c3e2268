"""CPU tests of the drop-in boundary: the C ABI library loads and exports every symbol
include/petar_b200.h declares; the C++ shim defines the symbols PeTar's force_gpu_cuda.hpp
declares; without a GPU the product fails loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from petar_b200 import engine
from petar_b200.types import EPJSoft, SPJQuad
from conftest import ROOT, _have_gpu


def _header_functions():
    src = open(os.path.join(ROOT, "include", "petar_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported():
    L = engine.load()
    names = _header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/petar_b200.h but not exported"
    assert sorted(engine.ABI_SYMBOLS) == names
    assert L.pb_abi_version() == engine.ABI_VERSION == 6


def test_shim_defines_petar_symbols():
    """The shim must define what reference src/force_gpu_cuda.hpp:120-132, 154-161, 165-168 declare and
    the globals of src/force_gpu_cuda.cu:7-10."""
    engine.load_shim()
    engine.load_shim(direct=True)
    out = subprocess.run(["nm", "-DC", os.path.join(engine.LIBDIR, "libpetar_b200_shim.so")], capture_output=True, text=True).stdout
    assert "CalcForceWithLinearCutoffCUDAMultiWalk::operator()(int, int, pb_EPISoft const**, int const*, int const**, int const*, int const**, int const*, pb_EPJSoft const*, int, pb_SPJQuad const*, int, bool)" in out
    assert "RetrieveForceCUDA(int, int, int const*, pb_ForceSoft**)" in out
    assert re.search(r"\bB gpu_profile\b", out) and re.search(r"\bB gpu_counter\b", out)
    out = subprocess.run(["nm", "-DC", os.path.join(engine.LIBDIR, "libpetar_b200_shim_direct.so")], capture_output=True, text=True).stdout
    assert "CalcForceWithLinearCutoffCUDA::operator()(int, int, pb_EPISoft const**, int const*, pb_EPJSoft const**, int const*, pb_SPJQuad const**, int const*)" in out


def test_library_contains_sm100a_code_only():
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(engine.LIBDIR, "libpetar_b200.so")], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly():
    L = engine.load()
    rc = L.pb_init(0, -1)
    assert rc == -1                                   # PB_ERR_NO_DEVICE
    assert b"no CPU fallback" in L.pb_last_error()
    epj = np.zeros(4, dtype=EPJSoft)
    spj = np.zeros(4, dtype=SPJQuad)
    rc = L.pb_upload_j(epj.ctypes.data, 4, C.byref(engine.LAYOUT_EPJ), spj.ctypes.data, 4, C.byref(engine.LAYOUT_SPJ))
    assert rc == -1
    with pytest.raises(engine.PbError):
        engine.check(rc, "pb_upload_j")


def test_options_validate():
    L = engine.load()
    import ctypes as C
    v = C.c_longlong(-1)
    assert L.pb_get_option(b"coords", C.byref(v)) == 0 and v.value == 2           # the drop-in default
    assert L.pb_set_option(b"coords", 1) == 0 and L.pb_set_option(b"coords", 0) == 0
    assert L.pb_get_option(b"coords", C.byref(v)) == 0 and v.value == 0
    assert L.pb_set_option(b"coords", 2) == 0
    assert L.pb_set_option(b"coords", 3) == -3 and L.pb_get_option(b"nosuchkey", C.byref(v)) == -3
    assert L.pb_set_option(b"streams", 0) == -3
    assert L.pb_set_option(b"nosuchkey", 1) == -3
    # kernel variants and the in-place options: documented defaults, range checks (no GPU needed to set them)
    for key, default, bad in ((b"ws", 1, 2), (b"sp2i", 1, 2), (b"fuse_reduce", 1, 2), (b"raw_result", 0, 2), (b"raw_upload", 0, 2), (b"ep_runs", 0, 2)):
        assert L.pb_get_option(key, C.byref(v)) == 0 and v.value == default, key
        assert L.pb_set_option(key, bad) == -3 and L.pb_set_option(key, 1 - default) == 0 and L.pb_set_option(key, default) == 0, key
    assert L.pb_set_params(-1.0, 0.0, 1.0) == -3


def test_host_packers_hi_lo_split_is_exact():
    """Device j format: x = x_hi + x_lo with x_hi = float(x); the pair must reproduce x to ~2^-48."""
    L = engine.load()
    rng = np.random.default_rng(0)
    n = 1000
    epj = np.zeros(n, dtype=EPJSoft)
    epj["pos"] = rng.normal(size=(n, 3)) * 3.0
    epj["mass"] = rng.random(n)
    epj["r_search"] = rng.random(n) * 0.01
    out = np.zeros((n, 8), dtype=np.float32)
    assert L.pb_pack_epj_host(epj.ctypes.data, n, C.byref(engine.LAYOUT_EPJ), out.ctypes.data) == 0
    assert np.array_equal(out[:, 0:3], epj["pos"].astype(np.float32))
    rec = out[:, 0:3].astype(np.float64) + out[:, 4:7].astype(np.float64)
    assert np.abs(rec - epj["pos"]).max() < 2.0 ** -45
    assert np.array_equal(out[:, 3], epj["mass"].astype(np.float32))
    assert np.array_equal(out[:, 7], epj["r_search"].astype(np.float32))

    spj = np.zeros(n, dtype=SPJQuad)
    spj["pos"] = rng.normal(size=(n, 3))
    spj["mass"] = rng.random(n)
    spj["quad"] = rng.normal(size=(n, 6))
    o2 = np.zeros((n, 16), dtype=np.float32)
    assert L.pb_pack_spj_host(spj.ctypes.data, n, C.byref(engine.LAYOUT_SPJ), o2.ctypes.data) == 0
    q = spj["quad"]
    tr = q[:, 0] + q[:, 1] + q[:, 2]                                                   # all in fp64, then cast
    assert np.array_equal(o2[:, 7], (3 * q[:, 0] - tr).astype(np.float32))             # q'xx = 3 qxx - tr
    assert np.array_equal(o2[:, 8], (3 * q[:, 1] - tr).astype(np.float32))             # q'yy
    assert np.array_equal(o2[:, 9], (3 * q[:, 2] - tr).astype(np.float32))             # q'zz
    assert np.array_equal(o2[:, 10:12], (3 * q[:, 3:5]).astype(np.float32))            # q'xy q'xz
    assert np.array_equal(o2[:, 12], (3 * q[:, 5]).astype(np.float32))                 # q'yz
    assert np.array_equal(o2[:, 13], tr.astype(np.float32))
    assert np.abs(o2[:, 7] + o2[:, 8] + o2[:, 9]).max() < 1e-5                         # traceless


def test_indexed_packer_equals_gather_then_pack():
    L = engine.load()
    rng = np.random.default_rng(3)
    n = 5000
    epj = np.zeros(n, dtype=EPJSoft)
    epj["pos"] = rng.normal(size=(n, 3))
    epj["mass"] = rng.random(n)
    epj["r_search"] = rng.random(n)
    idx = rng.integers(0, n, 1234).astype(np.int64)
    a = np.zeros((len(idx), 8), dtype=np.float32)
    b = np.zeros((len(idx), 8), dtype=np.float32)
    g = np.ascontiguousarray(epj[idx])
    assert L.pb_pack_epj_host(g.ctypes.data, len(idx), C.byref(engine.LAYOUT_EPJ), a.ctypes.data) == 0
    assert L.pb_pack_epj_host_indexed(epj.ctypes.data, idx.ctypes.data, len(idx), C.byref(engine.LAYOUT_EPJ), b.ctypes.data) == 0
    assert np.array_equal(a, b)


def test_layouts_match_petar_structs():
    assert (engine.LAYOUT_EPI.stride, engine.LAYOUT_EPI.off_pos, engine.LAYOUT_EPI.off_rsearch) == (48, 8, 32)
    assert (engine.LAYOUT_EPJ.stride, engine.LAYOUT_EPJ.off_pos, engine.LAYOUT_EPJ.off_mass, engine.LAYOUT_EPJ.off_rsearch) == (120, 16, 8, 80)
    assert (engine.LAYOUT_SPJ.stride, engine.LAYOUT_SPJ.off_pos, engine.LAYOUT_SPJ.off_mass, engine.LAYOUT_SPJ.off_quad) == (80, 8, 0, 32)
    assert (engine.LAYOUT_FORCE.stride, engine.LAYOUT_FORCE.off_acc, engine.LAYOUT_FORCE.off_pot, engine.LAYOUT_FORCE.off_nngb) == (40, 0, 24, 32)


def test_header_abi_version_matches_binding():
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "petar_b200.h")).read()
    assert int(re.search(r"#define\s+PB_ABI_VERSION\s+(\d+)", hdr).group(1)) == engine.ABI_VERSION


def test_graft_entry_build():
    """The driver's "does it build" check: compiles every native library (and the checker) and loads them."""
    import __graft_entry__ as g
    g.build()
