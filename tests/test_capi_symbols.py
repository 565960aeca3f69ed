"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU and exports exactly the
symbols include/stormb200.h declares; no compute is called here."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "stormb200.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"SB_API\s+[\w\s\*]+?\b(sb_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from stormruler_b200 import build, capi
    build.build()
    return capi.load()


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("sb_ctx_create", "sb_vec_alloc", "sb_op_create", "sb_apply", "sb_eval", "sb_dot",
                 "sb_norm2", "sb_cg_solve", "sb_bicgstab_solve", "sb_solve_host"):
        assert must in syms


def test_header_is_plain_c():
    """The drop-in boundary is a C ABI: include/stormb200.h must compile as C99 (and as C++) on its own."""
    for cmd in (["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c", HEADER],
                ["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", HEADER]):
        res = subprocess.run(cmd, capture_output=True, text=True)
        assert res.returncode == 0, res.stderr


def test_library_exports_every_declared_symbol(lib):
    from stormruler_b200 import capi
    syms = declared_symbols()
    assert sorted(capi.SIGNATURES) == syms, "python binding and header disagree"
    for s in syms:
        assert hasattr(lib, s), f"{s} not exported by libstormb200.so"
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (sb_\w+)", out))
    assert exported == set(syms), f"unexpected exports: {sorted(exported ^ set(syms))}"


def test_library_is_built_for_sm_100a():
    from stormruler_b200 import capi
    out = subprocess.run(["cuobjdump", "--list-elf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_no_silent_cpu_fallback(lib):
    """Without a CUDA device the context refuses to come up (and says why); with one it works."""
    import ctypes as C
    h = C.c_void_p()
    rc = lib.sb_ctx_create(0, C.byref(h))
    if rc == 0:
        lib.sb_ctx_destroy(h)
    else:
        assert rc < 0 and h.value is None and len(lib.sb_last_error()) > 0
    assert lib.sb_ctx_create(10_000, C.byref(h)) < 0
    assert lib.sb_version() >= 100


def test_dropin_library_loads_when_built():
    """libstorm_dropin.so (reference solver templates on DeviceVector) is built wherever the reference
    tree is mounted; it must load and export its entry points without a GPU."""
    from stormruler_b200 import dropin
    if not dropin.available():
        pytest.skip("libstorm_dropin.so not built (needs the StormRuler sources at build time)")
    L = dropin.load()
    for name in ("dropin_solve", "dropin_last_error", "dropin_reset_rng", "dropin_selftest_errors"):
        assert hasattr(L, name)
