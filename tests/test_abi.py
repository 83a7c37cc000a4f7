"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports exactly the symbols
include/msfm_match.h declares.  No compute calls (there is no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "msfm_match.h")


def header_functions(name="msfm_match.h"):
    src = open(os.path.join(ROOT, "include", name)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    return sorted(set(re.findall(r"\b(msfm_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(native_lib):
    from metricsfm_b200 import _lib
    assert header_functions() == sorted(_lib.EXPORTED_SYMBOLS)
    assert header_functions("msfm_sched.h") == sorted(_lib.SCHED_SYMBOLS)
    assert header_functions("msfm_multi.h") == sorted(_lib.MULTI_SYMBOLS)


def test_scheduler_and_multi_gpu_engine_are_exported(native_lib):
    """include/msfm_sched.h + msfm_multi.h: exported by the CUDA library; the scheduler also by the host-only library."""
    from metricsfm_b200 import _lib
    from metricsfm_b200.build import LIB_PATH, SCHED_LIB, build_host_libs
    build_host_libs()
    out = subprocess.check_output(["nm", "-D", "--defined-only", LIB_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    for name in _lib.SCHED_SYMBOLS + _lib.MULTI_SYMBOLS:
        assert name in exported, name
    assert not [s for s in exported if "internal" in s]          # the inter-unit helpers stay hidden
    host = subprocess.check_output(["nm", "-D", "--defined-only", SCHED_LIB], text=True)
    for name in _lib.SCHED_SYMBOLS:
        assert f" T {name}" in host


def test_library_exports_every_declared_symbol(native_lib):
    from metricsfm_b200.build import LIB_PATH
    out = subprocess.check_output(["nm", "-D", "--defined-only", LIB_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    for name in header_functions():
        assert name in exported, f"{name} declared in msfm_match.h but not exported"
        assert hasattr(native_lib, name)
    # plain C linkage only: no mangled C++ symbols leak out of the boundary
    assert not [s for s in exported if s.startswith("_Z") and "msfm" in s and "kernel" not in s and "device_stub" not in s]


def test_abi_version_and_status_strings(native_lib):
    assert native_lib.msfm_abi_version() == 3  # v2: msfm_config.keep_float, msfm_params.rescore_band
    assert native_lib.msfm_status_string(0) == b"ok"
    assert b"sm_100" in native_lib.msfm_status_string(6)


def test_sass_is_blackwell_native(native_lib):
    """The hot kernel must contain tcgen05 MMA (UTCIMMA), TMEM loads (LDTM) and TMA loads (UTMALDG)."""
    from metricsfm_b200.build import LIB_PATH
    sass = subprocess.check_output(["cuobjdump", "-sass", LIB_PATH], text=True)
    for mnemonic in ("UTCIMMA", "LDTM", "UTMALDG", "UBLKCP"):
        assert mnemonic in sass, mnemonic
    assert "sm_100a" in sass


def test_create_fails_cleanly_without_gpu(native_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from metricsfm_b200._lib import Config
    cfg = Config()
    cfg.device, cfg.max_images, cfg.arena_rows = 0, 4, 1024
    h = C.c_void_p()
    st = native_lib.msfm_create(C.byref(cfg), C.byref(h))
    assert st != 0 and not h.value          # an error code, never an abort, never a silent CPU path
    assert native_lib.msfm_create(None, C.byref(h)) == 1
    assert native_lib.msfm_destroy(None) == 0


def test_python_host_refuses_to_run_without_library(monkeypatch, tmp_path):
    from metricsfm_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "missing.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "metricsfm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle_quantize_f32", ""), f"{f} references the oracle"
