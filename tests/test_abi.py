"""CPU: the C-ABI library builds, loads and exports every symbol include/sustaingym_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "sustaingym_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(sg_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported(native_lib):
    names = declared_symbols()
    assert len(names) >= 10
    for name in names:
        assert hasattr(native_lib, name), f"{name} declared in the header but not exported"


def test_ctypes_signatures_cover_header(native_lib):
    from sustaingym_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_abi_version_and_error_string(native_lib):
    assert native_lib.sg_abi_version() == 2
    assert isinstance(native_lib.sg_last_error(), bytes)


def test_struct_layout_matches_header(native_lib):
    """sizeof of the ctypes mirrors == what the C compiler lays out (checked with a gcc probe)."""
    import subprocess
    import tempfile
    from sustaingym_b200 import _lib
    src = '#include <stdio.h>\n#include "sustaingym_b200.h"\nint main(){printf("%zu %zu %zu %zu\\n",' \
          'sizeof(sg_building_model),sizeof(sg_ev_model),sizeof(sg_ev_state),sizeof(sg_ev_session));' \
          'printf("%zu %zu %zu\\n",sizeof(sg_ev_gmm),sizeof(sg_cogen_model),(size_t)SG_COGEN_SCRATCH_FLOATS(10));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "p.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "p")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes[0] == ctypes.sizeof(_lib.BuildingModel)
    assert sizes[1] == ctypes.sizeof(_lib.EVModel)
    assert sizes[2] == ctypes.sizeof(_lib.EVState)
    assert sizes[3] == 16
    assert sizes[4] == ctypes.sizeof(_lib.EVGmm)
    assert sizes[5] == ctypes.sizeof(_lib.CogenModel)
    assert sizes[6] == _lib.cogen_scratch_floats(10)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from sustaingym_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    import pytest
    with pytest.raises(_lib.NativeLibraryError):
        _lib.load()


def test_argument_checks_need_no_gpu(native_lib):
    """Entry points validate their arguments before touching CUDA: rc != 0 and a message, no launch."""
    from sustaingym_b200 import _lib
    rc = native_lib.sg_ev_sample_gmm_days(None, 4, 0, 1, 100.0, 0.5, 8, 128, None, None, None, None, None)
    assert rc == 1 and b"NULL model" in native_lib.sg_last_error()
    g = _lib.EVGmm(30, 0, 54, 0, None, None, None, None, None)
    rc = native_lib.sg_ev_sample_gmm_days(ctypes.byref(g), 4, 0, 1, 100.0, 0.5, 8, 128, None, None, None, None, None)
    assert rc == 1 and b"session-count" in native_lib.sg_last_error()
    rc = native_lib.sg_tc_gemm_selftest(None, None, None, 100, 8, None)
    assert rc != 0
