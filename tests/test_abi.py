"""The C-ABI library loads and exports every symbol include/motionbert_b200.h declares; host-side logic that
needs no GPU.  No compute calls here."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from motionbert_b200 import _lib, build as _build_mod  # noqa: F401
from motionbert_b200 import build


def _declared(header="motionbert_b200.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mb_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_header_and_binding_agree():
    assert _declared() == sorted(_lib.EXPORTS)
    assert _declared("motionbert_b200_test.h") == sorted(_lib.TEST_EXPORTS)


def test_library_exports_every_declared_symbol(lib):
    for name in _declared():
        assert hasattr(lib, name), name


def test_product_library_carries_no_test_code_and_the_test_twin_carries_everything(lib):
    """libmotionbert_b200.so: production kernels + the ABI of include/motionbert_b200.h only.  The test twin (same source,
    -DMB_TEST_KERNELS) adds the hooks of include/motionbert_b200_test.h and the reference kernels."""
    for name in _lib.TEST_EXPORTS:
        with pytest.raises(AttributeError):
            getattr(lib, name)
    tl = _lib.load_test()
    for name in _lib.EXPORTS + _lib.TEST_EXPORTS:
        assert hasattr(tl, name), name
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if os.path.exists(cuobjdump):
        names = subprocess.run([cuobjdump, "-res-usage", _lib.LIB_PATH], capture_output=True, text=True).stdout
        for k in ("gemm_ref_kernel", "gemm_tc_kernel", "attn_t_ref_kernel", "attn_s_kernel", "attn_t2_kernel"):
            assert k not in names, k
        assert "gemm2_kernel" in names and "attn_t16_kernel" in names


def test_version_and_desc_layout(lib):
    assert lib.mb_version() == 1
    assert ctypes.sizeof(_lib.MbDesc) == 12 * 4


def test_create_fails_loudly_without_a_b200(lib):
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    d = _lib.MbDesc(3, 3, 512, 512, 5, 8, 1024, 17, 243, 1e-6, 0.0, 0)
    h = ctypes.c_void_p()
    rc = lib.mb_create(ctypes.byref(d), ctypes.byref(h))
    assert rc < 0 and not h.value
    assert len(lib.mb_last_error()) > 0
    with pytest.raises(_lib.MbError):
        _lib.check(rc, "mb_create")


def test_bad_descriptor_is_rejected(lib):
    for bad in ((3, 3, 500, 512, 5, 8, 1024, 17, 243, 1e-6, 0.0, 0),      # C % 256
                (3, 3, 512, 512, 5, 7, 1024, 17, 243, 1e-6, 0.0, 0),      # heads !| C
                (3, 3, 512, 512, 5, 8, 1024, 17, 300, 1e-6, 0.0, 0),      # maxlen > 256
                (3, 3, 512, 512, 5, 8, 1024, 17, 243, 1e-6, 0.0, 9)):     # math mode
        d = _lib.MbDesc(*bad)
        h = ctypes.c_void_p()
        assert lib.mb_create(ctypes.byref(d), ctypes.byref(h)) == -1
        assert b"" != lib.mb_last_error()
    assert lib.mb_create(None, ctypes.byref(ctypes.c_void_p())) == -2


def test_sass_contains_blackwell_instructions():
    """The shipped kernels are tcgen05 / TMA code, not legacy mma.sync (B200_PROFILING.md mnemonics)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTCQMMA" in sass and "UTMALDG" in sass and "LDTM" in sass and "STTM" in sass
    assert "HMMA." not in sass.replace("UTCHMMA", "")


def test_kernel_flag_constants_match_the_header():
    """MB_FLAG_* / MB_MATH_* of the ctypes binding are the header's values (the flags are A/B switches: a drifted constant
    would silently select another kernel path)."""
    src = open(os.path.join(ROOT, "include", "motionbert_b200.h")).read()
    flags = {k: int(v, 16) for k, v in re.findall(r"#define\s+(MB_FLAG_[A-Z0-9_]+)\s+(0x[0-9a-fA-F]+)u", src)}
    assert len(flags) >= 8 and len(set(flags.values())) == len(flags)          # distinct bits
    for k, v in flags.items():
        assert getattr(_lib, k) == v, k
        assert v & (v - 1) == 0, k                                              # single bit each
    maths = {k: int(v) for k, v in re.findall(r"(MB_MATH_[A-Z0-9]+)\s*=\s*(\d+)", src)}
    assert maths == {"MB_MATH_BF16X3": 0, "MB_MATH_BF16": 1, "MB_MATH_F16C": 2}
    for k, v in maths.items():
        assert getattr(_lib, k) == v


def test_product_library_contains_the_fused_mlp_kernel_with_blackwell_instructions():
    """mlp_fused_kernel ships in the product library and is a tcgen05 / TMA kernel: UTCHMMA + UTCQMMA (the F16C pass pair),
    LDTM, UTMALDG and UTMASTG in its SASS, no legacy HMMA."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    build.build()
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    i = sass.find("Function : _ZN2mb16mlp_fused_kernel")
    assert i >= 0
    j = sass.find("Function : ", i + 20)
    body = sass[i:j if j > 0 else len(sass)]
    for mnem in ("UTCHMMA.2CTA", "UTCQMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR"):
        assert mnem in body, mnem
    assert "HMMA.16816" not in body and " HMMA." not in body
    assert body.count("MUFU.EX2") >= 100 and body.count("MUFU.RCP") < 30       # one-MUFU GELU (RCP only in the LN statistics)
