"""The C-ABI shared library loads and exports every symbol include/backscrub_b200.h
declares; on a GPU-less box every compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests.conftest import ROOT, model_path


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "backscrub_b200.h")).read()
    return sorted(set(re.findall(r"BSB_API[^;(]*?\b(bsb_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def product():
    import backscrub_b200 as bs
    if not os.path.exists(bs.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return bs


def test_header_symbols_match_binding_list():
    from backscrub_b200 import _binding
    assert _declared_symbols() == sorted(_binding.SYMBOLS)


def test_library_exports_every_declared_symbol(product):
    lib = product.lib()
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.bsb_version()


def test_library_is_cuda_code_for_sm_100a(product):
    """The product .so carries sm_100a SASS (cuobjdump lists the ELF) — i.e. it is the CUDA build."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", product.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_gpu(product):
    lib = product.lib()
    if lib.bsb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(product.BackscrubError):
        product.bs_maskgen_new(model_path("mlkit"), 2, 640, 480)
    a = np.zeros((4, 4, 3), np.uint8)
    with pytest.raises(product.BackscrubError):
        product.alpha_blend(a, a, np.zeros((4, 4), np.uint8))
    # raw C ABI: NULL + message, like the reference's nullptr + ondebug contract
    msgs = []
    cb = product._binding.DEBUG_CB(lambda _c, m: msgs.append(m))
    null_stage = C.cast(None, product._binding.STAGE_CB)
    h = lib.bsb_maskgen_new(model_path("mlkit").encode(), 2, 640, 480, cb, null_stage, null_stage, null_stage, None)
    assert not h and msgs and b"CUDA" in msgs[0]
    assert lib.bsb_alpha_blend(0, None, None, None, None, 0) == 0


def test_product_package_never_imports_oracle_or_emulator():
    pkg = os.path.join(ROOT, "backscrub_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle/oracle_", "").replace("the oracle", "").replace("CPU oracle", "") \
                    or "import" not in text or f in ("bsb_common.h",), f
                assert "pyoracle" not in text and "liboracle" not in text, f
                assert "libbsb_emu" not in text, f
