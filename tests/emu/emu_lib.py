"""Loads tests/emu/libbsb_emu.so — the CUDA sources compiled against the kernel-logic
emulator (cuemu.h).  TEST-ONLY: lets `-m "not gpu"` tests drive the planner, the C ABI and
every kernel's indexing on a GPU-less box.  Never used by the product, bench or smoke."""
import os
import subprocess

from backscrub_b200 import _binding

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def emu():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
        _LIB = _binding.bind(os.path.join(_HERE, "libbsb_emu.so"))
    return _LIB
