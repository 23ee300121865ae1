"""The reference-compatible C++ API (include/libbackscrub.h + include/background.h), driven the way
app/deepseg.cc's CalcMask drives the reference (bs_maskgen_new / process / delete with callbacks),
compiled against a minimal cv::Mat stand-in because this image has no OpenCV C++ headers."""
import os
import subprocess

import numpy as np
import pytest

from tests.conftest import ROOT, model_path

CPP = os.path.join(ROOT, "tests", "cpp")


@pytest.fixture(scope="module")
def shim():
    lib = os.path.join(ROOT, "backscrub_b200", "libbackscrub_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()
    subprocess.check_call(["make", "-C", CPP, "-s"])
    return os.path.join(CPP, "shim_test")


def test_shim_compiles_and_fails_loudly_without_gpu(shim, tmp_path):
    import backscrub_b200 as bs
    if bs.device_count() > 0:
        pytest.skip("a GPU is present")
    r = subprocess.run([shim, model_path("meet_lite"), "640", "480", "/dev/null", "0", str(tmp_path / "m")], capture_output=True, text=True)
    assert r.returncode == 3 and "nullptr" in r.stdout and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_shim_masks_match_oracle(shim, tmp_path):
    from oracle import pyoracle as po
    from tests import synth
    W, H, n = 640, 480, 3
    frames = np.stack([synth.frame(W, H, t=t) for t in range(n)])
    fin, fout = tmp_path / "frames.bgr", tmp_path / "masks.out"
    frames.tofile(fin)
    r = subprocess.run([shim, model_path("mlkit"), str(W), str(H), str(fin), str(n), str(fout)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "callbacks: " + "PIM" * n in r.stdout          # onprep, oninfer, onmask in order, once per frame
    assert f"grab_background rc=0 size={W}x{H}" in r.stdout
    masks = np.fromfile(fout, np.uint8).reshape(n, H, W)
    o = po.MaskGen(model_path("mlkit"), W, H)
    for t in range(n):
        assert np.array_equal(masks[t], o.process(frames[t]))
