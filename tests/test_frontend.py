"""Host-side callers of the hot path through the C ABI, on the kernel-logic emulator build
(CPU) and on the real library (`-m gpu`):
  * bsb_calcmask_*   — app/deepseg.cc:159-286 (worker thread, latest frame wins, a mask is handed out once)
  * bsb_background_* — app/background.cc:13-202 (still / paced looping video provider, frame numbers, GPU resize)."""
import ctypes as C
import time

import numpy as np
import pytest

from backscrub_b200 import _binding as B
from tests import synth
from tests.conftest import model_path


def _calcmask_checks(lib, key, W, H):
    from oracle import pyoracle as po
    cm = lib.bsb_calcmask_new(model_path(key).encode(), 2, W, H, 0)
    assert cm, lib.bsb_last_error()
    try:
        o = po.MaskGen(model_path(key), W, H)
        out = np.full((H, W), 7, np.uint8)
        assert lib.bsb_calcmask_get_output_mask(cm, out.ctypes.data_as(C.c_void_p), W) == 0      # nothing yet: untouched
        assert (out == 7).all()
        frames = [synth.frame(W, H, t=t) for t in range(3)]
        for t, fr in enumerate(frames):          # synchronous use: every frame is processed, masks follow the oracle
            assert lib.bsb_calcmask_set_input_frame(cm, fr.ctypes.data_as(C.c_void_p), W * 3) == 1
            t0 = time.time()
            while lib.bsb_calcmask_frames_done(cm) < t + 1:
                assert time.time() - t0 < 600
                time.sleep(0.005)
            assert lib.bsb_calcmask_get_output_mask(cm, out.ctypes.data_as(C.c_void_p), W) == 1
            assert np.array_equal(out, o.process(fr)), f"frame {t}"
            assert lib.bsb_calcmask_mask_serial(cm) == t + 1
            keep = out.copy()
            assert lib.bsb_calcmask_get_output_mask(cm, out.ctypes.data_as(C.c_void_p), W) == 0  # handed out once
            assert np.array_equal(out, keep)
        ns = (C.c_long * 5)()
        assert lib.bsb_calcmask_timings(cm, ns) == 1
        assert ns[1] > 0 and ns[2] > 0 and ns[3] > 0 and ns[4] >= ns[1] + ns[2] + ns[3]   # prep, infer, mask inside the loop time
        # burst: frames arrive faster than the worker: the pending frame is replaced (latest frame wins), nothing queues up
        done0 = lib.bsb_calcmask_frames_done(cm)
        burst = [synth.frame(W, H, t=10 + t) for t in range(4)]
        for fr in burst:
            lib.bsb_calcmask_set_input_frame(cm, fr.ctypes.data_as(C.c_void_p), W * 3)
        t0 = time.time()
        while lib.bsb_calcmask_mask_serial(cm) < 3 + len(burst):
            assert time.time() - t0 < 600
            time.sleep(0.005)
        time.sleep(0.05)
        done = lib.bsb_calcmask_frames_done(cm) - done0
        assert 1 <= done <= len(burst)
        assert lib.bsb_calcmask_mask_serial(cm) == 3 + len(burst)          # the newest mask belongs to the last frame set
    finally:
        lib.bsb_calcmask_delete(cm)
    lib.bsb_calcmask_delete(None)
    assert not lib.bsb_calcmask_new(b"/nonexistent/segm_x.tflite", 2, W, H, 0)


def _provider_checks(lib):
    from oracle import pyoracle as po
    img = np.ascontiguousarray(synth.background()[:180, :320])
    bg = lib.bsb_background_new_still(0, img.ctypes.data_as(C.c_void_p), 320, 180, 320 * 3, 0)
    assert bg
    out = np.zeros((120, 200, 3), np.uint8)
    assert lib.bsb_background_grab(bg, 200, 120, out.ctypes.data_as(C.c_void_p), 600) == 1
    assert np.array_equal(out, po.resize_linear_u8(img, 200, 120))
    half = np.zeros((90, 160, 3), np.uint8)                                 # exact 2x down-scale: cv::resize's INTER_AREA path
    assert lib.bsb_background_grab(bg, 160, 90, half.ctypes.data_as(C.c_void_p), 480) == 1
    assert np.array_equal(half, po.resize_linear_u8(img, 160, 90))
    assert lib.bsb_background_frame(bg) == 1 and lib.bsb_background_running(bg) == 0
    assert lib.bsb_background_grab(bg, 0, 120, out.ctypes.data_as(C.c_void_p), 600) == -1
    lib.bsb_background_delete(bg)

    # video: a python frame source behind the two callbacks
    vid = np.stack([np.roll(img, 16 * i, axis=1) for i in range(5)])
    state = {"pos": 0, "reads": 0, "rewinds": 0}

    def read(_u, data, w, h, pitch):
        if state["pos"] >= len(vid):
            return 0
        data[0] = vid[state["pos"]].ctypes.data
        w[0], h[0], pitch[0] = 320, 180, 960
        state["pos"] += 1; state["reads"] += 1
        return 1

    def rewind(_u):
        state["pos"] = 0; state["rewinds"] += 1
        return 1
    rcb, wcb = B.BG_READ_CB(read), B.BG_REWIND_CB(rewind)
    t0 = time.time()
    bg = lib.bsb_background_new_video(0, 50.0, 0, rcb, wcb, None, vid[0].ctypes.data_as(C.c_void_p), 320, 180, 960, 0)
    assert bg
    seen = []
    while time.time() - t0 < 0.6:
        n = lib.bsb_background_grab(bg, 200, 120, out.ctypes.data_as(C.c_void_p), 600)
        assert 0 <= n <= 5
        seen.append(n)
        # whatever frame is current, the grab is that decoded frame resized
        assert any(np.array_equal(out, po.resize_linear_u8(v, 200, 120)) for v in vid)
        time.sleep(0.01)
    elapsed = time.time() - t0
    lib.bsb_background_delete(bg)                    # joins the reader
    assert state["rewinds"] >= 1                     # 5 frames at 50 fps: the loop wrapped several times
    # paced to 50 fps: about 30 frames in 0.6 s (far fewer than an unpaced reader would pull)
    assert 0.5 * 50 * elapsed <= state["reads"] <= 1.5 * 50 * elapsed + 3, (state, elapsed)
    assert len(set(seen)) >= 3


def test_calcmask_on_the_emulator():
    from tests.emu.emu_lib import emu
    _calcmask_checks(emu(), "meet_lite", 320, 240)


def test_background_provider_on_the_emulator():
    from tests.emu.emu_lib import emu
    _provider_checks(emu())


@pytest.mark.gpu
def test_calcmask_gpu():
    import backscrub_b200 as bs
    _calcmask_checks(bs.lib(), "mlkit", 640, 480)


@pytest.mark.gpu
def test_background_provider_gpu():
    import backscrub_b200 as bs
    _provider_checks(bs.lib())
