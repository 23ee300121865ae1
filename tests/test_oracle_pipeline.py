"""Oracle pipeline (bs_maskgen_process + composite) vs a cv2-composed restatement of
lib/libbackscrub.cc:279-376 / app/deepseg.cc:87-134 around the oracle's own CNN output,
geometry for every BASELINE config, IIR closed form, and committed regression goldens."""
import os

import cv2
import numpy as np
import pytest

from oracle import pyoracle as po
from tests import synth
from tests.conftest import ROOT, model_path

GEOM = [  # SURVEY.md §8 geometry table: (model, W, H, roidim, in_roidim)
    ("mlkit", 640, 480, [80, 0, 480, 480], [0, 0, 256, 256]),
    ("deeplab", 1280, 720, [280, 0, 720, 720], [0, 0, 257, 257]),
    ("meet_full", 1280, 720, [0, 0, 1280, 720], [0, 0, 256, 144]),
    ("meet_full", 640, 480, [0, 0, 640, 480], [32, 0, 192, 144]),
    ("bodypix", 3840, 2160, [840, 0, 2160, 2160], [0, 0, 257, 257]),
    ("meet_lite", 640, 480, [0, 0, 640, 480], [16, 0, 128, 96]),
]


@pytest.mark.parametrize("key,W,H,roi,in_roi", GEOM)
def test_geometry(key, W, H, roi, in_roi):
    g = po.MaskGen(model_path(key), W, H)
    assert g.roidim == roi and g.in_roidim == in_roi
    if key == "bodypix":
        assert g.out_roidim == [0, 0, 33, 33]   # documented deviation (reference throws)
    else:
        assert g.out_roidim == in_roi


def test_model_type_from_name():
    L = po.lib()
    assert L.or_model_type_from_name(b"/x/body-pix-float.tflite") == po.MODEL_BODYPIX
    assert L.or_model_type_from_name(b"deeplabv3_257.tflite") == po.MODEL_DEEPLAB
    assert L.or_model_type_from_name(b"segm_full_v679.tflite") == po.MODEL_MEET
    assert L.or_model_type_from_name(b"selfiesegmentation.tflite") == po.MODEL_MLKIT
    assert L.or_model_type_from_name(b"unknown.tflite") == po.MODEL_UNKNOWN
    with pytest.raises(RuntimeError):
        po.MaskGen(os.path.join(ROOT, "tests", "conftest.py"), 640, 480)


def _cv_pipeline(g, key, frame, bg_raw, ofinal_prev):
    """cv2 restatement of process() around the oracle's model output."""
    x, y, w, h = g.roidim
    ix, iy, iw, ih = g.in_roidim
    mh, mw, _ = g.in_hwc
    in_bgr = np.zeros((mh, mw, 3), np.uint8)
    in_bgr[iy:iy + ih, ix:ix + iw] = cv2.resize(frame[y:y + h, x:x + w], (iw, ih))
    in_rgb = cv2.cvtColor(in_bgr, cv2.COLOR_BGR2RGB)
    assert np.array_equal(g.in_u8, in_rgb)
    filt = cv2.bilateralFilter(in_rgb, 5, 100.0, 100.0)
    d = np.abs(filt.astype(int) - g.filtered_u8.astype(int))
    assert d.max() <= 1 and (d != 0).mean() <= 1e-4
    out = g.output_f32
    if key == "deeplab":
        val = np.where(out.argmax(-1) == 15, 0, 255)
    elif key.startswith("meet"):
        val = np.where(out[..., 0] < out[..., 1], 0, 255)
    else:
        val = np.where(out[..., 0] > np.float32(0.65), 0, 255)
    ofinal = ((val & 0xE0) | (ofinal_prev >> 3)).astype(np.uint8)
    assert np.array_equal(g.ofinal, ofinal)
    ox, oy, ow, oh = g.out_roidim
    up = cv2.resize(ofinal[oy:oy + oh, ox:ox + ow], (w, h))
    mask = np.full((g.H, g.W), 255, np.uint8)
    mask[y:y + h, x:x + w] = cv2.blur(up, (5, 5))
    bg = cv2.resize(bg_raw, (g.W, g.H))
    m = mask.astype(int)[..., None]
    blend = ((bg.astype(int) * m + frame.astype(int) * (255 - m)) // 255).astype(np.uint8)
    return mask, blend, ofinal


@pytest.mark.parametrize("key,W,H", [("mlkit", 640, 480), ("meet_full", 640, 480), ("meet_full", 1280, 720),
                                     ("meet_lite", 640, 480), ("bodypix", 640, 480)])
def test_pipeline_vs_cv2_composition(key, W, H):
    g = po.MaskGen(model_path(key), W, H)
    bg_raw = synth.background()
    ofinal = np.zeros(g.out_hwc[:2], np.uint8)
    for t in range(3):
        fr = synth.frame(W, H, t=t)
        out, yuyv, mask = g.composite(fr, bg_raw)
        mask_cv, blend_cv, ofinal = _cv_pipeline(g, key, fr, bg_raw, ofinal)
        assert np.array_equal(mask, mask_cv)
        assert np.array_equal(out, blend_cv)
        assert np.array_equal(yuyv, po.convert_rgb_to_yuyv(out))
    assert 0.05 < (mask < 128).mean() < 0.6   # a person is found


@pytest.mark.slow
def test_pipeline_deeplab_720p():
    g = po.MaskGen(model_path("deeplab"), 1280, 720)
    fr = synth.frame(1280, 720, t=0)
    out, yuyv, mask = g.composite(fr, synth.background())
    mask_cv, blend_cv, _ = _cv_pipeline(g, "deeplab", fr, synth.background(), np.zeros((257, 257), np.uint8))
    assert np.array_equal(mask, mask_cv) and np.array_equal(out, blend_cv)
    assert (mask[:, :280] == 255).all() and (mask[:, 1000:] == 255).all()   # outside roidim stays background


def test_iir_closed_form():
    """after >= 3 frames ofinal == 0xE0*b_t | 0x1C*b_{t-1} | 0x03*b_{t-2}."""
    g = po.MaskGen(model_path("mlkit"), 640, 480)
    rng = np.random.default_rng(5)
    hist = []
    for t in range(5):
        outp = rng.uniform(0, 1, (256, 256, 1)).astype(np.float32)
        g.post_from_output(outp)
        hist.append((outp[..., 0] <= np.float32(0.65)).astype(np.uint8))   # b = 1 <=> background
        if t >= 2:
            exp = 0xE0 * hist[-1] | 0x1C * hist[-2] | 0x03 * hist[-3]
            assert np.array_equal(g.ofinal, exp)


def test_meet_decision_nan_semantics():
    """lib/libbackscrub.cc:350-356: overflowing expf -> NaN -> comparison false -> 255."""
    g = po.MaskGen(model_path("meet_full"), 640, 480)
    outp = np.zeros((144, 256, 2), np.float32)
    outp[0, 0] = (1.0, 2.0)        # person
    outp[0, 1] = (2.0, 1.0)        # background
    outp[0, 2] = (10.0, 100.0)     # exp overflow -> inf/inf = NaN -> background
    outp[0, 3] = (-200.0, -100.0)  # both underflow to 0 -> 0/0 = NaN -> background
    g.post_from_output(outp)
    assert list(g.ofinal[0, :4]) == [0, 0xE0, 0xE0, 0xE0]


def test_regression_goldens():
    """Committed oracle outputs (tests/golden/make_golden.py) — guards the oracle itself
    against drift; NOT a reference pin (whole-model parity is unpinned, see oracle.h)."""
    path = os.path.join(ROOT, "tests", "golden", "pipeline_golden.npz")
    if not os.path.exists(path):
        pytest.skip("golden file not generated")
    z = np.load(path)
    for key in ("mlkit", "meet_full"):
        g = po.MaskGen(model_path(key), 640, 480)
        for t in range(int(z["n_frames"])):
            out, yuyv, mask = g.composite(synth.frame(640, 480, t=t), synth.background())
        assert np.array_equal(np.packbits(mask < 128), z[f"{key}_maskbits"])
        assert int(out.astype(np.int64).sum()) == int(z[f"{key}_outsum"])
        assert int(yuyv.astype(np.int64).sum()) == int(z[f"{key}_yuyvsum"])
