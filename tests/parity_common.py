"""Shared parity checks: the CUDA path (through the C ABI) vs the CPU oracle, stage by
stage and end to end.  Used with the product library on a GPU (`-m gpu`) and with the
kernel-logic emulator build on the GPU-less box (`-m "not gpu"`)."""
import numpy as np

from backscrub_b200 import api
from oracle import pyoracle as po
from tests import synth
from tests.conftest import model_path


def check_pipeline(lib, key, W, H, n_frames=3, batch=None, tensors=False, frame_kind="person", flags=0):
    """n consecutive frames of one stream through bsb_composite (batched) vs the oracle."""
    batch = batch or n_frames
    g = api.MaskGen(lib, model_path(key), W, H, max_batch=batch, flags=flags | (1 if tensors else 0))
    o = po.MaskGen(model_path(key), W, H)
    assert g.roidim == o.roidim and g.in_roidim == o.in_roidim and g.out_roidim == o.out_roidim
    assert g.in_hwc == o.in_hwc and g.out_hwc == o.out_hwc
    bg = synth.background()
    g.set_background(bg)
    assert np.array_equal(g.background(), po.resize_linear_u8(bg, W, H))
    frames = np.stack([synth.frame(W, H, t=t, kind=frame_kind) for t in range(n_frames)])
    person = 0.0
    for s in range(0, n_frames, batch):
        chunk = frames[s:s + batch]
        out, yuyv, mask = g.composite(chunk)
        for b in range(chunk.shape[0]):
            ro, ry, rm = o.composite(chunk[b], bg)
            assert np.array_equal(g.stage_u8(0, b), o.in_u8), f"{key}: resized RGB input differs (frame {s + b})"
            if tensors:
                assert np.array_equal(g.stage_u8(1, b), o.filtered_u8), f"{key}: bilateral output differs"
            assert np.array_equal(g.stage_u8(2, b), o.ofinal), f"{key}: decision/IIR state differs (frame {s + b})"
            assert np.array_equal(mask[b], rm), f"{key}: mask differs (frame {s + b})"
            assert np.array_equal(out[b], ro), f"{key}: composite differs (frame {s + b})"
            assert np.array_equal(yuyv[b], ry), f"{key}: YUYV differs (frame {s + b})"
            person = float((rm < 128).mean())
    g.close()
    return person


def check_tensors(lib, key, W=640, H=480):
    """Every materialised activation of the CUDA graph executor vs the oracle interpreter, bit for bit."""
    g = api.MaskGen(lib, model_path(key), W, H, max_batch=1, flags=1)
    o = po.MaskGen(model_path(key), W, H)
    fr = synth.frame(W, H, t=1)
    g.set_background(synth.background())
    g.composite(fr)
    o.process(fr)
    m = po.Model(model_path(key))
    m.invoke(o.input_f32)
    compared = 0
    for t in range(m.n_tensors):
        if m.is_const(t):
            continue
        try:
            ref = m.tensor(t)
        except KeyError:
            continue
        got = g.tensor(t, ref.size)
        if got is None:
            continue      # folded into a neighbouring kernel by the planner
        compared += 1
        assert np.array_equal(got.view(np.uint32), ref.ravel().view(np.uint32)), \
            f"{key}: tensor {t} {m.shape(t)} differs, max |d| = {np.abs(got - ref.ravel()).max()}"
    g.close()
    return compared


def check_infer_batch(lib, key, n=3):
    """bsb_infer on a batch == oracle interpreter frame by frame (bit-exact)."""
    g = api.MaskGen(lib, model_path(key), 640, 480, max_batch=n)
    m = po.Model(model_path(key))
    rng = np.random.default_rng(11)
    x = rng.uniform(-1 if key == "deeplab" else 0, 1, (n, *g.in_hwc)).astype(np.float32)
    out = g.infer(x)
    for b in range(n):
        ref = m.invoke(x[b])[0]
        assert np.array_equal(out[b].view(np.uint32), ref.view(np.uint32)), f"{key}: frame {b} differs"
    g.close()


def check_stage_functions(lib):
    rng = np.random.default_rng(3)
    for (w, h) in [(640, 480), (34, 6), (2, 2)]:
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        b = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        m = rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert np.array_equal(api.alpha_blend(lib, a, b, m), po.alpha_blend(a, b, m))
        assert np.array_equal(api.convert_rgb_to_yuyv(lib, a), po.convert_rgb_to_yuyv(a))
    for (sw, sh, dw, dh) in [(1280, 720, 640, 480), (100, 37, 313, 211), (64, 64, 32, 32), (640, 480, 640, 480), (960, 540, 1280, 720)]:
        s = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        assert np.array_equal(api.resize_u8c3(lib, s, dw, dh), po.resize_linear_u8(s, dw, dh)), (sw, sh, dw, dh)


def check_yuyv_ingest(lib, key="meet_lite", W=640, H=480, n=2):
    """Camera-format path: YUYV frames in -> (GPU YUYV->BGR) -> pipeline, vs oracle conversion + oracle pipeline."""
    import cv2
    g = api.MaskGen(lib, model_path(key), W, H, max_batch=n)
    o = po.MaskGen(model_path(key), W, H)
    bg = synth.background()
    g.set_background(bg)
    bgr = np.stack([synth.frame(W, H, t=t) for t in range(n)])
    yuyv_in = np.stack([po.convert_rgb_to_yuyv(f) for f in bgr])          # any valid YUYV camera frame will do
    rng = np.random.default_rng(9)
    yuyv_in[0, :8] = rng.integers(0, 256, (8, W, 2), dtype=np.uint8)       # include out-of-gamut / extreme codes
    for b in range(n):
        assert np.array_equal(api.convert_yuyv_to_bgr(lib, yuyv_in[b]), po.yuyv_to_bgr(yuyv_in[b]))
    out, yuyv, mask = g.composite_yuyv(yuyv_in)
    for b in range(n):
        fr = po.yuyv_to_bgr(yuyv_in[b])
        ro, ry, rm = o.composite(fr, bg)
        assert np.array_equal(mask[b], rm) and np.array_equal(out[b], ro) and np.array_equal(yuyv[b], ry), b
    g.close()


def check_mask_only_and_callbacks(lib, key="mlkit", W=640, H=480):
    """bs_maskgen_process semantics: callbacks fire in order once per call; mask aliases
    context storage; consecutive calls advance the IIR like the oracle."""
    events = []
    g = api.MaskGen(lib, model_path(key), W, H, onprep=lambda: events.append("prep"),
                    oninfer=lambda: events.append("infer"), onmask=lambda: events.append("mask"))
    o = po.MaskGen(model_path(key), W, H)
    for t in range(3):
        fr = synth.frame(W, H, t=t)
        assert np.array_equal(g.process(fr), o.process(fr))
    assert events == ["prep", "infer", "mask"] * 3
    g.reset_state()
    o2 = po.MaskGen(model_path(key), W, H)
    fr = synth.frame(W, H, t=7)
    assert np.array_equal(g.process(fr), o2.process(fr))
    g.close()


def check_errors(lib, tmp_path):
    import pytest
    with pytest.raises(api.BackscrubError, match="unknown model type"):
        p = tmp_path / "mystery.tflite"
        p.write_bytes(open(model_path("mlkit"), "rb").read())
        api.MaskGen(lib, str(p), 640, 480)
    with pytest.raises(api.BackscrubError, match="unable to load model"):
        api.MaskGen(lib, str(tmp_path / "missing_selfie.tflite"), 640, 480)
    with pytest.raises(api.BackscrubError, match="malformed|too short"):
        p = tmp_path / "bad_selfie.tflite"
        p.write_bytes(b"\x10\0\0\0TFL3" + b"\xff" * 64)
        api.MaskGen(lib, str(p), 640, 480)
    g = api.MaskGen(lib, model_path("mlkit"), 640, 480, max_batch=1)
    with pytest.raises(api.BackscrubError):
        g.composite(np.zeros((2, 480, 640, 3), np.uint8))          # n > max_batch
    with pytest.raises(api.BackscrubError, match="background"):
        g.composite(np.zeros((480, 640, 3), np.uint8))             # no background set
    with pytest.raises(api.BackscrubError):
        g.process(np.zeros((10, 10, 3), np.uint8))
    g.close()
