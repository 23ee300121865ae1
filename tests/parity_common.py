"""Shared parity checks: the CUDA path (through the C ABI) vs the CPU oracle, stage by
stage and end to end.  Used with the product library on a GPU (`-m gpu`) and with the
kernel-logic emulator build on the GPU-less box (`-m "not gpu"`)."""
import numpy as np

from backscrub_b200 import api
from oracle import pyoracle as po
from tests import synth
from tests.conftest import model_path


FLAG_EXACT = 16


def exact_flag(key):
    """DeepLab / BodyPix run their 1x1 convs on the tensor cores by default (3xTF32: decisions agree with the oracle on
    the committed fixtures, activations to ~1e-6); the bit-for-bit comparisons against the oracle select the exact
    fp32 path with BSB_FLAG_EXACT.  The default path has its own checks (check_tc_default)."""
    return FLAG_EXACT if key in ("deeplab", "bodypix") else 0


def check_pipeline(lib, key, W, H, n_frames=3, batch=None, tensors=False, frame_kind="person", flags=0):
    """n consecutive frames of one stream through bsb_composite (batched) vs the oracle."""
    batch = batch or n_frames
    flags |= exact_flag(key)
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
    g = api.MaskGen(lib, model_path(key), 640, 480, max_batch=n, flags=exact_flag(key))
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


def check_overlapped_host_call(lib, key="meet_lite", W=640, H=480, n=19, oracle_frames=(0, 1, 8, 9, 18)):
    """bsb_composite_yuyv with enough frames to take the chunked, copy/compute-overlapped schedule (chunks of 8 + 8 + 3 on
    three streams) must return exactly what the serial schedule (e2e_chunk = 0) returns — the temporal smoother's state
    carries across the chunk borders — over two consecutive calls, and both must equal the oracle."""
    bg = synth.background()
    bgr = np.stack([synth.frame(W, H, t=t) for t in range(n)])
    yuyv_in = np.stack([po.convert_rgb_to_yuyv(f) for f in bgr])
    res = {}
    try:
        for chunk in (0, 8):
            assert lib.bsb_set_tuning(b"e2e_chunk", chunk)
            g = api.MaskGen(lib, model_path(key), W, H, max_batch=n)
            g.set_background(bg)
            first = g.composite_yuyv(yuyv_in)
            second = g.composite_yuyv(yuyv_in[::-1].copy())        # state carried from the first call
            res[chunk] = (first, second)
            g.close()
    finally:
        lib.bsb_set_tuning(b"e2e_chunk", 16)
    for call in range(2):
        for a, b, what in zip(res[0][call], res[8][call], ("composite", "YUYV", "mask")):
            assert np.array_equal(a, b), f"{key}: overlapped schedule changes the {what} (call {call})"
    o = po.MaskGen(model_path(key), W, H)
    out, yuyv, mask = res[8][0]
    for b in range(n):
        ro, ry, rm = o.composite(po.yuyv_to_bgr(yuyv_in[b]), bg)
        if b in oracle_frames:
            assert np.array_equal(mask[b], rm) and np.array_equal(out[b], ro) and np.array_equal(yuyv[b], ry), b


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
    with pytest.raises(api.BackscrubError, match="must be odd"):
        g.set_bgblur(24)                                           # app/deepseg.cc:423-426
    with pytest.raises(api.BackscrubError, match="invalid background"):
        g.set_background_ring(np.zeros((0, 4, 4, 3), np.uint8))
    with pytest.raises(api.BackscrubError):
        g.process(np.zeros((10, 10, 3), np.uint8))
    g.close()


def check_app_stage_functions(lib):
    """Gaussian taps for every legal strength; blur / flip stage kernels vs the oracle (pinned on cv2)."""
    for k in range(1, 256, 2):
        assert np.array_equal(api.gaussian_taps(lib, k), po.gaussian_kernel_q8(k)), k
    rng = np.random.default_rng(5)
    for (w, h, k) in [(640, 480, 25), (300, 40, 3), (37, 21, 9), (70, 50, 51), (33, 9, 101), (16, 16, 1),
                      (100, 70, 31), (65, 33, 27), (10, 8, 25), (130, 64, 13), (64, 32, 5), (4, 3, 29)]:
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(api.gaussian_blur(lib, a, k), po.gaussian_blur(a, k)), (w, h, k)
    for k in range(3, 42, 2):                                    # every fused-kernel strength (3..31) and the first generic ones
        w, h = int(rng.integers(3, 150)), int(rng.integers(3, 80))
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(api.gaussian_blur(lib, a, k), po.gaussian_blur(a, k)), (w, h, k)
    a = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    for fh, fv in [(1, 0), (0, 1), (1, 1), (0, 0)]:
        assert np.array_equal(api.flip(lib, a, fh, fv), po.flip(a, fh, fv))


def check_app_options(lib, key="meet_lite", W=640, H=480):
    """The main loop's options around the blend (app/deepseg.cc:649-681): default green background, bgblur of the
    grabbed background and of the camera frame, flip, virtual-camera resize, animated background ring."""
    bg = synth.background()
    frames = np.stack([synth.frame(W, H, t=t) for t in range(4)])

    def run(setup, ref_kw, bg_raw=bg, n=2, batch=2):
        g = api.MaskGen(lib, model_path(key), W, H, max_batch=batch)
        o = po.MaskGen(model_path(key), W, H)
        if bg_raw is not None:
            g.set_background(bg_raw)
        setup(g)
        for s in range(0, n, batch):
            out, yuyv, mask = g.composite(frames[s:s + batch])
            for b in range(out.shape[0]):
                ro, ry, rm = o.composite_ex(frames[s + b], bg_raw, **ref_kw)
                assert np.array_equal(mask[b], rm), (ref_kw, s + b)
                assert out[b].shape == ro.shape and np.array_equal(out[b], ro), (ref_kw, s + b)
                assert np.array_equal(yuyv[b], ry), (ref_kw, s + b)
        g.close()

    # no background, no blur: plain green (app/deepseg.cc:603)
    g = api.MaskGen(lib, model_path(key), W, H)
    o = po.MaskGen(model_path(key), W, H)
    green = np.zeros((H, W, 3), np.uint8); green[..., 1] = 255
    assert np.array_equal(g.background(), green)
    out, yuyv, mask = g.composite(frames[0])
    ro, ry, rm = o.composite(frames[0], green)
    assert np.array_equal(out, ro) and np.array_equal(yuyv, ry) and np.array_equal(mask, rm)
    g.close()

    run(lambda g: g.set_bgblur(25), dict(bgblur=25))                                   # blurred still background
    run(lambda g: g.set_bgblur(25), dict(bgblur=25), bg_raw=None)                      # blurred camera frame
    run(lambda g: g.set_bgblur(7), dict(bgblur=7), bg_raw=None, n=3, batch=1)
    run(lambda g: g.set_output(flip_h=True), dict(flip_h=True))
    run(lambda g: g.set_output(flip_v=True, out_size=(320, 240)), dict(flip_v=True, out_size=(320, 240)))   # 2x down: INTER_AREA path
    run(lambda g: g.set_output(True, True, (854, 480)), dict(flip_h=True, flip_v=True, out_size=(854, 480)))
    run(lambda g: (g.set_bgblur(11), g.set_output(out_size=(1280, 720))), dict(bgblur=11, out_size=(1280, 720)))

    # options can be switched on a live context (graphs are re-captured) and switched off again
    g = api.MaskGen(lib, model_path(key), W, H, max_batch=1)
    o = po.MaskGen(model_path(key), W, H)
    g.set_background(bg)
    for t, (k, fh) in enumerate([(0, False), (5, True), (0, False), (25, False)]):
        g.set_bgblur(k); g.set_output(flip_h=fh)
        out, yuyv, mask = g.composite(frames[t])
        ro, ry, rm = o.composite_ex(frames[t], bg, bgblur=k, flip_h=fh)
        assert np.array_equal(out, ro) and np.array_equal(yuyv, ry) and np.array_equal(mask, rm), t
    g.close()

    # animated background: ring of 3 images, one step per camera frame, batches of 2 -> wraps inside a batch
    ring = np.stack([bg, bg[::-1].copy(), np.roll(bg, 100, axis=1)])
    g = api.MaskGen(lib, model_path(key), W, H, max_batch=2)
    o = po.MaskGen(model_path(key), W, H)
    g.set_background_ring(ring, advance=1)
    g.set_bgblur(9)
    t = 0
    for s in (0, 2):
        out, yuyv, mask = g.composite(frames[s:s + 2])
        for b in range(2):
            ro, ry, rm = o.composite_ex(frames[s + b], ring[t % 3], bgblur=9)
            assert np.array_equal(out[b], ro) and np.array_equal(yuyv[b], ry) and np.array_equal(mask[b], rm), (s, b)
            t += 1
    g.set_background_cursor(2)                                   # caller-paced: pick an image explicitly
    out, _, _ = g.composite(frames[:1])
    ro, _, _ = o.composite_ex(frames[0], ring[2], bgblur=9)
    assert np.array_equal(out[0], ro)
    g.close()


def check_pointwise_variants(lib):
    """Every exact FFMA pointwise kernel (shape heuristics, classic tiles, register-tiled) gives the oracle's bits,
    including ragged M / K / N tails."""
    rng = np.random.default_rng(11)
    for (M, K, N, act) in [(33 * 33, 96, 32, 0), (300, 160, 64, 3), (129, 40, 21, 0), (257, 512, 256, 3), (1000, 24, 130, 1), (64, 16, 8, 0),
                           (38000, 20, 128, 3)]:   # last one: enough 128-wide tiles for the 8x8 register tile
        A = rng.standard_normal((M, K)).astype(np.float32)
        W = rng.standard_normal((N, K)).astype(np.float32)
        b = rng.standard_normal(N).astype(np.float32)
        ref = po.conv2d(A.reshape(1, M, K), W.reshape(N, 1, 1, K), b, padding=po.PAD_VALID, act=act).reshape(M, N)
        for variant in (0, 2, 3, 4, 5, 8, 16, 32, 64):
            got = api.pointwise(lib, A, W, b, act=act, variant=variant)
            assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (M, K, N, variant)


def check_app_option_edges(lib, key="meet_lite", W=640, H=480):
    """Less common combinations of the per-frame options: YUYV-only output behind flip + resize, mask-only calls in
    camera-blur mode, a caller-paced background ring, and the even-width rule for YUYV at the output size."""
    import pytest
    bg = synth.background()
    fr = np.stack([synth.frame(W, H, t=t) for t in range(2)])
    g = api.MaskGen(lib, model_path(key), W, H, max_batch=2)
    o = po.MaskGen(model_path(key), W, H)
    g.set_background(bg)
    g.set_output(flip_h=True, out_size=(320, 240))
    out, yuyv, mask = g.composite(fr, want_out=False, want_yuyv=True, want_mask=False)
    assert out is None and mask is None
    for b in range(2):
        assert np.array_equal(yuyv[b], o.composite_ex(fr[b], bg, flip_h=True, out_size=(320, 240))[1])
    g.close()

    g = api.MaskGen(lib, model_path(key), W, H)
    o = po.MaskGen(model_path(key), W, H)
    g.set_bgblur(25)
    assert np.array_equal(g.composite(fr[0], want_out=False, want_yuyv=False)[2], o.process(fr[0]))
    g.close()

    ring = np.stack([bg, bg[::-1].copy()])
    g = api.MaskGen(lib, model_path(key), W, H, max_batch=2)
    o = po.MaskGen(model_path(key), W, H)
    g.set_background_ring(ring, advance=0)
    g.set_background_cursor(1)
    out, _, _ = g.composite(fr)
    for b in range(2):
        assert np.array_equal(out[b], o.composite_ex(fr[b], ring[1])[0])
    with pytest.raises(api.BackscrubError, match="out of range"):
        g.set_background_cursor(2)
    g.close()

    g = api.MaskGen(lib, model_path(key), W, H)
    g.set_output(out_size=(321, 240))
    with pytest.raises(api.BackscrubError, match="even width"):
        g.composite(fr[0])
    assert g.composite(fr[0], want_yuyv=False)[0].shape == (240, 321, 3)
    assert g.process(fr[1]).shape == (H, W)
    g.close()


def check_post_variants(lib, key="meet_full", W=1280, H=720, n=3):
    """The fused post stage under every kernel variant and tile class: TMA-staged kernel vs the k_post_fast fallback
    (bsb_set_tuning("post_tma", 0/1)), BGR frames vs camera YUYV read in place, person-like / pure-noise / constant
    frames (mixed / person / background tiles), every output subset, still and animated backgrounds.  All bit-exact
    against the oracle."""
    bg = synth.background()
    ring = np.stack([bg, bg[::-1].copy(), np.roll(bg, 100, axis=1)])
    try:
        for tma in (1, 0):
            assert lib.bsb_set_tuning(b"post_tma", tma)
            for kind in ("person", "noise", "const"):
                g = api.MaskGen(lib, model_path(key), W, H, max_batch=n, flags=exact_flag(key))
                o = po.MaskGen(model_path(key), W, H)
                g.set_background(bg)
                bgr = np.stack([synth.frame(W, H, t=t, kind=kind) for t in range(n)])
                yin = np.stack([po.convert_rgb_to_yuyv(f) for f in bgr])
                ref = [o.composite(po.yuyv_to_bgr(yin[b]), bg) for b in range(n)]
                out, yuyv, mask = g.composite_yuyv(yin)                              # camera YUYV in, all outputs
                for b in range(n):
                    assert np.array_equal(mask[b], ref[b][2]) and np.array_equal(out[b], ref[b][0]) and np.array_equal(yuyv[b], ref[b][1]), (tma, kind, b)
                g.close()
            # output subsets from YUYV input (a fresh context each time: the IIR starts at zero like the oracle's)
            bgr = np.stack([synth.frame(W, H, t=t) for t in range(n)])
            yin = np.stack([po.convert_rgb_to_yuyv(f) for f in bgr])
            for want in ((True, False, False), (False, True, False), (False, False, True), (True, True, False)):
                g = api.MaskGen(lib, model_path(key), W, H, max_batch=n, flags=exact_flag(key))
                o = po.MaskGen(model_path(key), W, H)
                g.set_background(bg)
                bufs = [np.zeros((n, H, W, 3), np.uint8) if want[0] else None, np.zeros((n, H, W, 2), np.uint8) if want[1] else None,
                        np.zeros((n, H, W), np.uint8) if want[2] else None]
                g.composite_yuyv_into(yin, *bufs)
                for b in range(n):
                    r = o.composite(po.yuyv_to_bgr(yin[b]), bg)
                    for got, exp in zip(bufs, r):
                        if got is not None:
                            assert np.array_equal(got[b], exp), (tma, want, b)
                g.close()
            # BGR input + animated background ring (one image per frame, wraps inside the batch)
            g = api.MaskGen(lib, model_path(key), W, H, max_batch=n, flags=exact_flag(key))
            o = po.MaskGen(model_path(key), W, H)
            g.set_background_ring(ring, advance=1)
            for rep in range(2):
                out, yuyv, mask = g.composite(bgr)
                for b in range(n):
                    ro, ry, rm = o.composite(bgr[b], ring[(rep * n + b) % 3])
                    assert np.array_equal(mask[b], rm) and np.array_equal(out[b], ro) and np.array_equal(yuyv[b], ry), (tma, "ring", rep, b)
            # the same ring from camera YUYV
            out, yuyv, mask = g.composite_yuyv(yin)
            for b in range(n):
                ro, ry, rm = o.composite(po.yuyv_to_bgr(yin[b]), ring[(2 * n + b) % 3])
                assert np.array_equal(mask[b], rm) and np.array_equal(out[b], ro) and np.array_equal(yuyv[b], ry), (tma, "ring-yuyv", b)
            g.close()
    finally:
        lib.bsb_set_tuning(b"post_tma", 1)


def check_chain(lib, key, n=3, min_saved=15):
    """The one-kernel low-resolution chain (kernels_chain.cu) vs the oracle and vs the stand-alone kernels it replaces
    (bsb_set_tuning("cnn_chain", 0)): same bits, far fewer launches."""
    m = po.Model(model_path(key))
    rng = np.random.default_rng(23)
    outs, launches = {}, {}
    try:
        for chain in (1, 0):
            assert lib.bsb_set_tuning(b"cnn_chain", chain)
            g = api.MaskGen(lib, model_path(key), 640, 480, max_batch=n)
            x = rng.uniform(0, 1, (n, *g.in_hwc)).astype(np.float32) if chain else x
            outs[chain] = g.infer(x)
            g.set_background(synth.background())
            g.composite(np.stack([synth.frame(640, 480, t=t) for t in range(n)]))
            launches[chain] = g.launches_per_call
            g.close()
    finally:
        lib.bsb_set_tuning(b"cnn_chain", 1)
    assert launches[0] - launches[1] >= min_saved, (key, launches)
    assert np.array_equal(outs[1].view(np.uint32), outs[0].view(np.uint32)), f"{key}: chain and stand-alone kernels differ"
    for b in range(n):
        assert np.array_equal(outs[1][b].view(np.uint32), m.invoke(x[b])[0].view(np.uint32)), f"{key}: frame {b} differs from the oracle"


TC_FIXTURES = [("deeplab", 640, 480, 6), ("deeplab", 1280, 720, 4), ("bodypix", 640, 480, 6), ("bodypix", 1920, 1080, 3), ("bodypix", 3840, 2160, 2)]


def check_tc_default(lib, key, W, H, n):
    """The default path of the GEMM-dominated models (tensor-core 1x1 convs) on the committed synthetic fixtures: every
    per-frame decision, mask byte and composited byte equals the exact fp32 path (which the other tests pin to the
    oracle bit for bit) — 0 flips — and the logits stay within the reference's own fp32-conv tolerance."""
    bg = synth.background()
    for kind in ("person", "noise"):
        frames = np.stack([synth.frame(W, H, t=t, kind=kind) for t in range(n)])
        res = {}
        for name, flags in (("exact", FLAG_EXACT), ("default", 0)):
            g = api.MaskGen(lib, model_path(key), W, H, max_batch=n, flags=flags)
            assert g.uses_tensor_cores == (name == "default")
            g.set_background(bg)
            out, yuyv, mask = g.composite(frames)
            res[name] = (np.stack([g.stage_u8(2, b) for b in range(n)]), mask, out, yuyv)
            g.close()
        for a, b, what in zip(res["exact"], res["default"], ("decisions (ofinal)", "mask", "composite", "YUYV")):
            assert np.array_equal(a, b), f"{key} {W}x{H} {kind}: {what} differs between the exact and the tensor-core path ({int((a != b).sum())} bytes)"
    # logits: tensor-core vs oracle on one frame
    g = api.MaskGen(lib, model_path(key), 640, 480, max_batch=1)
    o = po.MaskGen(model_path(key), 640, 480)
    m = po.Model(model_path(key))
    o.process(synth.frame(640, 480, t=2))
    ref = m.invoke(o.input_f32)[0]
    got = g.infer(o.input_f32[None])[0]
    assert np.abs(got - ref).max() < 2e-3 * max(1.0, float(np.abs(ref).max()) / 10.0), float(np.abs(got - ref).max())
    g.close()


def check_fusion_switches(lib, key, n=2):
    """Every planner fusion of round 2 (stem + 1x1, one-launch pool + SE, low-resolution chain, resize folded into its
    1x1 conv, decoder stage kernel) can be switched off with bsb_set_tuning; the CNN output must not change by a bit,
    and it must equal the oracle's."""
    m = po.Model(model_path(key))
    g = api.MaskGen(lib, model_path(key), 640, 480, max_batch=n, flags=exact_flag(key))
    rng = np.random.default_rng(31)
    x = rng.uniform(-1 if key == "deeplab" else 0, 1, (n, *g.in_hwc)).astype(np.float32)
    base = g.infer(x)
    base_launches = None
    g.set_background(synth.background())
    frames = np.stack([synth.frame(640, 480, t=t) for t in range(n)])
    ref_out = g.composite(frames)
    base_launches = g.launches_per_call
    g.close()
    for b in range(n):
        assert np.array_equal(base[b].view(np.uint32), m.invoke(x[b])[0].view(np.uint32)), f"{key}: frame {b} differs from the oracle"
    defaults = {b"stem_pw": 0, b"pool_merge": 1, b"cnn_chain": 1, b"up_pw": 1, b"head": 1}
    if key == "deeplab":
        defaults = {b"dec_up": 1, b"dw_plane": 1}       # final resize folded into the argmax kernel; whole-plane atrous depthwise
    if key != "deeplab":
        defaults[b"pw_dws2"] = 1                         # 1x1 + stride-2 depthwise in one kernel
        defaults[b"up_staged"] = 1                       # resize + 1x1: staged interpolated rows
        defaults[b"stem_x2"] = 1                         # two output pixels per thread in the stem conv
    defaults[b"epi_static"] = 1                          # compile-time epilogues vs the generic run-time one
    defaults[b"dec_par"] = 1                             # frame-parallel decision + smoother vs one thread per pixel
    for sw, dflt in defaults.items():
        try:
            assert lib.bsb_set_tuning(sw, 1 - dflt)
            g = api.MaskGen(lib, model_path(key), 640, 480, max_batch=n, flags=exact_flag(key))
            # (CPU-suite time: a switch that only touches the decision stage needs the pipeline alone, a switch that only
            #  swaps CNN kernel variants needs the CNN output alone — the strictest check for it; planner switches get both)
            got = base if sw in (b"dec_par",) else g.infer(x)
            g.set_background(synth.background())
            out = ref_out if sw in (b"epi_static", b"stem_x2", b"up_staged") else g.composite(frames)
            if out is not ref_out:                                # (launches are counted while a pipeline call is enqueued)
                assert (g.launches_per_call >= base_launches) if dflt else (g.launches_per_call <= base_launches), (key, sw)
            g.close()
        finally:
            lib.bsb_set_tuning(sw, dflt)
        assert np.array_equal(got.view(np.uint32), base.view(np.uint32)), f"{key}: flipping {sw.decode()} changes the CNN output"
        for a, b in zip(out, ref_out):
            assert np.array_equal(a, b), f"{key}: flipping {sw.decode()} changes the composite"


def check_sub_batch(lib, key="deeplab", n=3, W=640, H=480, mbs=(0, 1, 8)):
    """Wide-layer segments executed a few frames at a time (engine.cu: find_segments / sub_batch_frames) produce the same
    bits as the whole batch at once, for any group size — including groups that do not divide the batch."""
    m = po.Model(model_path(key))
    rng = np.random.default_rng(11)
    x = None
    frames = np.stack([synth.frame(W, H, t=t) for t in range(n)])
    results = {}
    try:
        for mb in mbs:                  # off / one frame per group / two-ish frames per group (3 frames: groups of 2 + 1)
            assert lib.bsb_set_tuning(b"sub_batch_mb", mb)
            g = api.MaskGen(lib, model_path(key), W, H, max_batch=n, flags=exact_flag(key))
            g.set_background(synth.background())
            if x is None:
                x = rng.uniform(-1 if key == "deeplab" else 0, 1, (n, *g.in_hwc)).astype(np.float32)
            results[mb] = (g.infer(x), g.composite(frames), g.launches_per_call)
            g.close()
    finally:
        lib.bsb_set_tuning(b"sub_batch_mb", 0)
    base = results[0]
    assert np.array_equal(base[0][0].view(np.uint32), m.invoke(x[0])[0].view(np.uint32)), f"{key}: differs from the oracle"
    assert results[mbs[1]][2] > base[2], (key, "no segment was split", results[mbs[1]][2], base[2])
    for mb in mbs[1:]:
        assert np.array_equal(results[mb][0].view(np.uint32), base[0].view(np.uint32)), f"{key}: sub_batch_mb={mb} changes the CNN output"
        for a, b in zip(results[mb][1], base[1]):
            assert np.array_equal(a, b), f"{key}: sub_batch_mb={mb} changes the composite"


def check_mjpg_ingest(lib, key="meet_lite", W=640, H=480, n=2):
    """MJPG camera ingest (app/deepseg.cc:548-553): JPEG frames decoded on the GPU (NVJPG) feed the fused path.  The
    decoder is a library either side of the path (the reference uses libjpeg behind cv::VideoCapture); what is checked:
    the decoded frame is the same picture as libjpeg's (cv2.imdecode) up to IDCT / up-sampling rounding, and everything
    downstream of the decoded frame is bit-exact against the oracle."""
    import cv2
    import pytest
    g = api.MaskGen(lib, model_path(key), W, H, max_batch=n, flags=exact_flag(key))
    o = po.MaskGen(model_path(key), W, H)
    bg = synth.background()
    g.set_background(bg)
    frames = [synth.frame(W, H, t=t) for t in range(n)]
    params = [int(cv2.IMWRITE_JPEG_QUALITY), 92]
    if hasattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR"):
        params += [int(cv2.IMWRITE_JPEG_SAMPLING_FACTOR), int(cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422)]      # what UVC cameras send
    jpgs = [cv2.imencode(".jpg", f, params)[1].tobytes() for f in frames]
    dec = [g.decode_mjpg(j) for j in jpgs]
    for d, j in zip(dec, jpgs):
        ref = cv2.imdecode(np.frombuffer(j, np.uint8), cv2.IMREAD_COLOR)
        diff = np.abs(d.astype(np.int16) - ref.astype(np.int16))
        # NVJPG replicates the sub-sampled chroma where libjpeg interpolates it ("fancy up-sampling"): isolated pixels on
        # sharp colour edges differ by tens of levels, the picture as a whole by well under one (measured: max 29, mean 0.70)
        assert diff.mean() < 1.0 and (diff > 8).mean() < 0.01 and diff.max() <= 64, (int(diff.max()), float(diff.mean()), float((diff > 8).mean()))
    out, yuyv, mask = g.composite_mjpg(jpgs)
    for b in range(n):
        ro, ry, rm = o.composite(dec[b], bg)
        assert np.array_equal(mask[b], rm) and np.array_equal(out[b], ro) and np.array_equal(yuyv[b], ry), b
    small = cv2.imencode(".jpg", synth.frame(W // 2, H // 2))[1].tobytes()
    with pytest.raises(api.BackscrubError, match="frame size"):
        g.composite_mjpg([small])
    with pytest.raises(api.BackscrubError, match="JPEG"):
        g.composite_mjpg([b"\x00\x01not a jpeg at all" * 8])
    g.close()
