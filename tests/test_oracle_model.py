"""Oracle loader + interpreter vs an independent torch-CPU evaluation of the same
.tflite graphs, and vs the python flatbuffer reader (tools/tflite_graph.py)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests import synth
from tests.conftest import MODELS, model_path
from tools import tflite_graph as tg

EXPECT = {
    "mlkit": ([1, 256, 256, 3], [1, 256, 256, 1]),
    "meet_full": ([1, 144, 256, 3], [1, 144, 256, 2]),
    "meet_lite": ([1, 96, 160, 3], [1, 96, 160, 2]),
    "deeplab": ([1, 257, 257, 3], [1, 257, 257, 21]),
    "bodypix": ([1, 257, 257, 3], [1, 33, 33, 1]),
}


@pytest.mark.parametrize("key", list(MODELS))
def test_loader_matches_python_reader(key):
    m = po.Model(model_path(key))
    g = tg.load(model_path(key))
    assert m.n_tensors == len(g.tensors) and m.n_ops == len(g.ops)
    assert m.shape(m.input) == EXPECT[key][0] and m.shape(m.output) == EXPECT[key][1]
    assert (m.input, m.output) == (g.inputs[0], g.outputs[0])
    code = {v: k for k, v in tg.BUILTIN.items()}
    for i, op in enumerate(g.ops):
        kind, ins, out = m.op(i)
        assert kind == code[op.kind] and ins == op.inputs[:4] and out == op.outputs[0]
    # constants (incl. fp16 -> fp32 widening) identical to numpy's conversion
    for op in g.ops:
        if op.kind in ("CONV_2D", "DEPTHWISE_CONV_2D", "FULLY_CONNECTED", "CUSTOM"):
            for t in op.inputs[1:3]:
                assert np.array_equal(m.tensor(t), g.const_f32(t))


@pytest.mark.parametrize("key,tol", [("mlkit", 2e-4), ("meet_full", 5e-4), ("meet_lite", 5e-4),
                                     ("deeplab", 5e-4), ("bodypix", 5e-4)])
def test_interpreter_matches_torch_fp64(key, tol):
    """Every intermediate tensor of the oracle (fp32) vs torch fp64 evaluation."""
    from tests import torch_graph
    m = po.Model(model_path(key))
    _, h, w, _ = m.shape(m.input)
    g = po.MaskGen(model_path(key), 640, 480)
    g.process(synth.frame(640, 480, t=3))
    x = g.input_f32
    ref_out, ref_all, graph = torch_graph.run(model_path(key), x, keep=True)
    out = m.invoke(x)
    worst = 0.0
    for op in graph.ops:
        if op.kind == "DEQUANTIZE":
            continue
        t = op.outputs[0]
        got = m.tensor(t)[0].astype(np.float64)
        ref = ref_all[t]
        scale = max(1.0, float(np.abs(ref).max()))
        err = float(np.abs(got - ref).max()) / scale
        worst = max(worst, err)
        assert err < tol, f"op {op.idx} {op.kind} t{t}: rel-to-range err {err}"
    # decision agreement on the model output
    if key == "deeplab":
        assert (out[0].argmax(-1) == ref_out.argmax(-1)).mean() > 0.9995
    elif key.startswith("meet"):
        assert ((out[0][..., 0] < out[0][..., 1]) == (ref_out[..., 0] < ref_out[..., 1])).mean() > 0.9995
    else:
        assert ((out[0] > 0.65) == (ref_out > 0.65)).mean() > 0.9995


def test_flops_match_survey():
    # SURVEY.md Appendix A totals count 2*MAC of conv/dw/fc/tconv (pool/eltwise add a few %)
    for key, mflop in [("mlkit", 122.9), ("meet_full", 69.4), ("meet_lite", 29.1), ("deeplab", 1454.4), ("bodypix", 1261.4)]:
        got = po.Model(model_path(key)).flops / 1e6
        assert abs(got - mflop) / mflop < 0.05, (key, got)


def test_rejects_bad_file(tmp_path):
    p = tmp_path / "x_selfie.tflite"
    p.write_bytes(b"\0" * 64)
    with pytest.raises(RuntimeError):
        po.Model(str(p))
    with pytest.raises(RuntimeError):
        po.Model(str(tmp_path / "missing.tflite"))


# ---- third-party cross-check: OpenCV's dnn module imports .tflite files (cv2.dnn.readNetFromTFLite) and runs them
# with its own CPU kernels.  OpenCV is the reference's system dependency, and this is an implementation of the same
# graphs that shares no code with the oracle or with tests/torch_graph.py. ----
def _cv2dnn_forward(path, x, layer=None):
    import cv2
    net = cv2.dnn.readNetFromTFLite(path)
    net.setInput(np.ascontiguousarray(x.transpose(2, 0, 1)[None]))
    out = net.forward(layer) if layer else net.forward()
    return (out[0].transpose(1, 2, 0) if out.ndim == 4 else out), net


def test_mlkit_whole_model_matches_opencv_dnn():
    """BASELINE configs 1-2 model: the oracle's output agrees with OpenCV dnn's on the whole graph (136 ops: fp16
    DEQUANTIZE, convs, depthwise, SE blocks, hard-swish, bilinear resize, Convolution2DTransposeBias, logistic)."""
    g = po.MaskGen(model_path("mlkit"), 640, 480)
    g.process(synth.frame(640, 480, t=3))
    x = g.input_f32
    ref = po.Model(model_path("mlkit")).invoke(x)[0]
    got, _ = _cv2dnn_forward(model_path("mlkit"), x)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 5e-5                      # probabilities in [0, 1]
    assert np.array_equal(got > 0.65, ref > 0.65) or ((got > 0.65) != (ref > 0.65)).sum() <= 2


@pytest.mark.parametrize("key,n_ops", [("bodypix", 13), ("deeplab", 25)])
def test_mobilenet_prefix_matches_opencv_dnn(key, n_ops):
    """DeepLab / BodyPix: every layer up to the first atrous depthwise conv agrees with OpenCV dnn (its TFLite importer
    drops dilation_*_factor, so from that layer on OpenCV computes a different network and cannot serve as a pin)."""
    import cv2
    from tools import tflite_graph as tg
    g = po.MaskGen(model_path(key), 640, 480)
    g.process(synth.frame(640, 480, t=3))
    x = g.input_f32
    m = po.Model(model_path(key))
    m.invoke(x)
    gr = tg.load(model_path(key))
    net = cv2.dnn.readNetFromTFLite(model_path(key))
    names = set(net.getLayerNames())
    net.setInput(np.ascontiguousarray(x.transpose(2, 0, 1)[None]))
    checked = 0
    for op in gr.ops[:n_ops]:
        t = op.outputs[0]
        nm = gr.tensors[t].name
        layer = nm + "/activ" if nm + "/activ" in names else (nm if nm in names else None)
        if layer is None:
            continue
        assert op.opts.get("dil_h", 1) == 1
        out = net.forward(layer)
        ref = m.tensor(t)
        got = out[0].transpose(1, 2, 0).reshape(ref.shape)
        assert np.abs(got - ref).max() < 5e-4 * max(1.0, float(np.abs(ref).max())), (op.idx, op.kind)
        checked += 1
    assert checked >= n_ops - 1
    assert gr.ops[n_ops].kind == "DEPTHWISE_CONV_2D" and gr.ops[n_ops].opts["dil_h"] == 2
