"""Tensor-core (tcgen05 3xTF32) pointwise path — `-m gpu`.  Not bit-identical to the oracle by
construction (different summation order, split products), so the bar is the one the reference's own
XNNPACK-vs-builtin tests use for fp32 convs (3e-6 relative, conv_2d_tester.cc:104) on the GEMM, and
decision agreement / mask IoU >= 0.999 on whole models (BASELINE.json north_star)."""
import numpy as np
import pytest

from backscrub_b200 import api
from oracle import pyoracle as po
from tests import synth
from tests.conftest import model_path

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    import backscrub_b200 as bs
    L = bs.lib()
    assert L.bsb_device_count() > 0
    return L


@pytest.mark.parametrize("M,K,N", [(1089, 32, 192), (1089 * 3, 96, 16), (4225, 48, 12), (1089, 160, 256), (2178, 480, 80),
                                   (1089, 512, 256), (1089, 256, 21), (1500, 288, 48), (1089, 80, 480), (300, 256, 1), (128, 32, 16)])
def test_gemm_tc_vs_exact(lib, M, K, N):
    rng = np.random.default_rng(M + K + N)
    A = rng.standard_normal((M, K)).astype(np.float32) * 2
    W = rng.standard_normal((N, K)).astype(np.float32) * 0.2
    b = rng.standard_normal(N).astype(np.float32)
    exact = api.pointwise(lib, A, W, b, act=3, use_tc=False)
    ref64 = np.clip(A.astype(np.float64) @ W.astype(np.float64).T + b, 0, 6)
    assert np.abs(exact - ref64).max() < 1e-4
    got = api.pointwise(lib, A, W, b, act=3, use_tc=True)
    scale = np.abs(A).astype(np.float64) @ np.abs(W).astype(np.float64).T + 1.0     # |a|.|w| bound per output
    rel = np.abs(got - ref64) / scale
    assert rel.max() < 3e-6, rel.max()
    assert np.abs(got - exact).max() < 2e-4


def test_exact_pointwise_matches_oracle_bits(lib):
    rng = np.random.default_rng(5)
    A = rng.standard_normal((33 * 33, 96)).astype(np.float32)
    W = rng.standard_normal((32, 96)).astype(np.float32)
    b = rng.standard_normal(32).astype(np.float32)
    got = api.pointwise(lib, A, W, b, act=0, use_tc=False)
    ref = po.conv2d(A.reshape(33, 33, 96), W.reshape(32, 1, 1, 96), b, padding=po.PAD_VALID).reshape(-1, 32)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("key", ["deeplab", "bodypix"])
def test_model_tc_vs_oracle(lib, key):
    g = api.MaskGen(lib, model_path(key), 640, 480, max_batch=2, flags=4)       # BSB_FLAG_TENSOR_CORES
    o = po.MaskGen(model_path(key), 640, 480)
    m = po.Model(model_path(key))
    fr = synth.frame(640, 480, t=2)
    o.process(fr)
    x = o.input_f32
    ref = m.invoke(x)[0]
    out = g.infer(np.stack([x, x]))
    assert np.array_equal(out[0], out[1])
    err = np.abs(out[0] - ref).max()
    assert err < 2e-3, err
    if key == "deeplab":
        a, b = out[0].argmax(-1) == 15, ref.argmax(-1) == 15
    else:
        a, b = out[0][..., 0] > 0.65, ref[..., 0] > 0.65
    iou = (a & b).sum() / max(1, (a | b).sum())
    assert iou >= 0.999 and (a != b).sum() <= 3, (iou, (a != b).sum())
    g.close()


@pytest.mark.parametrize("key,W,H,n", [(k, w, h, n) for (k, w, h, n) in __import__("tests.parity_common", fromlist=["x"]).TC_FIXTURES])
def test_tc_default_path_has_no_flips_on_the_fixtures(lib, key, W, H, n):
    from tests import parity_common as pc
    pc.check_tc_default(lib, key, W, H, n)


def test_meet_stays_exact_by_default(lib):
    g = api.MaskGen(lib, model_path("meet_full"), 640, 480)
    assert not g.uses_tensor_cores
    g.close()
    g = api.MaskGen(lib, model_path("meet_full"), 640, 480, flags=4)
    assert g.uses_tensor_cores
    g.close()


def test_pipeline_tc_mask_iou(lib):
    W, H = 1280, 720
    g = api.MaskGen(lib, model_path("deeplab"), W, H, max_batch=2, flags=4)
    o = po.MaskGen(model_path("deeplab"), W, H)
    bg = synth.background()
    g.set_background(bg)
    frames = np.stack([synth.frame(W, H, t=t) for t in range(2)])
    out, yuyv, mask = g.composite(frames)
    for b in range(2):
        ro, ry, rm = o.composite(frames[b], bg)
        a, r = mask[b] < 128, rm < 128
        assert (a & r).sum() / max(1, (a | r).sum()) >= 0.999
        assert np.abs(out[b].astype(int) - ro.astype(int)).max() <= 255        # report, bound below on the bulk
        assert (np.abs(out[b].astype(int) - ro.astype(int)) > 1).mean() < 1e-3
    g.close()
