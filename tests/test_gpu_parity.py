"""GPU parity tests proper (`-m gpu`): the sm_100a library, called through the C ABI,
vs the CPU oracle on the same seeded inputs — bit-exact for every integer stage and,
because the kernels keep the oracle's accumulation order, for every CNN activation too."""
import numpy as np
import pytest

from tests import parity_common as pc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    import backscrub_b200 as bs
    L = bs.lib()
    assert L.bsb_device_count() > 0, "no CUDA device: the gpu tests must run on the B200 box"
    return L


@pytest.mark.parametrize("key,W,H", [("mlkit", 640, 480), ("meet_full", 640, 480), ("meet_lite", 640, 480),
                                     ("bodypix", 640, 480), ("meet_full", 1280, 720), ("deeplab", 640, 480)])
def test_pipeline_bit_exact(lib, key, W, H):
    person = pc.check_pipeline(lib, key, W, H, n_frames=4, batch=3, tensors=True)
    assert 0.05 < person < 0.6


def test_pipeline_deeplab_720p(lib):          # BASELINE config 3
    pc.check_pipeline(lib, "deeplab", 1280, 720, n_frames=2, batch=2)


@pytest.mark.parametrize("kind", ["noise", "const"])
def test_pipeline_edge_streams(lib, kind):
    pc.check_pipeline(lib, "mlkit", 640, 480, n_frames=3, frame_kind=kind)


def test_pipeline_ragged_geometry(lib):
    pc.check_pipeline(lib, "meet_lite", 324, 250, n_frames=2)
    pc.check_pipeline(lib, "mlkit", 322, 182, n_frames=2)
    pc.check_pipeline(lib, "meet_full", 720, 1280, n_frames=1)     # portrait: letter-boxed input
    pc.check_pipeline(lib, "meet_full", 128, 72, n_frames=2)       # mask resize is an exact 2x down-scale (INTER_AREA path)


@pytest.mark.parametrize("key", ["deeplab", "bodypix"])
def test_sub_batched_segments(lib, key):
    pc.check_sub_batch(lib, key, n=3)


def test_pipeline_fused_blocks(lib):
    pc.check_pipeline(lib, "mlkit", 640, 480, n_frames=3, flags=8)
    pc.check_pipeline(lib, "meet_full", 1280, 720, n_frames=3, flags=8)


def test_graph_and_eager_agree(lib):
    pc.check_pipeline(lib, "mlkit", 640, 480, n_frames=2, flags=2)  # BSB_FLAG_NO_GRAPH


@pytest.mark.parametrize("key", ["mlkit", "meet_full", "meet_lite", "deeplab", "bodypix"])
def test_every_tensor_bit_exact(lib, key):
    assert pc.check_tensors(lib, key) > 20


@pytest.mark.parametrize("key", ["mlkit", "meet_full", "deeplab"])
def test_infer_batch(lib, key):
    pc.check_infer_batch(lib, key, n=4)


def test_stage_functions(lib):
    pc.check_stage_functions(lib)


def test_yuyv_ingest(lib):
    pc.check_yuyv_ingest(lib, "meet_full", 1280, 720, n=3)
    pc.check_yuyv_ingest(lib, "mlkit", 640, 480, n=2)


@pytest.mark.parametrize("key,W,H", [("meet_full", 1280, 720), ("mlkit", 640, 480), ("bodypix", 1920, 1080), ("meet_lite", 336, 250)])
def test_post_kernel_variants(lib, key, W, H):
    pc.check_post_variants(lib, key, W, H, n=3 if W < 1920 else 2)


@pytest.mark.parametrize("key,n", [("meet_full", 32), ("meet_lite", 5), ("mlkit", 16)])
def test_chain_kernel(lib, key, n):
    pc.check_chain(lib, key, n=n)


@pytest.mark.parametrize("key", ["meet_full", "meet_lite", "mlkit", "bodypix", "deeplab"])
def test_fusion_switches(lib, key):
    pc.check_fusion_switches(lib, key, n=3)


def test_mjpg_ingest(lib):
    pc.check_mjpg_ingest(lib, "meet_lite", 640, 480, n=2)
    pc.check_mjpg_ingest(lib, "meet_full", 1280, 720, n=1)


def test_mask_only_and_callbacks(lib):
    pc.check_mask_only_and_callbacks(lib, "mlkit")


def test_errors(lib, tmp_path):
    pc.check_errors(lib, tmp_path)


def test_large_batch_stream_consistency(lib):
    """Size-independent property at full batch: one 32-frame launch == 32 single-frame launches
    (same stream, IIR advancing in order), and batches of different streams are independent."""
    from backscrub_b200 import api
    from tests import synth
    from tests.conftest import model_path
    W, H, n = 1280, 720, 32
    frames = np.stack([synth.frame(W, H, t=t) for t in range(n)])
    bg = synth.background()
    a = api.MaskGen(lib, model_path("meet_full"), W, H, max_batch=n)
    b = api.MaskGen(lib, model_path("meet_full"), W, H, max_batch=1)
    a.set_background(bg); b.set_background(bg)
    out_a, yuyv_a, mask_a = a.composite(frames)
    for t in range(n):
        o, y, m = b.composite(frames[t])
        assert np.array_equal(out_a[t], o) and np.array_equal(yuyv_a[t], y) and np.array_equal(mask_a[t], m), t
    # outside roidim nothing but background; mask extremes blend exactly
    assert (mask_a[:, :, :] <= 255).all()
    sel = mask_a == 255
    assert np.array_equal(out_a[sel], np.broadcast_to(a.background(), out_a.shape)[sel])
    sel0 = mask_a == 0
    assert np.array_equal(out_a[sel0], frames[sel0])
    a.close(); b.close()


def test_4k_bodypix(lib):                     # BASELINE config 5 geometry (documented out_roidim deviation)
    pc.check_pipeline(lib, "bodypix", 3840, 2160, n_frames=1)


def test_overlapped_host_call(lib):
    pc.check_overlapped_host_call(lib, "meet_full", 1280, 720, n=19)


@pytest.mark.parametrize("key,W,H,n,batch", [("meet_full", 1280, 720, 27, 10), ("mlkit", 640, 480, 11, 4), ("deeplab", 640, 480, 7, 3)])
def test_decision_window_across_calls(lib, key, W, H, n, batch):
    """Frame-parallel decision + temporal smoother (k_decision_par / k_decision_up_par): batches longer than the eight
    frame lanes of a block, and the state byte carried from call to call (27 frames in calls of 10 + 10 + 7, ...)."""
    pc.check_pipeline(lib, key, W, H, n_frames=n, batch=batch)


def test_app_stage_functions(lib):            # `-p bgblur:k` Gaussian taps + blur, cv::flip
    pc.check_app_stage_functions(lib)


def test_app_options(lib):                    # green default, bgblur (still / camera), flip, vcam resize, animated ring
    pc.check_app_options(lib)


def test_app_options_720p_meet(lib):          # the BASELINE config-4 geometry with the blur-my-background mode
    pc.check_app_options(lib, "meet_full", 1280, 720)


def test_pointwise_variants(lib):             # classic / register-tiled exact FFMA kernels: same bits
    pc.check_pointwise_variants(lib)


@pytest.mark.parametrize("key,n", [("bodypix", 8), ("deeplab", 16)])
def test_infer_batch_atrous_strips(lib, key, n):   # dilated depthwise strips + register-tiled pointwise at real batch sizes
    pc.check_infer_batch(lib, key, n=n)
