"""CUDA kernel / planner / C-ABI logic on the GPU-less box: the product's .cu sources
compiled against tests/emu/cuemu.h (a test-only kernel-logic emulator) must reproduce the
oracle bit for bit.  This guards indexing, fusion and host logic before the `-m gpu`
parity tests run the real sm_100a build on a B200."""
import numpy as np
import pytest

from tests import parity_common as pc
from tests.emu.emu_lib import emu


@pytest.fixture(scope="module")
def lib():
    return emu()


@pytest.mark.parametrize("key,W,H", [("mlkit", 640, 480), ("meet_full", 640, 480), ("meet_lite", 640, 480),
                                     ("bodypix", 640, 480), ("meet_full", 1280, 720)])
def test_pipeline_bit_exact(lib, key, W, H):
    person = pc.check_pipeline(lib, key, W, H, n_frames=3, batch=2, tensors=True)
    assert 0.05 < person < 0.6


@pytest.mark.slow
def test_pipeline_deeplab(lib):
    pc.check_pipeline(lib, "deeplab", 640, 480, n_frames=2, batch=2)


@pytest.mark.parametrize("kind", ["noise", "const"])
def test_pipeline_edge_streams(lib, kind):
    pc.check_pipeline(lib, "mlkit", 640, 480, n_frames=2, frame_kind=kind)


def test_pipeline_fused_blocks(lib):
    """BSB_FLAG_FUSE_BLOCKS: expand+depthwise+SE+project in one kernel must stay bit-exact."""
    pc.check_pipeline(lib, "mlkit", 640, 480, n_frames=2, flags=8)
    pc.check_pipeline(lib, "meet_full", 640, 480, n_frames=2, flags=8)


def test_pipeline_ragged_geometry(lib):
    """odd sizes: ROI not tile-aligned, width not a multiple of 16, letter-boxed model input."""
    pc.check_pipeline(lib, "meet_lite", 324, 250, n_frames=2)
    pc.check_pipeline(lib, "mlkit", 322, 182, n_frames=1)
    # frame exactly half the model output: cv::resize of the mask takes its INTER_AREA 2x2 path (lib/libbackscrub.cc:366)
    pc.check_pipeline(lib, "meet_full", 128, 72, n_frames=2)


@pytest.mark.parametrize("key", ["mlkit", "meet_full", "bodypix"])
def test_every_tensor_bit_exact(lib, key):
    assert pc.check_tensors(lib, key) > 20


def test_post_variants_on_the_emulator(lib):
    """camera YUYV read in place by the pre-processing kernel, output subsets, rings (the TMA kernel itself needs a GPU)"""
    pc.check_post_variants(lib, "meet_lite", 320, 240, n=2)


@pytest.mark.parametrize("key", ["meet_lite", "mlkit"])
def test_chain_kernel(lib, key):
    pc.check_chain(lib, key, n=2)


@pytest.mark.parametrize("key", ["meet_lite", "mlkit", "deeplab"])
def test_fusion_switches(lib, key):
    pc.check_fusion_switches(lib, key, n=2)


def test_decision_window_across_calls(lib):
    """11 frames in calls of 4 + 4 + 3: the frame-parallel decision kernel rebuilds each state byte from the three latest
    decisions, the first two frames of a call from the state byte the previous call left (lib/libbackscrub.cc:314-361)."""
    pc.check_pipeline(lib, "meet_lite", 320, 240, n_frames=11, batch=4)


def test_sub_batched_segments(lib):
    pc.check_sub_batch(lib, "deeplab", n=3, mbs=(0, 8))      # (the one-frame-per-group case runs in the GPU suite)


def test_infer_batch(lib):
    pc.check_infer_batch(lib, "meet_lite", n=3)


@pytest.mark.slow
@pytest.mark.parametrize("key,n", [("bodypix", 8), ("deeplab", 5)])
def test_infer_batch_atrous_strips(lib, key, n):
    """batches large enough that the dilated depthwise layers take the strip kernel (dilation 2 / 4)."""
    pc.check_infer_batch(lib, key, n=n)


def test_stage_functions(lib):
    pc.check_stage_functions(lib)


def test_pointwise_variants(lib):
    pc.check_pointwise_variants(lib)


def test_app_stage_functions(lib):
    pc.check_app_stage_functions(lib)


def test_app_options(lib):
    pc.check_app_options(lib)


def test_app_option_edges(lib):
    pc.check_app_option_edges(lib)


def test_yuyv_ingest(lib):
    pc.check_yuyv_ingest(lib)


def test_mask_only_and_callbacks(lib):
    pc.check_mask_only_and_callbacks(lib, "meet_lite")


def test_errors(lib, tmp_path):
    pc.check_errors(lib, tmp_path)


def test_planner_fuses(lib):
    """launch counts: fused plan is far below one-kernel-per-TFLite-op (136 / 131 / 70 / 28 ops)."""
    from backscrub_b200 import api
    from tests.conftest import model_path
    for key, max_launches in [("mlkit", 80), ("meet_full", 84), ("deeplab", 70), ("bodypix", 33)]:
        g = api.MaskGen(lib, model_path(key), 640, 480)
        assert g.launches_per_call <= max_launches, (key, g.launches_per_call)
        g.close()


@pytest.mark.slow
def test_depthwise_plane_and_strips():
    """The whole-plane depthwise kernel (default for the 33x33 atrous layers) and the strip kernel it replaced
    (bsb_set_tuning("dw_plane", 0)) must both give the oracle's bits.  Run in a subprocess so the emulator's launch
    trace (CUEMU_TRACE) can prove which kernel ran."""
    import os
    import subprocess
    import sys
    from tests.conftest import ROOT
    code = ("import sys\n"
            "from tests import parity_common as pc\n"
            "from tests.emu.emu_lib import emu\n"
            "lib = emu()\n"
            "assert lib.bsb_set_tuning(b'dw_plane', int(sys.argv[1]))\n"
            "assert not lib.bsb_set_tuning(b'no_such_switch', 1)\n"
            "pc.check_infer_batch(lib, 'bodypix', n=2)\n"
            "pc.check_infer_batch(lib, 'deeplab', n=1)\n"
            "assert pc.check_tensors(lib, 'bodypix') > 20\n"
            "print('plane ok')\n")
    for v in ("1", "0"):
        out = subprocess.run([sys.executable, "-c", code, v], cwd=ROOT, env=dict(os.environ, CUEMU_TRACE="1", PYTHONPATH=ROOT),
                             capture_output=True, text=True, timeout=900)
        assert out.returncode == 0 and "plane ok" in out.stdout, out.stderr[-2000:]
        assert ("k_depthwise_plane" in out.stderr) == (v == "1")


def test_overlapped_host_call_on_the_emulator(lib):
    """the chunked schedule of bsb_composite_yuyv (the emulator runs its three streams in program order)"""
    pc.check_overlapped_host_call(lib, "meet_lite", 320, 240, n=17, oracle_frames=(0, 8, 16))
