"""pytest configuration: markers, paths, shared fixtures.

`-m "not gpu"` runs on a GPU-less box (oracle vs the reference's golden vectors and
cv2, host logic, C-ABI symbol checks, kernel-logic emulation); `-m gpu` are the
parity tests proper and call the CUDA path through the C-ABI.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODELS = {
    "mlkit": "selfiesegmentation_mlkit-256x256-2021_01_19-v1215.f16.tflite",
    "meet_full": "segm_full_v679.tflite",
    "meet_lite": "segm_lite_v681.tflite",
    "deeplab": "deeplabv3_257_mv_gpu.tflite",
    "bodypix": "body-pix-float-050-8.tflite",
}


def model_path(key: str) -> str:
    return os.path.join(ROOT, "models", MODELS[key])


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


@pytest.fixture(scope="session")
def root():
    return ROOT
