"""Deterministic synthetic frames: a real person image, translated
sinusoidally, plus seeded uniform noise; also pure-noise and constant streams."""
import os

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CACHE = {}


def person_base(width: int, height: int) -> np.ndarray:
    key = (width, height)
    if key not in _CACHE:
        img = cv2.imread(os.path.join(ROOT, "backgrounds", "screenshot.jpg"))
        left = img[:, : img.shape[1] // 2]
        _CACHE[key] = cv2.resize(left, (width, height), interpolation=cv2.INTER_LINEAR)
    return _CACHE[key]


def frame(width: int, height: int, t: int = 0, stream: int = 0, kind: str = "person") -> np.ndarray:
    """BGR u8 frame t of stream s."""
    rng = np.random.default_rng(1000 * stream + t)
    if kind == "noise":
        return rng.integers(0, 256, (height, width, 3), dtype=np.uint8)
    if kind == "const":
        return np.full((height, width, 3), 128, np.uint8)
    base = person_base(width, height)
    dx = int(np.floor(8 * np.sin(2 * np.pi * t / 64)))
    shifted = np.roll(base, dx, axis=1)
    noise = rng.integers(-3, 4, shifted.shape, dtype=np.int16)
    return np.clip(shifted.astype(np.int16) + noise, 0, 255).astype(np.uint8)


def background() -> np.ndarray:
    return cv2.imread(os.path.join(ROOT, "backgrounds", "background_bauhaus.png"))


def yuyv_frame(width: int, height: int, t: int = 0, stream: int = 0, kind: str = "person") -> np.ndarray:
    """Camera-format (YUYV, H x W x 2) version of frame(): BT.601 limited-range encode of the BGR frame with plain
    numpy — any valid YUYV frame is a legitimate camera frame, so this does not need to (and does not) depend on the
    oracle.  Used by bench.py to synthesise its input."""
    bgr = frame(width, height, t, stream, kind).astype(np.float32)
    b, g, r = bgr[..., 0], bgr[..., 1], bgr[..., 2]
    y = 16.0 + 0.257 * r + 0.504 * g + 0.098 * b
    u = 128.0 - 0.148 * r - 0.291 * g + 0.439 * b
    v = 128.0 + 0.439 * r - 0.368 * g - 0.071 * b
    out = np.empty((height, width, 2), np.uint8)
    out[..., 0] = np.clip(np.rint(y), 0, 255)
    uu = (u[:, 0::2] + u[:, 1::2]) * 0.5
    vv = (v[:, 0::2] + v[:, 1::2]) * 0.5
    out[:, 0::2, 1] = np.clip(np.rint(uu), 0, 255)
    out[:, 1::2, 1] = np.clip(np.rint(vv), 0, 255)
    return out
