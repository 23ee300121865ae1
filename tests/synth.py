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
