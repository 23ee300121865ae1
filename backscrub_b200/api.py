"""Host-side mirror of the reference's interface for the per-frame path.

Names follow the reference: `bs_maskgen_new / bs_maskgen_process / bs_maskgen_delete /
bs_tensorflow_version` (lib/libbackscrub.h:13-39), `alpha_blend`, `convert_rgb_to_yuyv`
(app/deepseg.cc:87-134), `load_background / grab_background` (app/background.h:14-23);
numpy arrays stand in for cv::Mat.  Everything computes on the GPU through the C ABI
(include/backscrub_b200.h); there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _binding as B


class BackscrubError(RuntimeError):
    pass


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class MaskGen:
    """One mask-generation context == one video stream (holds the temporal IIR state).

    `lib` is the bound shared library (the product library by default)."""

    def __init__(self, lib, modelname: str, width: int, height: int, *, threads: int = 2, device: int = 0,
                 max_batch: int = 1, flags: int = 0, ondebug=None, onprep=None, oninfer=None, onmask=None):
        self._lib = lib
        self._msgs = []
        self._user_debug = ondebug

        def _dbg(_ctx, msg):
            text = msg.decode("utf-8", "replace")
            self._msgs.append(text)
            if self._user_debug:
                self._user_debug(text)

        mk = lambda f: B.STAGE_CB(lambda _c: f()) if f else C.cast(None, B.STAGE_CB)
        self._cbs = (B.DEBUG_CB(_dbg), mk(onprep), mk(oninfer), mk(onmask))  # keep alive
        self._h = lib.bsb_maskgen_new_ex(modelname.encode(), width, height, device, max_batch, flags, *self._cbs, None)
        if not self._h:
            raise BackscrubError("".join(self._msgs).strip() or lib.bsb_last_error().decode())
        self.width, self.height, self.max_batch, self.device = width, height, max_batch, device
        self.out_width, self.out_height = width, height
        r, i, o = (C.c_int * 4)(), (C.c_int * 4)(), (C.c_int * 4)()
        ih, oh = (C.c_int * 3)(), (C.c_int * 3)()
        lib.bsb_geometry(self._h, r, i, o, ih, oh)
        self.roidim, self.in_roidim, self.out_roidim = list(r), list(i), list(o)
        self.in_hwc, self.out_hwc = list(ih), list(oh)

    # -- lifecycle
    def close(self):
        if getattr(self, "_h", None):
            self._lib.bsb_maskgen_delete(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def _fail(self, what):
        raise BackscrubError(f"{what}: {self._lib.bsb_last_error().decode()}")

    # -- the reference's calls
    def process(self, frame: np.ndarray) -> np.ndarray:
        """bs_maskgen_process: BGR u8 H x W x 3 -> mask u8 H x W (255 = background)."""
        frame = np.ascontiguousarray(frame, np.uint8)
        if frame.shape != (self.height, self.width, 3):
            raise BackscrubError(f"frame shape {frame.shape} != {(self.height, self.width, 3)}")
        mp, pitch = B.u8p(), C.c_size_t()
        if not self._lib.bsb_maskgen_process(self._h, _ptr(frame), self.width * 3, C.byref(mp), C.byref(pitch)):
            self._fail("bs_maskgen_process")
        return np.ctypeslib.as_array(mp, shape=(self.height, pitch.value))[:, : self.width].copy()

    def set_background(self, bg_raw: np.ndarray):
        bg_raw = np.ascontiguousarray(bg_raw, np.uint8)
        if not self._lib.bsb_set_background(self._h, _ptr(bg_raw), bg_raw.shape[1], bg_raw.shape[0], bg_raw.shape[1] * 3):
            self._fail("bsb_set_background")

    def set_background_ring(self, frames: np.ndarray, advance: int = 1):
        """Animated background: frames [count, bh, bw, 3] -> device-resident ring (app/background.cc video branch)."""
        frames = np.ascontiguousarray(frames, np.uint8)
        c, bh, bw, _ = frames.shape
        if not self._lib.bsb_set_background_ring(self._h, _ptr(frames), c, bw, bh, bw * 3, bh * bw * 3, int(advance)):
            self._fail("bsb_set_background_ring")

    def set_background_cursor(self, index: int):
        if not self._lib.bsb_set_background_cursor(self._h, int(index)):
            self._fail("bsb_set_background_cursor")

    def set_bgblur(self, ksize: int):
        """`-p bgblur:k` (app/deepseg.cc:657-658); 0 = off."""
        if not self._lib.bsb_set_bgblur(self._h, int(ksize)):
            self._fail("bsb_set_bgblur")

    def set_output(self, flip_h=False, flip_v=False, out_size=None):
        """cv::flip + cv::resize to the virtual-camera size (app/deepseg.cc:667-679); out_size = (w, h)."""
        ow, oh = out_size if out_size else (0, 0)
        if not self._lib.bsb_set_output(self._h, int(bool(flip_h)), int(bool(flip_v)), int(ow), int(oh)):
            self._fail("bsb_set_output")
        w, h = C.c_int(), C.c_int()
        self._lib.bsb_output_size(self._h, C.byref(w), C.byref(h))
        self.out_width, self.out_height = w.value, h.value

    def background(self) -> np.ndarray:
        out = np.empty((self.height, self.width, 3), np.uint8)
        if not self._lib.bsb_get_background(self._h, _ptr(out), self.width * 3):
            self._fail("bsb_get_background")
        return out

    def composite(self, frames: np.ndarray, want_out=True, want_yuyv=True, want_mask=True):
        """Fused path on HOST buffers: frames [n, H, W, 3] (or [H, W, 3]) -> (out, yuyv, mask)."""
        single = frames.ndim == 3
        frames = np.ascontiguousarray(frames[None] if single else frames, np.uint8)
        n = frames.shape[0]
        if frames.shape[1:] != (self.height, self.width, 3):
            raise BackscrubError(f"frame shape {frames.shape[1:]} != {(self.height, self.width, 3)}")
        fb, npx = self.height * self.width * 3, self.height * self.width
        ow, oh = self.out_width, self.out_height
        out = np.empty((n, oh, ow, 3), np.uint8) if want_out else None
        yuyv = np.empty((n, oh, ow, 2), np.uint8) if want_yuyv else None
        mask = np.empty((n, self.height, self.width), np.uint8) if want_mask else None
        ok = self._lib.bsb_composite(self._h, n, _ptr(frames), self.width * 3, fb,
                                     _ptr(out) if want_out else None, ow * 3, ow * oh * 3,
                                     _ptr(yuyv) if want_yuyv else None, ow * oh * 2,
                                     _ptr(mask) if want_mask else None, npx)
        if not ok:
            self._fail("bsb_composite")
        pick = (lambda a: a[0] if single and a is not None else a)
        return pick(out), pick(yuyv), pick(mask)

    def composite_into(self, frames: np.ndarray, out=None, yuyv=None, mask=None):
        """bsb_composite into caller-provided (e.g. pinned) host arrays; frames [n, H, W, 3]."""
        n = frames.shape[0]
        fb, npx = self.height * self.width * 3, self.height * self.width
        ow, opx = self.out_width, self.out_width * self.out_height
        ok = self._lib.bsb_composite(self._h, n, _ptr(frames), self.width * 3, fb,
                                     _ptr(out) if out is not None else None, ow * 3, opx * 3,
                                     _ptr(yuyv) if yuyv is not None else None, opx * 2,
                                     _ptr(mask) if mask is not None else None, npx)
        if not ok:
            self._fail("bsb_composite")

    def composite_yuyv_into(self, yuyv_frames: np.ndarray, out=None, yuyv=None, mask=None):
        """bsb_composite_yuyv: camera YUYV frames [n, H, W, 2] in, results into caller-provided host arrays."""
        n = yuyv_frames.shape[0]
        npx, opx = self.height * self.width, self.out_width * self.out_height
        ok = self._lib.bsb_composite_yuyv(self._h, n, _ptr(yuyv_frames), npx * 2,
                                          _ptr(out) if out is not None else None, opx * 3,
                                          _ptr(yuyv) if yuyv is not None else None, opx * 2,
                                          _ptr(mask) if mask is not None else None, npx)
        if not ok:
            self._fail("bsb_composite_yuyv")

    def composite_yuyv(self, yuyv_frames: np.ndarray):
        yuyv_frames = np.ascontiguousarray(yuyv_frames, np.uint8)
        n = yuyv_frames.shape[0]
        out = np.empty((n, self.out_height, self.out_width, 3), np.uint8)
        yuyv = np.empty((n, self.out_height, self.out_width, 2), np.uint8)
        mask = np.empty((n, self.height, self.width), np.uint8)
        self.composite_yuyv_into(yuyv_frames, out, yuyv, mask)
        return out, yuyv, mask

    def decode_mjpg(self, jpeg: bytes) -> np.ndarray:
        """one MJPG camera frame -> BGR, decoded on the GPU (NVJPG); what composite_mjpg feeds the pipeline"""
        buf = np.frombuffer(jpeg, np.uint8)
        out = np.empty((self.height, self.width, 3), np.uint8)
        if not self._lib.bsb_decode_mjpg(self._h, _ptr(buf), buf.size, _ptr(out)):
            self._fail("bsb_decode_mjpg")
        return out

    def composite_mjpg(self, jpegs):
        """MJPG camera ingest (app/deepseg.cc:548-553): a list of JPEG byte strings -> (out, yuyv, mask)"""
        bufs = [np.frombuffer(j, np.uint8) for j in jpegs]
        n = len(bufs)
        ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
        sizes = (C.c_size_t * n)(*[b.size for b in bufs])
        out = np.empty((n, self.out_height, self.out_width, 3), np.uint8)
        yuyv = np.empty((n, self.out_height, self.out_width, 2), np.uint8)
        mask = np.empty((n, self.height, self.width), np.uint8)
        opx = self.out_width * self.out_height
        if not self._lib.bsb_composite_mjpg(self._h, n, ptrs, sizes, _ptr(out), opx * 3, _ptr(yuyv), opx * 2, _ptr(mask), self.height * self.width):
            self._fail("bsb_composite_mjpg")
        return out, yuyv, mask

    def composite_device(self, n, d_frames, d_out=0, d_yuyv=0, d_mask=0, sync=False):
        """Fused path on DEVICE pointers (ints, e.g. torch.Tensor.data_ptr()); tightly packed frames."""
        fb, npx, opx = self.height * self.width * 3, self.height * self.width, self.out_width * self.out_height
        if not self._lib.bsb_composite_device(self._h, n, d_frames, fb, d_out or None, opx * 3, d_yuyv or None, opx * 2,
                                              d_mask or None, npx, int(sync)):
            self._fail("bsb_composite_device")

    def composite_yuyv_device(self, n, d_yuyv_in, d_out=0, d_yuyv=0, d_mask=0, sync=False):
        """Fused path from DEVICE-resident camera YUYV frames (tightly packed)."""
        npx, opx = self.height * self.width, self.out_width * self.out_height
        if not self._lib.bsb_composite_yuyv_device(self._h, n, d_yuyv_in, d_out or None, opx * 3, d_yuyv or None, opx * 2,
                                                   d_mask or None, npx, int(sync)):
            self._fail("bsb_composite_yuyv_device")

    def synchronize(self):
        if not self._lib.bsb_synchronize(self._h):
            self._fail("bsb_synchronize")

    @property
    def stream(self) -> int:
        return self._lib.bsb_stream(self._h) or 0

    # -- introspection
    def infer(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float32)
        n = 1 if x.ndim == 3 else x.shape[0]
        out = np.empty((n, *self.out_hwc), np.float32)
        if not self._lib.bsb_infer(self._h, n, _ptr(x), _ptr(out)):
            self._fail("bsb_infer")
        return out[0] if x.ndim == 3 else out

    def tensor(self, index: int, count: int):
        buf = np.empty(count, np.float32)
        n = self._lib.bsb_get_tensor(self._h, index, _ptr(buf), count)
        if n < 0:
            self._fail("bsb_get_tensor")
        return buf[:n] if n else None

    def stage_u8(self, which: int, frame: int = 0) -> np.ndarray:
        shape = self.in_hwc if which in (0, 1) else self.out_hwc[:2]
        buf = np.empty(int(np.prod(shape)), np.uint8)
        n = self._lib.bsb_get_stage_u8(self._h, which, frame, _ptr(buf), buf.size)
        if n < 0:
            self._fail("bsb_get_stage_u8")
        return buf.reshape(shape)

    def reset_state(self):
        if not self._lib.bsb_reset_state(self._h):
            self._fail("bsb_reset_state")

    def time_stage(self, stage: int, n_frames: int, iters: int) -> float:
        """ms per run of one stage (0 pre, 1 cnn, 2 decision, 3 post, 4 all) on device-resident buffers."""
        ms = self._lib.bsb_time_stage(self._h, stage, n_frames, iters)
        if ms < 0:
            self._fail("bsb_time_stage")
        return ms

    @property
    def uses_tensor_cores(self) -> bool:
        return bool(self._lib.bsb_uses_tensor_cores(self._h))

    @property
    def yuyv_native(self) -> bool:
        """the last composite_yuyv call read the camera YUYV frames in place (no BGR frame was materialised)"""
        return bool(self._lib.bsb_yuyv_native(self._h))

    @property
    def launches_per_call(self) -> int:
        return self._lib.bsb_launches_per_call(self._h, 1)

    @property
    def flops(self) -> float:
        return self._lib.bsb_model_flops(self._h)


def alpha_blend(lib, srca, srcb, mask, device=0):
    """app/deepseg.cc:108-134 (srca weight = mask, srcb weight = 255 - mask)."""
    srca, srcb, mask = (np.ascontiguousarray(a, np.uint8) for a in (srca, srcb, mask))
    out = np.empty_like(srca)
    if not lib.bsb_alpha_blend(device, _ptr(srca), _ptr(srcb), _ptr(mask), _ptr(out), mask.size):
        raise BackscrubError(lib.bsb_last_error().decode())
    return out


def convert_rgb_to_yuyv(lib, rgb, device=0):
    """app/deepseg.cc:87-106."""
    rgb = np.ascontiguousarray(rgb, np.uint8)
    h, w, _ = rgb.shape
    out = np.empty((h, w, 2), np.uint8)
    if not lib.bsb_convert_rgb_to_yuyv(device, _ptr(rgb), _ptr(out), w, h):
        raise BackscrubError(lib.bsb_last_error().decode())
    return out


def gaussian_blur(lib, src, ksize, device=0):
    """cv::GaussianBlur(src, Size(k,k), 0) 8UC3 (app/deepseg.cc:657-658)."""
    src = np.ascontiguousarray(src, np.uint8)
    out = np.empty_like(src)
    if not lib.bsb_gaussian_blur(device, _ptr(src), _ptr(out), src.shape[1], src.shape[0], int(ksize)):
        raise BackscrubError(lib.bsb_last_error().decode())
    return out


def gaussian_taps(lib, ksize):
    q = (C.c_int * 255)()
    if not lib.bsb_gaussian_taps(int(ksize), q):
        raise BackscrubError(lib.bsb_last_error().decode())
    return np.array(q[:ksize], np.int64)


def flip(lib, src, flip_h, flip_v, device=0):
    """cv::flip (app/deepseg.cc:667-673)."""
    src = np.ascontiguousarray(src, np.uint8)
    out = np.empty_like(src)
    if not lib.bsb_flip(device, _ptr(src), _ptr(out), src.shape[1], src.shape[0], int(bool(flip_h)), int(bool(flip_v))):
        raise BackscrubError(lib.bsb_last_error().decode())
    return out


def resize_u8c3(lib, src, dw, dh, device=0):
    """cv::resize(src, Size(dw, dh)) — what grab_background does (app/background.cc:178-194)."""
    src = np.ascontiguousarray(src, np.uint8)
    out = np.empty((dh, dw, 3), np.uint8)
    if not lib.bsb_resize_u8c3(device, _ptr(src), src.shape[1], src.shape[0], _ptr(out), dw, dh):
        raise BackscrubError(lib.bsb_last_error().decode())
    return out


def pointwise(lib, A, W, bias=None, act=0, use_tc=False, device=0, variant=None):
    """1x1 convolution stage: A [M, K] x W [N, K]^T (+ bias, activation) on the GPU."""
    A = np.ascontiguousarray(A, np.float32); W = np.ascontiguousarray(W, np.float32)
    M, K = A.shape; N = W.shape[0]
    b = np.ascontiguousarray(bias, np.float32) if bias is not None else None
    out = np.empty((M, N), np.float32)
    if not lib.bsb_pointwise(device, int(use_tc) if variant is None else int(variant), M, K, N, _ptr(A), _ptr(W), _ptr(b) if b is not None else None, act, _ptr(out)):
        raise BackscrubError(lib.bsb_last_error().decode())
    return out


def convert_yuyv_to_bgr(lib, yuyv, device=0):
    """cv::cvtColor(COLOR_YUV2BGR_YUYV) — the camera-frame ingest conversion."""
    yuyv = np.ascontiguousarray(yuyv, np.uint8)
    h, w, _ = yuyv.shape
    out = np.empty((h, w, 3), np.uint8)
    if not lib.bsb_convert_yuyv_to_bgr(device, _ptr(yuyv), _ptr(out), w, h):
        raise BackscrubError(lib.bsb_last_error().decode())
    return out
