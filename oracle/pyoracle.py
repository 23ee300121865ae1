"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs — never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MODEL_UNKNOWN, MODEL_BODYPIX, MODEL_DEEPLAB, MODEL_MEET, MODEL_MLKIT = range(5)
ACT = {"NONE": 0, "RELU": 1, "RELU_N1_TO_1": 2, "RELU6": 3}
PAD_SAME, PAD_VALID = 0, 1

f32p = C.POINTER(C.c_float)
u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int)


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference exists)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
    ref = os.path.join(_HERE, "_ref", "libref_tconv.so")
    if os.path.isdir("/root/reference") and not os.path.exists(ref):
        subprocess.call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        L.or_expf.restype = C.c_float
        L.or_expf.argtypes = [C.c_float]
        L.or_half_to_float.restype = C.c_float
        L.or_half_to_float.argtypes = [C.c_uint16]
        L.or_model_load.restype = C.c_void_p
        L.or_model_load.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
        L.or_model_free.argtypes = [C.c_void_p]
        for fn in ("or_model_num_tensors", "or_model_num_ops", "or_model_input", "or_model_output"):
            getattr(L, fn).argtypes = [C.c_void_p]
        L.or_model_tensor_shape.argtypes = [C.c_void_p, C.c_int, i32p]
        L.or_model_tensor_is_const.argtypes = [C.c_void_p, C.c_int]
        L.or_model_tensor_data.restype = f32p
        L.or_model_tensor_data.argtypes = [C.c_void_p, C.c_int]
        L.or_model_op.argtypes = [C.c_void_p, C.c_int, i32p, i32p, i32p, i32p]
        L.or_model_invoke.argtypes = [C.c_void_p, f32p]
        L.or_model_flops.restype = C.c_double
        L.or_model_flops.argtypes = [C.c_void_p]
        L.or_maskgen_new.restype = C.c_void_p
        L.or_maskgen_new.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_size_t]
        L.or_maskgen_delete.argtypes = [C.c_void_p]
        L.or_maskgen_geometry.argtypes = [C.c_void_p, i32p, i32p, i32p, i32p, i32p]
        L.or_maskgen_process.argtypes = [C.c_void_p, u8p, C.c_size_t, u8p]
        L.or_maskgen_post_from_output.argtypes = [C.c_void_p, f32p, u8p]
        for fn in ("or_maskgen_in_u8", "or_maskgen_filtered_u8", "or_maskgen_ofinal"):
            getattr(L, fn).restype = u8p
            getattr(L, fn).argtypes = [C.c_void_p]
        for fn in ("or_maskgen_input_f32", "or_maskgen_output_f32"):
            getattr(L, fn).restype = f32p
            getattr(L, fn).argtypes = [C.c_void_p]
        L.or_maskgen_model.restype = C.c_void_p
        L.or_maskgen_model.argtypes = [C.c_void_p]
        L.or_composite.argtypes = [C.c_void_p, u8p, C.c_size_t, u8p, C.c_int, C.c_int, C.c_size_t, u8p, u8p, u8p]
        L.or_composite_ex.argtypes = [C.c_void_p, u8p, C.c_size_t, u8p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, u8p, u8p, u8p]
        L.or_gaussian_kernel_q8.argtypes = [C.c_int, C.POINTER(C.c_int)]
        L.or_gaussian_blur_u8c3.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, u8p, C.c_size_t, C.c_int]
        L.or_flip_u8c3.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.or_model_type_from_name.argtypes = [C.c_char_p]
        _LIB = L
    return _LIB


def _f(a):
    return a.ctypes.data_as(f32p)


def _u(a):
    return a.ctypes.data_as(u8p)


def _cf(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _cu(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


# ---------------------------------------------------------------- scalar
def expf(x: float) -> float:
    return float(lib().or_expf(C.c_float(x)))


def expf_array(x: np.ndarray) -> np.ndarray:
    L = lib()
    return np.array([L.or_expf(C.c_float(float(v))) for v in x.ravel()], dtype=np.float32).reshape(x.shape)


# ---------------------------------------------------------------- single ops (HWC arrays)
def conv_out_size(in_size, k, stride, dil, padding):
    o, p = C.c_int(), C.c_int()
    lib().or_conv_out_size(in_size, k, stride, dil, padding, C.byref(o), C.byref(p))
    return o.value, p.value


def conv2d(x, w, bias, stride=(1, 1), dil=(1, 1), padding=PAD_SAME, act=0):
    x, w = _cf(x), _cf(w)
    ih, iw, ic = x.shape
    oc, kh, kw, _ = w.shape
    oh, _ = conv_out_size(ih, kh, stride[0], dil[0], padding)
    ow, _ = conv_out_size(iw, kw, stride[1], dil[1], padding)
    out = np.empty((oh, ow, oc), np.float32)
    b = _cf(bias) if bias is not None else None
    lib().or_conv2d(_f(x), ih, iw, ic, _f(w), oc, kh, kw, _f(b) if b is not None else None,
                    stride[0], stride[1], dil[0], dil[1], padding, act, _f(out), oh, ow)
    return out


def depthwise_conv2d(x, w, bias, stride=(1, 1), dil=(1, 1), padding=PAD_SAME, mult=1, act=0):
    x, w = _cf(x), _cf(w)
    ih, iw, ic = x.shape
    _, kh, kw, od = w.shape
    oh, _ = conv_out_size(ih, kh, stride[0], dil[0], padding)
    ow, _ = conv_out_size(iw, kw, stride[1], dil[1], padding)
    out = np.empty((oh, ow, od), np.float32)
    b = _cf(bias) if bias is not None else None
    lib().or_depthwise_conv2d(_f(x), ih, iw, ic, _f(w), kh, kw, _f(b) if b is not None else None,
                              stride[0], stride[1], dil[0], dil[1], padding, mult, act, _f(out), oh, ow)
    return out


def average_pool(x, fh, fw, stride=(1, 1), padding=PAD_VALID, act=0):
    x = _cf(x)
    ih, iw, c = x.shape
    oh, _ = conv_out_size(ih, fh, stride[0], 1, padding)
    ow, _ = conv_out_size(iw, fw, stride[1], 1, padding)
    out = np.empty((oh, ow, c), np.float32)
    lib().or_average_pool(_f(x), ih, iw, c, fh, fw, stride[0], stride[1], padding, act, _f(out), oh, ow)
    return out


def fully_connected(x, w, bias, act=0):
    x, w = _cf(x), _cf(w)
    batches, d = x.shape
    od = w.shape[0]
    out = np.empty((batches, od), np.float32)
    b = _cf(bias) if bias is not None else None
    lib().or_fully_connected(_f(x), batches, d, _f(w), od, _f(b) if b is not None else None, act, _f(out))
    return out


def resize_bilinear(x, oh, ow, align_corners=False, half_pixel=False):
    x = _cf(x)
    ih, iw, c = x.shape
    out = np.empty((oh, ow, c), np.float32)
    lib().or_resize_bilinear(_f(x), ih, iw, c, _f(out), oh, ow, int(align_corners), int(half_pixel))
    return out


def _unary(fn, x, *extra):
    x = _cf(x)
    out = np.empty_like(x)
    getattr(lib(), fn)(_f(x), _f(out), C.c_size_t(x.size), *extra)
    return out


def hard_swish(x): return _unary("or_hard_swish", x)
def logistic(x): return _unary("or_logistic", x)
def relu(x, act=1): return _unary("or_relu", x, act)


def add(a, b, act=0):
    a, b = _cf(a), _cf(b)
    out = np.empty_like(a)
    lib().or_add(_f(a), _f(b), _f(out), C.c_size_t(a.size), act)
    return out


def mul(a, b, act=0):
    a, b = _cf(a), _cf(b)
    c = a.shape[-1]
    out = np.empty_like(a)
    lib().or_mul(_f(a), _f(b), _f(out), C.c_size_t(a.size // c), c, int(b.size == c and a.size != c), act)
    return out


def tconv_bias(x, w, bias, stride=(2, 2), padding_same=True):
    x, w, bias = _cf(x), _cf(w), _cf(bias)
    ih, iw, ic = x.shape
    oc, kh, kw, _ = w.shape
    ph = max(0, kh - (ih - 1) % stride[0] - 1) if padding_same else 0
    pw = max(0, kw - (iw - 1) % stride[1] - 1) if padding_same else 0
    oh = stride[0] * (ih - 1) + kh - ph
    ow = stride[1] * (iw - 1) + kw - pw
    out = np.empty((oh, ow, oc), np.float32)
    lib().or_tconv_bias(_f(x), ih, iw, ic, _f(w), oc, kh, kw, _f(bias), stride[0], stride[1],
                        int(padding_same), _f(out), oh, ow)
    return out


_REF = None


def ref_tconv_lib():
    """oracle/_ref/libref_tconv.so — the reference's own transpose_conv_bias.cc (or None)."""
    global _REF
    if _REF is None:
        build()
        p = os.path.join(_HERE, "_ref", "libref_tconv.so")
        if not os.path.exists(p):
            return None
        _REF = C.CDLL(p)
        _REF.ref_tconv_bias.argtypes = [f32p, C.c_int, C.c_int, C.c_int, f32p, C.c_int, C.c_int, C.c_int, f32p,
                                        C.c_int, C.c_int, C.c_int, f32p, i32p, i32p]
    return _REF


def ref_tconv_bias(x, w, bias, stride=(2, 2), padding_same=True):
    R = ref_tconv_lib()
    x, w, bias = _cf(x), _cf(w), _cf(bias)
    ih, iw, ic = x.shape
    oc, kh, kw, _ = w.shape
    oh, ow = C.c_int(), C.c_int()
    rc = R.ref_tconv_bias(_f(x), ih, iw, ic, _f(w), oc, kh, kw, _f(bias), stride[0], stride[1],
                          int(padding_same), None, C.byref(oh), C.byref(ow))
    assert rc == 0
    out = np.empty((oh.value, ow.value, oc), np.float32)
    rc = R.ref_tconv_bias(_f(x), ih, iw, ic, _f(w), oc, kh, kw, _f(bias), stride[0], stride[1],
                          int(padding_same), _f(out), C.byref(oh), C.byref(ow))
    assert rc == 0
    return out


# ---------------------------------------------------------------- image ops
def resize_linear_u8(src, dw, dh):
    src = _cu(src)
    sh, sw = src.shape[:2]
    cn = 1 if src.ndim == 2 else src.shape[2]
    dst = np.empty((dh, dw) if src.ndim == 2 else (dh, dw, cn), np.uint8)
    lib().or_resize_linear_u8(_u(src), sw, sh, C.c_size_t(sw * cn), _u(dst), dw, dh, C.c_size_t(dw * cn), cn)
    return dst


def bilateral_d5(src, sigma_color=100.0, sigma_space=100.0):
    src = _cu(src)
    h, w, _ = src.shape
    dst = np.empty_like(src)
    lib().or_bilateral_d5_u8c3(_u(src), _u(dst), w, h, C.c_double(sigma_color), C.c_double(sigma_space))
    return dst


def convert_u8_f32(src, alpha, beta):
    src = _cu(src)
    dst = np.empty(src.shape, np.float32)
    lib().or_convert_u8_f32(_u(src), _f(dst), C.c_size_t(src.size), C.c_float(alpha), C.c_float(beta))
    return dst


def box_blur5(src):
    src = _cu(src)
    h, w = src.shape
    dst = np.empty_like(src)
    lib().or_box_blur5_u8(_u(src), C.c_size_t(w), _u(dst), C.c_size_t(w), w, h)
    return dst


def rgb2yuv(src):
    src = _cu(src)
    dst = np.empty_like(src)
    lib().or_rgb2yuv_u8(_u(src), _u(dst), C.c_size_t(src.size // 3))
    return dst


def convert_rgb_to_yuyv(src):
    src = _cu(src)
    h, w, _ = src.shape
    dst = np.empty((h, w, 2), np.uint8)
    lib().or_convert_rgb_to_yuyv(_u(src), _u(dst), w, h)
    return dst


def yuyv_to_bgr(yuyv):
    yuyv = _cu(yuyv)
    h, w, _ = yuyv.shape
    dst = np.empty((h, w, 3), np.uint8)
    lib().or_yuyv_to_bgr(_u(yuyv), _u(dst), w, h)
    return dst


def gaussian_kernel_q8(k):
    q = (C.c_int * 255)()
    if lib().or_gaussian_kernel_q8(k, q):
        raise ValueError(f"bad gaussian kernel size {k}")
    return np.array(q[:k], np.int64)


def gaussian_blur(src, k):
    src = _cu(src)
    h, w, _ = src.shape
    dst = np.empty_like(src)
    if lib().or_gaussian_blur_u8c3(_u(src), w, h, C.c_size_t(w * 3), _u(dst), C.c_size_t(w * 3), k):
        raise ValueError(f"bad gaussian kernel size {k}")
    return dst


def flip(src, flip_h, flip_v):
    src = _cu(src)
    h, w, _ = src.shape
    dst = np.empty_like(src)
    lib().or_flip_u8c3(_u(src), _u(dst), w, h, int(flip_h), int(flip_v))
    return dst


class FrameOpts(C.Structure):
    _fields_ = [("bgblur_k", C.c_int), ("flip_h", C.c_int), ("flip_v", C.c_int), ("out_w", C.c_int), ("out_h", C.c_int)]


def alpha_blend(srca, srcb, mask):
    srca, srcb, mask = _cu(srca), _cu(srcb), _cu(mask)
    out = np.empty_like(srca)
    lib().or_alpha_blend(_u(srca), _u(srcb), _u(mask), _u(out), C.c_size_t(mask.size))
    return out


# ---------------------------------------------------------------- model / pipeline
class Model:
    def __init__(self, path: str):
        err = C.create_string_buffer(256)
        self.h = lib().or_model_load(path.encode(), err, 256)
        if not self.h:
            raise RuntimeError(err.value.decode())
        self.path = path

    def close(self):
        if self.h:
            lib().or_model_free(self.h)
            self.h = None

    def __del__(self):
        self.close()

    @property
    def n_tensors(self): return lib().or_model_num_tensors(self.h)
    @property
    def n_ops(self): return lib().or_model_num_ops(self.h)
    @property
    def input(self): return lib().or_model_input(self.h)
    @property
    def output(self): return lib().or_model_output(self.h)
    @property
    def flops(self): return lib().or_model_flops(self.h)

    def shape(self, t):
        s = (C.c_int * 4)()
        r = lib().or_model_tensor_shape(self.h, t, s)
        return list(s)[:r]

    def is_const(self, t): return bool(lib().or_model_tensor_is_const(self.h, t))

    def tensor(self, t) -> np.ndarray:
        shp = self.shape(t)
        n = int(np.prod(shp)) if shp else 1
        p = lib().or_model_tensor_data(self.h, t)
        if not p:
            raise KeyError(t)
        return np.ctypeslib.as_array(p, shape=(n,)).reshape(shp).copy()

    def op(self, i):
        kind, n_in, out = C.c_int(), C.c_int(), C.c_int()
        ins = (C.c_int * 4)()
        lib().or_model_op(self.h, i, C.byref(kind), ins, C.byref(n_in), C.byref(out))
        return kind.value, list(ins)[:n_in.value], out.value

    def invoke(self, x: np.ndarray) -> np.ndarray:
        x = _cf(x)
        assert x.size == int(np.prod(self.shape(self.input)))
        rc = lib().or_model_invoke(self.h, _f(x))
        if rc:
            raise RuntimeError(f"or_model_invoke rc={rc}")
        return self.tensor(self.output)


class MaskGen:
    """Deterministic restatement of bs_maskgen_new/process (+ composite)."""

    def __init__(self, model_path: str, width: int, height: int):
        err = C.create_string_buffer(256)
        self.h = lib().or_maskgen_new(model_path.encode(), width, height, err, 256)
        if not self.h:
            raise RuntimeError(err.value.decode())
        self.W, self.H = width, height
        r, i, o = (C.c_int * 4)(), (C.c_int * 4)(), (C.c_int * 4)()
        ih, oh = (C.c_int * 3)(), (C.c_int * 3)()
        lib().or_maskgen_geometry(self.h, r, i, o, ih, oh)
        self.roidim, self.in_roidim, self.out_roidim = list(r), list(i), list(o)
        self.in_hwc, self.out_hwc = list(ih), list(oh)

    def close(self):
        if self.h:
            lib().or_maskgen_delete(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def process(self, frame_bgr: np.ndarray) -> np.ndarray:
        frame_bgr = _cu(frame_bgr)
        assert frame_bgr.shape == (self.H, self.W, 3)
        mask = np.empty((self.H, self.W), np.uint8)
        rc = lib().or_maskgen_process(self.h, _u(frame_bgr), C.c_size_t(self.W * 3), _u(mask))
        if rc:
            raise RuntimeError(f"or_maskgen_process rc={rc}")
        return mask

    def post_from_output(self, out_f32: np.ndarray) -> np.ndarray:
        out_f32 = _cf(out_f32)
        mask = np.empty((self.H, self.W), np.uint8)
        lib().or_maskgen_post_from_output(self.h, _f(out_f32), _u(mask))
        return mask

    def composite(self, frame_bgr, bg_raw, want_yuyv=True):
        frame_bgr, bg_raw = _cu(frame_bgr), _cu(bg_raw)
        bh, bw = bg_raw.shape[:2]
        out = np.empty((self.H, self.W, 3), np.uint8)
        yuyv = np.empty((self.H, self.W, 2), np.uint8) if want_yuyv else None
        mask = np.empty((self.H, self.W), np.uint8)
        rc = lib().or_composite(self.h, _u(frame_bgr), C.c_size_t(self.W * 3), _u(bg_raw), bw, bh,
                                C.c_size_t(bw * 3), _u(out), _u(yuyv) if want_yuyv else None, _u(mask))
        if rc:
            raise RuntimeError(f"or_composite rc={rc}")
        return out, yuyv, mask

    def composite_ex(self, frame_bgr, bg_raw=None, bgblur=0, flip_h=False, flip_v=False, out_size=None, want_yuyv=True, reuse=False):
        """app/deepseg.cc:640-681 with its options; bg_raw None = blur-the-camera-frame mode."""
        frame_bgr = _cu(frame_bgr)
        ow, oh = out_size if out_size else (self.W, self.H)
        o = FrameOpts(int(bgblur), int(flip_h), int(flip_v), ow, oh)
        if bg_raw is not None:
            bg_raw = _cu(bg_raw)
            bh, bw = bg_raw.shape[:2]
        else:
            bh = bw = 0
        # result arrays are kept per context (same-shape calls reuse them when `reuse`): a timed multi-threaded run then does
        # not spend its time in mmap / page faults of three fresh frame-sized arrays per frame
        if reuse and getattr(self, "_bufs", None) is not None and self._bufs[0].shape == (oh, ow, 3):
            out, yuyv_b, mask = self._bufs
        else:
            out, yuyv_b, mask = np.empty((oh, ow, 3), np.uint8), np.empty((oh, ow, 2), np.uint8), np.empty((self.H, self.W), np.uint8)
            if reuse:
                self._bufs = (out, yuyv_b, mask)
        yuyv = yuyv_b if want_yuyv else None
        rc = lib().or_composite_ex(self.h, _u(frame_bgr), C.c_size_t(self.W * 3), _u(bg_raw) if bg_raw is not None else None,
                                   bw, bh, C.c_size_t(bw * 3), C.byref(o), _u(out), _u(yuyv) if want_yuyv else None, _u(mask))
        if rc:
            raise RuntimeError(f"or_composite_ex rc={rc}")
        return out, yuyv, mask

    def _arr_u8(self, fn, shape):
        p = getattr(lib(), fn)(self.h)
        return np.ctypeslib.as_array(p, shape=(int(np.prod(shape)),)).reshape(shape).copy()

    def _arr_f32(self, fn, shape):
        p = getattr(lib(), fn)(self.h)
        return np.ctypeslib.as_array(p, shape=(int(np.prod(shape)),)).reshape(shape).copy()

    @property
    def in_u8(self): return self._arr_u8("or_maskgen_in_u8", self.in_hwc)
    @property
    def filtered_u8(self): return self._arr_u8("or_maskgen_filtered_u8", self.in_hwc)
    @property
    def input_f32(self): return self._arr_f32("or_maskgen_input_f32", self.in_hwc)
    @property
    def output_f32(self): return self._arr_f32("or_maskgen_output_f32", self.out_hwc)
    @property
    def ofinal(self): return self._arr_u8("or_maskgen_ofinal", self.out_hwc[:2])
