"""Independent evaluation of a .tflite graph with torch CPU ops (test helper).

This is a *second* restatement of the TFLite op semantics (NHWC/OHWI, TF SAME padding,
half-pixel / align-corners bilinear, k2s2 transpose-conv-bias), written against torch's
own conv/pool kernels, used to cross-check the C oracle's interpreter end to end
(whole-model outputs are otherwise unpinned).  fp64 by default so it
also bounds the oracle's fp32 rounding error.
"""
import numpy as np
import torch
import torch.nn.functional as F

from tools import tflite_graph as tg


def _act(x, act):
    if act == "RELU":
        return torch.relu(x)
    if act == "RELU6":
        return torch.clamp(x, 0, 6)
    if act == "RELU_N1_TO_1":
        return torch.clamp(x, -1, 1)
    return x


def _same_pad(in_size, k, stride, dil):
    eff = (k - 1) * dil + 1
    out = (in_size + stride - 1) // stride
    total = max(0, (out - 1) * stride + eff - in_size)
    return total // 2, total - total // 2


def _resize(x, oh, ow, align_corners, half_pixel):
    # x: [1, C, H, W]; literal restatement of reference/resize_bilinear.h in float64
    _, c, ih, iw = x.shape
    hs = (ih - 1) / (oh - 1) if (align_corners and oh > 1) else ih / oh
    ws = (iw - 1) / (ow - 1) if (align_corners and ow > 1) else iw / ow
    def coords(n_out, scale, n_in):
        v = torch.arange(n_out, dtype=x.dtype)
        s = (v + 0.5) * scale - 0.5 if half_pixel else v * scale
        lo = torch.clamp(torch.floor(s), min=0).long()
        hi = torch.clamp(torch.ceil(s), max=n_in - 1).long()
        return s, lo, hi
    sy, y0, y1 = coords(oh, np.float32(hs).item(), ih)
    sx, x0, x1 = coords(ow, np.float32(ws).item(), iw)
    dy = (sy - y0.to(x.dtype)).view(1, 1, oh, 1)
    dx = (sx - x0.to(x.dtype)).view(1, 1, 1, ow)
    g = lambda yy, xx: x[:, :, yy][:, :, :, xx]
    return g(y0, x0) * (1 - dy) * (1 - dx) + g(y1, x0) * dy * (1 - dx) + g(y0, x1) * (1 - dy) * dx + g(y1, x1) * dy * dx


def run(path, inp_hwc, dtype=torch.float64, keep=False):
    """Evaluate the graph; returns output HWC numpy (and all tensors when keep=True)."""
    g = tg.load(path)
    T = {}
    x = torch.from_numpy(np.asarray(inp_hwc, np.float32)).to(dtype)
    T[g.inputs[0]] = x.permute(2, 0, 1).unsqueeze(0)          # NCHW
    cf = lambda t: torch.from_numpy(g.const_f32(t)).to(dtype)
    for op in g.ops:
        k = op.kind
        if k == "DEQUANTIZE":
            continue
        o = op.opts
        a = T.get(op.inputs[0]) if op.inputs and op.inputs[0] in T else None
        if k == "CONV_2D":
            w = cf(op.inputs[1]).permute(0, 3, 1, 2)            # OHWI -> OIHW
            b = cf(op.inputs[2])
            if o["padding"] == 0:
                pt, pb = _same_pad(a.shape[2], w.shape[2], o["stride_h"], o["dil_h"])
                pl, pr = _same_pad(a.shape[3], w.shape[3], o["stride_w"], o["dil_w"])
                a = F.pad(a, (pl, pr, pt, pb))
            y = F.conv2d(a, w, b, stride=(o["stride_h"], o["stride_w"]), dilation=(o["dil_h"], o["dil_w"]))
            y = _act(y, o["act"])
        elif k == "DEPTHWISE_CONV_2D":
            w = cf(op.inputs[1])                                # [1, kh, kw, C]
            c = w.shape[3]
            w = w.permute(3, 0, 1, 2)                           # [C, 1, kh, kw]
            b = cf(op.inputs[2])
            if o["padding"] == 0:
                pt, pb = _same_pad(a.shape[2], w.shape[2], o["stride_h"], o["dil_h"])
                pl, pr = _same_pad(a.shape[3], w.shape[3], o["stride_w"], o["dil_w"])
                a = F.pad(a, (pl, pr, pt, pb))
            y = F.conv2d(a, w, b, stride=(o["stride_h"], o["stride_w"]), dilation=(o["dil_h"], o["dil_w"]), groups=c)
            y = _act(y, o["act"])
        elif k == "AVERAGE_POOL_2D":
            assert o["fh"] == a.shape[2] and o["fw"] == a.shape[3], "only global pools occur"
            y = _act(a.mean(dim=(2, 3), keepdim=True), o["act"])
        elif k == "FULLY_CONNECTED":
            w = cf(op.inputs[1]); b = cf(op.inputs[2])
            v = a.permute(0, 2, 3, 1).reshape(-1, w.shape[1])
            y = _act(v @ w.t() + b, o["act"]).reshape(1, 1, 1, -1).permute(0, 3, 1, 2)
        elif k == "RESIZE_BILINEAR":
            oh, ow = [int(v) for v in g.tensors[op.inputs[1]].data]
            y = _resize(a, oh, ow, o["align_corners"], o["half_pixel"])
        elif k == "HARD_SWISH":
            y = a * torch.clamp(a + 3, 0, 6) / 6
        elif k == "LOGISTIC":
            y = torch.sigmoid(a)
        elif k == "RELU":
            y = torch.relu(a)
        elif k == "RELU6":
            y = torch.clamp(a, 0, 6)
        elif k == "ADD":
            y = _act(T[op.inputs[0]] + T[op.inputs[1]], o.get("act", "NONE"))
        elif k == "MUL":
            y = _act(T[op.inputs[0]] * T[op.inputs[1]], o.get("act", "NONE"))
        elif k == "CONCATENATION":
            y = torch.cat([T[i] for i in op.inputs], dim=1)
        elif k == "CUSTOM":
            assert o["custom"] == "Convolution2DTransposeBias" and o["stride_w"] == 2
            w = cf(op.inputs[1]).permute(3, 0, 1, 2)            # OHWI -> [I, O, kh, kw]
            y = F.conv_transpose2d(a, w, cf(op.inputs[2]), stride=2)
        else:
            raise NotImplementedError(k)
        T[op.outputs[0]] = y
    out = T[g.outputs[0]][0].permute(1, 2, 0).to(torch.float64).numpy()
    if keep:
        return out, {k: v[0].permute(1, 2, 0).to(torch.float64).numpy() for k, v in T.items()}, g
    return out
