"""Minimal .tflite (schema v3, "TFL3") flatbuffer reader -> python graph description.

Test/inspection tooling only (not on the product path).  The field numbering
follows the on-disk format documented in the reference's vendored schema
(tensorflow/lite/schema/schema.fbs: Model :1231, SubGraph :1169, Tensor :195,
Buffer :1191, OperatorCode :1108, Operator :1134; option tables :510-:721).
Used by tests to (a) cross-check the C oracle's loader and (b) evaluate the
graph with torch CPU ops as an independent restatement.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

BUILTIN = {0: "ADD", 1: "AVERAGE_POOL_2D", 2: "CONCATENATION", 3: "CONV_2D",
           4: "DEPTHWISE_CONV_2D", 6: "DEQUANTIZE", 9: "FULLY_CONNECTED",
           14: "LOGISTIC", 18: "MUL", 19: "RELU", 21: "RELU6",
           23: "RESIZE_BILINEAR", 32: "CUSTOM", 117: "HARD_SWISH"}
ACT = {0: "NONE", 1: "RELU", 2: "RELU_N1_TO_1", 3: "RELU6"}
TTYPE = {0: "f32", 1: "f16", 2: "i32"}


class FB:
    """Tiny flatbuffer accessor."""

    def __init__(self, buf: bytes):
        self.b = buf

    def u8(self, o): return self.b[o]
    def i8(self, o): return struct.unpack_from("<b", self.b, o)[0]
    def u16(self, o): return struct.unpack_from("<H", self.b, o)[0]
    def i32(self, o): return struct.unpack_from("<i", self.b, o)[0]
    def u32(self, o): return struct.unpack_from("<I", self.b, o)[0]

    def indirect(self, o): return o + self.u32(o)

    def field(self, table, slot):
        """Offset of field `slot` inside `table`, or 0 when absent."""
        vt = table - self.i32(table)
        vtsize = self.u16(vt)
        fo = 4 + 2 * slot
        if fo >= vtsize:
            return 0
        off = self.u16(vt + fo)
        return table + off if off else 0

    def vec(self, table, slot):
        f = self.field(table, slot)
        if not f:
            return 0, 0
        v = self.indirect(f)
        return v + 4, self.u32(v)

    def scalar(self, table, slot, kind, default=0):
        f = self.field(table, slot)
        if not f:
            return default
        return getattr(self, kind)(f)

    def string(self, table, slot):
        f = self.field(table, slot)
        if not f:
            return ""
        s = self.indirect(f)
        n = self.u32(s)
        return self.b[s + 4:s + 4 + n].decode("utf-8", "replace")

    def table_vec(self, table, slot):
        start, n = self.vec(table, slot)
        return [self.indirect(start + 4 * i) for i in range(n)]

    def int_vec(self, table, slot):
        start, n = self.vec(table, slot)
        return list(struct.unpack_from("<%di" % n, self.b, start)) if n else []


@dataclass
class Tensor:
    idx: int
    name: str
    shape: list
    dtype: str
    buffer: int
    data: np.ndarray | None = None     # constant data (as stored), None for activations


@dataclass
class Op:
    idx: int
    kind: str
    inputs: list
    outputs: list
    opts: dict = field(default_factory=dict)


@dataclass
class Graph:
    tensors: list
    ops: list
    inputs: list
    outputs: list

    def const_f32(self, t: int) -> np.ndarray:
        """Constant tensor as float32; follows a DEQUANTIZE producer (fp16 weight storage)."""
        ten = self.tensors[t]
        if ten.data is not None:
            return ten.data.astype(np.float32).reshape(ten.shape)
        for op in self.ops:
            if op.kind == "DEQUANTIZE" and op.outputs[0] == t:
                src = self.tensors[op.inputs[0]]
                return src.data.astype(np.float32).reshape(src.shape)
        raise KeyError(f"tensor {t} ({ten.name}) is not constant")

    def is_const(self, t: int) -> bool:
        if self.tensors[t].data is not None:
            return True
        return any(op.kind == "DEQUANTIZE" and op.outputs[0] == t and
                   self.tensors[op.inputs[0]].data is not None for op in self.ops)


def load(path: str) -> Graph:
    buf = open(path, "rb").read()
    fb = FB(buf)
    assert buf[4:8] == b"TFL3", "not a TFL3 flatbuffer"
    model = fb.indirect(0)
    opcodes = []
    for oc in fb.table_vec(model, 1):
        dep = fb.scalar(oc, 0, "i8", 0)
        new = fb.scalar(oc, 3, "i32", 0)
        code = max(dep, new)
        opcodes.append((code, fb.string(oc, 1)))
    buffers = []
    for bt in fb.table_vec(model, 4):
        start, n = fb.vec(bt, 0)
        buffers.append((start, n))
    sg = fb.table_vec(model, 2)[0]
    tensors = []
    for i, tt in enumerate(fb.table_vec(sg, 0)):
        shape = fb.int_vec(tt, 0)
        ty = fb.scalar(tt, 1, "i8", 0)
        bidx = fb.scalar(tt, 2, "u32", 0)
        name = fb.string(tt, 3)
        data = None
        start, n = buffers[bidx]
        if n:
            dt = {0: np.float32, 1: np.float16, 2: np.int32}[ty]
            data = np.frombuffer(buf, dtype=dt, count=n // np.dtype(dt).itemsize, offset=start).copy()
        tensors.append(Tensor(i, name, shape, TTYPE.get(ty, str(ty)), bidx, data))
    ops = []
    for i, ot in enumerate(fb.table_vec(sg, 3)):
        code, custom = opcodes[fb.scalar(ot, 0, "u32", 0)]
        kind = BUILTIN.get(code, f"OP{code}")
        ins = fb.int_vec(ot, 1)
        outs = fb.int_vec(ot, 2)
        opts = {}
        bo = fb.field(ot, 4)
        bo = fb.indirect(bo) if bo else 0
        if kind == "CONV_2D" and bo:
            opts = dict(padding=fb.scalar(bo, 0, "i8"), stride_w=fb.scalar(bo, 1, "i32"),
                        stride_h=fb.scalar(bo, 2, "i32"), act=ACT[fb.scalar(bo, 3, "i8")],
                        dil_w=fb.scalar(bo, 4, "i32", 1), dil_h=fb.scalar(bo, 5, "i32", 1))
        elif kind == "DEPTHWISE_CONV_2D" and bo:
            opts = dict(padding=fb.scalar(bo, 0, "i8"), stride_w=fb.scalar(bo, 1, "i32"),
                        stride_h=fb.scalar(bo, 2, "i32"), mult=fb.scalar(bo, 3, "i32"),
                        act=ACT[fb.scalar(bo, 4, "i8")],
                        dil_w=fb.scalar(bo, 5, "i32", 1), dil_h=fb.scalar(bo, 6, "i32", 1))
        elif kind == "AVERAGE_POOL_2D" and bo:
            opts = dict(padding=fb.scalar(bo, 0, "i8"), stride_w=fb.scalar(bo, 1, "i32"),
                        stride_h=fb.scalar(bo, 2, "i32"), fw=fb.scalar(bo, 3, "i32"),
                        fh=fb.scalar(bo, 4, "i32"), act=ACT[fb.scalar(bo, 5, "i8")])
        elif kind == "RESIZE_BILINEAR":
            opts = dict(align_corners=bool(fb.scalar(bo, 2, "u8")) if bo else False,
                        half_pixel=bool(fb.scalar(bo, 3, "u8")) if bo else False)
        elif kind == "FULLY_CONNECTED" and bo:
            opts = dict(act=ACT[fb.scalar(bo, 0, "i8")], keep_num_dims=bool(fb.scalar(bo, 2, "u8")))
        elif kind in ("ADD", "MUL") and bo:
            opts = dict(act=ACT[fb.scalar(bo, 0, "i8")])
        elif kind == "CONCATENATION" and bo:
            opts = dict(axis=fb.scalar(bo, 0, "i32"), act=ACT[fb.scalar(bo, 1, "i8")])
        elif kind == "CUSTOM":
            start, n = fb.vec(ot, 5)
            raw = buf[start:start + n]
            opts = dict(custom=custom)
            if custom == "Convolution2DTransposeBias" and n >= 12:
                p, sw, sh = struct.unpack_from("<iii", raw, 0)
                opts.update(padding_c_enum=p, stride_w=sw, stride_h=sh)
        ops.append(Op(i, kind, ins, outs, opts))
    return Graph(tensors, ops, fb.int_vec(sg, 1), fb.int_vec(sg, 2))


def dump(path: str) -> None:
    g = load(path)
    T = g.tensors
    print(f"# {path}: {len(T)} tensors, {len(g.ops)} ops, in={g.inputs} out={g.outputs}")
    for op in g.ops:
        if op.kind == "DEQUANTIZE":
            continue
        ins = []
        for t in op.inputs:
            if t < 0:
                ins.append("-")
            else:
                c = "c" if g.is_const(t) else "t"
                ins.append(f"{c}{t}{T[t].shape}")
        outs = [f"t{t}{T[t].shape}" for t in op.outputs]
        o = {k: v for k, v in op.opts.items() if not (k.startswith("dil") and v == 1)}
        print(f"{op.idx:4d} {op.kind:20s} {' '.join(ins)} -> {' '.join(outs)} {o}")


if __name__ == "__main__":
    import sys
    for p in sys.argv[1:]:
        dump(p)
