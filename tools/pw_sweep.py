"""Time the exact FFMA pointwise kernels per layer shape on the GPU (old heuristics = variant 2, register-tiled = 3).
    python tools/pw_sweep.py [batch]
"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import backscrub_b200 as bs  # noqa: E402
from tests.conftest import MODELS, model_path  # noqa: E402
from tools import tflite_graph as tg  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
lib = bs.lib()
seen = collections.OrderedDict()
for key in MODELS:
    g = tg.load(model_path(key))
    for op in g.ops:
        if op.kind == "CONV_2D":
            w = g.tensors[op.inputs[1]].shape
            o = g.tensors[op.outputs[0]].shape
            if w[1] == 1 and w[2] == 1 and o[1] * o[2] > 1:
                seen.setdefault((o[1] * o[2], w[3], w[0]), []).append(key)
# variants: 2 = rows/classic heuristics, 16/32/64 = classic kernel with that N tile, 4 / 8 = register-tiled 8x4 / 8x8
print(f"batch {batch}: rows/frame K N | heur ms | rows | bn16 | bn32 | bn64 | tile4 | tile8 | best TFLOP/s | models")
for (rows, K, N), keys in sorted(seen.items(), key=lambda kv: -kv[0][0] * kv[0][1] * kv[0][2]):
    M = rows * batch
    if M < 4096:
        continue
    ts = [lib.bsb_time_pointwise(0, v, M, K, N, 20) for v in (0, 5, 16, 32, 64, 4, 8)]
    fl = 2.0 * M * K * N
    print(f"{rows:6d} {K:4d} {N:4d} | " + " | ".join(f"{t:7.4f}" for t in ts) + f" | {fl / min(ts) / 1e9:7.2f} | {','.join(sorted(set(keys)))}")
