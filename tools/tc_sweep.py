#!/usr/bin/env python
"""1x1-conv GEMM shapes of DeepLab / BodyPix at a real batch: exact FFMA kernels vs the tensor-core kernels
(ms per launch, useful TFLOP/s = 2*M*K*N / t), plus the accuracy of the tensor-core result against fp64.
    python tools/tc_sweep.py [batch]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

import backscrub_b200 as bs
from backscrub_b200 import api

L = bs.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
SHAPES = [(1089, 512, 256), (1089, 480, 160), (1089, 160, 480), (1089, 256, 256), (1089, 1280 // 4, 256), (1089, 256, 21), (1089, 96, 576 // 4 * 4),
          (4225, 144, 32), (4225, 32, 192), (16641, 16, 96), (16641, 96, 24)]
rng = np.random.default_rng(1)
for (m, K, N) in SHAPES:
    M = m * B
    row = {"M": M, "K": K, "N": N}
    flop = 2.0 * M * K * N
    t = L.bsb_time_pointwise(0, 0, M, K, N, 10)
    row["ffma_ms"] = round(t, 4); row["ffma_tflops"] = round(flop / t / 1e9, 1)
    for name, tv, mask in (("tc_r1", 1, 0), ("tc2", 2, 0), ("tc2_maskhi", 2, 1)):
        L.bsb_set_tuning(b"tc_variant", tv); L.bsb_set_tuning(b"tc_mask_hi", mask)
        t = L.bsb_time_pointwise(0, 1, M, K, N, 10)
        row[name + "_ms"] = round(t, 4); row[name + "_tflops"] = round(flop / t / 1e9, 1) if t > 0 else None
        # accuracy on one frame's worth of rows
        A = (rng.standard_normal((m, K)) * 2).astype(np.float32); W = (rng.standard_normal((N, K)) * 0.2).astype(np.float32)
        b = rng.standard_normal(N).astype(np.float32)
        try:
            got = api.pointwise(L, A, W, b, act=0, use_tc=True)
            ref = A.astype(np.float64) @ W.astype(np.float64).T + b
            scale = np.abs(A).astype(np.float64) @ np.abs(W).astype(np.float64).T + 1.0
            row[name + "_relerr"] = float((np.abs(got - ref) / scale).max())
        except Exception as e:
            row[name + "_relerr"] = str(e)[:60]
    L.bsb_set_tuning(b"tc_variant", 2); L.bsb_set_tuning(b"tc_mask_hi", 0)
    print(json.dumps(row), flush=True)
