"""Micro-probe of the pointwise kernels (run under ncu): one tensor-core and one FFMA launch per shape."""
import sys
import numpy as np
sys.path.insert(0, ".")
import backscrub_b200 as bs
from backscrub_b200 import api
L = bs.lib()
rng = np.random.default_rng(0)
for (M, K, N) in [(17424, 512, 256), (17424, 480, 160), (17424, 160, 256), (17424, 80, 480)]:
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) * 0.1).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    for tc in (True, False):
        out = api.pointwise(L, A, W, b, act=3, use_tc=tc)
    print(M, K, N, "ok", float(out.mean()))
