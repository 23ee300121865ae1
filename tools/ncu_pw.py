"""Small driver for ncu captures of single kernels through the C ABI (no bench harness around them)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import backscrub_b200 as bs
from backscrub_b200 import api
lib = bs.lib()
what = sys.argv[1]
if what == "tile":
    for (M, K, N) in [(34848, 512, 256), (34848, 480, 160)]:
        print(M, K, N, lib.bsb_time_pointwise(0, 3, M, K, N, 2))
elif what == "gauss":
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8)
    out = api.gaussian_blur(lib, img, 25)
    print(out.mean())
