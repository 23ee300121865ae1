"""How far the tensor-core (3xTF32) pointwise path is from the exact path where it matters: per-frame decision flips
(ofinal), mask differences, and logit error, on the synthetic fixtures the parity tests use.
    python tools/probe/tc_flips.py [tc_min_k ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

import backscrub_b200 as bs
from backscrub_b200 import api
from tests import synth
from tests.conftest import model_path

L = bs.lib()
mins = [int(x) for x in sys.argv[1:]] or [160, 16]
CASES = [("deeplab", 640, 480, 6), ("deeplab", 1280, 720, 4), ("bodypix", 640, 480, 6), ("bodypix", 1920, 1080, 3), ("bodypix", 3840, 2160, 2)]
bg = synth.background()
for mk in mins:
    L.bsb_set_tuning(b"tc_min_k", mk)
    for key, W, H, n in CASES:
        for kind in ("person", "noise"):
            frames = np.stack([synth.frame(W, H, t=t, kind=kind) for t in range(n)])
            res = {}
            for name, flags in (("exact", 0), ("tc", 4)):
                g = api.MaskGen(L, model_path(key), W, H, max_batch=n, flags=flags)
                g.set_background(bg)
                out, yuyv, mask = g.composite(frames)
                res[name] = (np.stack([g.stage_u8(2, b) for b in range(n)]), mask, out)
                g.close()
            of_e, m_e, o_e = res["exact"]; of_t, m_t, o_t = res["tc"]
            print(json.dumps({"tc_min_k": mk, "model": key, "W": W, "H": H, "frames": n, "kind": kind,
                              "ofinal_pixels_differing": int((of_e != of_t).sum()), "mask_bytes_differing": int((m_e != m_t).sum()),
                              "out_bytes_differing": int((o_e != o_t).sum()), "person_fraction": float((m_e < 128).mean())}), flush=True)
L.bsb_set_tuning(b"tc_min_k", 160)
