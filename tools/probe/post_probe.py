"""debug aid: the smallest calls that reach k_post_tma (mask-only, then all outputs), for compute-sanitizer"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import backscrub_b200 as bs
from backscrub_b200 import api
from tests import synth
from tests.conftest import model_path
W, H = 640, 480
g = api.MaskGen(bs.lib(), model_path("mlkit"), W, H, max_batch=1)
fr = synth.frame(W, H, t=0)
print("mask-only", g.process(fr).mean(), flush=True)
g.set_background(synth.background())
out, yuyv, mask = g.composite(fr)
print("composite", out.mean(), yuyv.mean(), mask.mean(), flush=True)
