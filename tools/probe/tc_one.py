"""one tensor-core GEMM launch loop for ncu: python tools/probe/tc_one.py M K N"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import backscrub_b200 as bs
L = bs.lib()
M, K, N = (int(x) for x in sys.argv[1:4])
print(L.bsb_time_pointwise(0, 1, M, K, N, 3))
