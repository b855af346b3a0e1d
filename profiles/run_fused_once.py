"""One fused-inference call on synthetic data (for ncu captures): python profiles/run_fused_once.py [n] [bn_group]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcn_pose_b200.engine import LcnEngine
from lcn_pose_b200 import _lib as L
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256 * 22 * 8
bn = int(sys.argv[2]) if len(sys.argv) > 2 else 256
eng = LcnEngine(F=64, in_F=2, num_layers=3, neighbour_matrix=L.neighbour_matrix(3), path="bf16")
eng.init_params(42)
x = torch.rand((n, 34), device="cuda") - 0.5
for _ in range(3):
    out = eng.forward(x, bn_group=bn)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
