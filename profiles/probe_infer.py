"""Throughput probe of the inference path (config 3 shape): forward at BN group 256 over a large resident
batch, and the evaluator.  Prints poses/s and the fraction of the bf16 tensor roofline."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lcn_pose_b200.engine import LcnEngine, eval_mpjpe  # noqa: E402
from lcn_pose_b200 import _lib as L  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
bn = int(sys.argv[2]) if len(sys.argv) > 2 else 256
reps = 5
nm = np.zeros((17, 17), np.float32)
L.load().lcn_neighbour_matrix(3, nm.ctypes.data_as(__import__("ctypes").c_void_p))
eng = LcnEngine(F=64, in_F=2, num_layers=3, neighbour_matrix=nm, path="bf16")
eng.init_params(42)
x = (torch.rand((n, 34), device="cuda") - 0.5)
out = torch.empty((n, 51), device="cuda")
for _ in range(2):
    eng.forward(x, bn_group=bn, out=out)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for i in range(reps):
    eng.forward(x, bn_group=bn, out=out)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = min(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
pps = n / ms * 1e3
print(json.dumps({"what": "forward", "n": n, "bn_group": bn, "ms": ms, "poses_per_s": pps,
                  "tensor_frac_burst": pps * 8713600 / 1652.9e12}))
gt = torch.randn((n, 17, 3), device="cuda") * 300 + torch.tensor([0., 0., 4500.], device="cuda")
pred = out.view(n, 17, 3).contiguous()
box = torch.tensor([0., 0., 999., 999.], device="cuda").repeat(n, 1)
cam = torch.tensor([1145., 1143., 512., 515.], device="cuda").repeat(n, 1)
rd = gt[:, 0, 2].contiguous()
for p2 in (False, True):
    for want in (False, True):
        eval_mpjpe(pred, gt, box, cam, rd, p2, want_err=want)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            eval_mpjpe(pred, gt, box, cam, rd, p2, want_err=want)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        byt = 444 + (68 if want else 0)
        print(json.dumps({"what": "eval", "protocol2": p2, "want_err": want, "ms": ms, "poses_per_s": n / ms * 1e3,
                          "hbm_gbs": n * byt / ms / 1e6, "hbm_frac": n * byt / ms / 1e6 / 6551.0}))
