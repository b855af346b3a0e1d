"""Per-phase timeline (SM clock cycles) of the fused inference kernel for cluster 0 / CTA 0, second group.
Needs a library built with EXTRA=-DLCN_TC_PROFILE."""
import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcn_pose_b200.engine import LcnEngine
from lcn_pose_b200 import _lib as L
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256 * 148
bn = int(sys.argv[2]) if len(sys.argv) > 2 else 256
eng = LcnEngine(F=64, in_F=2, num_layers=3, neighbour_matrix=L.neighbour_matrix(3), path="bf16")
eng.init_params(42)
x = torch.rand((n, 34), device="cuda") - 0.5
eng.forward(x, bn_group=bn)
torch.cuda.synchronize()
buf = (C.c_ulonglong * 512)()
fn = eng.lib.lcn_debug_read_stack_prof
fn.argtypes = [C.c_void_p]
assert fn(buf) == 0
t = np.array(buf[:], dtype=np.int64)
names = ["epi_start", "tfull", "pass1", "stats_xchg", "pass2", "stored", "layer_end", "-", "mma_k0", "mma_k8", "mma_k16"]
t0 = t[0]
for l in range(8):
    row = t[16 * l: 16 * l + 11]
    print(f"   mma wait-full total {int(t[16*l+11])}  producer wait-empty total {int(t[16*l+12])}  mma issue {int(t[16*l+13])} commit {int(t[16*l+14])}  producer iter total {int(t[16*l+15])}")
    print(f"layer {l}: " + "  ".join(f"{nm}={int(v - t0)}" for nm, v in zip(names, row) if v > 0 and nm != "-"))
