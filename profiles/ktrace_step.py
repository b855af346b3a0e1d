"""Start-to-start timeline of one train step as replayed from the CUDA graph (profiling build, -DLCN_KTRACE): block (0,0)
of every kernel stamps %globaltimer right after its griddepcontrol.wait.  Run with
LCN_B200_LIB=lcn_pose_b200/liblcn_b200_prof.so python profiles/ktrace_step.py"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.gpu_helpers import make_pair, synth_xy, dev
eng, cfg, p = make_pair(L=3, knn=3, path='bf16')
x, y = synth_xy(4096)
xd, yd = dev(x), dev(y)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def read(reset):
    rows = []
    for tu in ("kernels", "gemm"):
        fn = getattr(eng.lib, "lcn_ktrace_read_" + tu)
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        buf = (C.c_ulonglong * (2 * 8192))(); n = C.c_uint(0)
        assert fn(buf, C.byref(n), reset) == 0
        a = np.array(buf[:2 * min(n.value, 8192)], dtype=np.uint64).reshape(-1, 2)
        rows += [(int(t), tu, int(s) >> 32, (int(s) >> 20) & 0xfff, (int(s) >> 8) & 0xfff, int(s) & 0xff) for t, s in a]
    return sorted(rows)
for _ in range(5): eng.train_step_graph(xd, yd, 0.25)
torch.cuda.synchronize(); read(1)
flush.zero_(); torch.cuda.synchronize()
eng.train_step_graph(xd, yd, 0.25)
torch.cuda.synchronize()
rows = read(1)
t0 = rows[0][0]
print("n kernels", len(rows))
prev = t0
for t, tu, gx, gy, bx, by in rows:
    print(f"{(t - t0) / 1e3:9.2f} us  +{(t - prev) / 1e3:7.2f}  {tu:8s} grid=({gx},{gy}) block=({bx},{by})")
    prev = t
