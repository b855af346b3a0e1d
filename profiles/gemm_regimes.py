"""Duration of the mid-layer forward GEMM (lcn_layer_gemm, B=4096) by CUDA events in different neighbourhoods:
back to back, after a tiny kernel, after an L2 flush, after a big elementwise kernel."""
import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.gpu_helpers import make_pair, synth_xy, dev
from lcn_pose_b200 import _lib as L
eng, cfg, p = make_pair(L=3, knn=3, path='bf16')
x, _ = synth_xy(4096)
xd = dev(x)
for _ in range(3): eng.forward(xd, bn_group=4096, training=True)
torch.cuda.synchronize()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
tiny = torch.zeros(32, device='cuda')
big = torch.zeros(8 << 20, device='cuda')
def gemm():
    L.check(eng.lib.lcn_layer_gemm(eng.h, eng.params.data_ptr(), eng.ws.data_ptr(), eng.ws.numel(), 4096, 4096, 2, 0, st))
def timed(pre, reps=30):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        pre()
        a.record(); gemm(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) * 1e3 for a, b in ev)
    return round(ts[0], 1), round(ts[len(ts) // 2], 1), round(ts[-1], 1)
for name, pre in (("back-to-back", lambda: None), ("after tiny kernel", lambda: tiny.add_(1)), ("after 32MB elementwise", lambda: big.add_(1)),
                  ("after 256MB flush", lambda: flush.zero_()), ("back-to-back again", lambda: None)):
    print(f"{name:26s} us (min, median, max): {timed(pre)}")
