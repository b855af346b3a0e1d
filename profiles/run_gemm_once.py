"""Mid-layer forward GEMM (lcn_layer_gemm, layer 2, B=4096) a few times, L2 flushed before each: the launch that
bench.py's roofline block times.  For `ncu --set full -k regex:k_tc_gemm` captures."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.gpu_helpers import make_pair, synth_xy, dev
from lcn_pose_b200 import _lib as L
eng, cfg, p = make_pair(L=3, knn=3, path=os.environ.get('LCN_BENCH_PATH', 'bf16'))
x, _ = synth_xy(4096)
xd = dev(x)
for _ in range(3): eng.forward(xd, bn_group=4096, training=True)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for _ in range(4):
    flush.zero_()
    L.check(eng.lib.lcn_layer_gemm(eng.h, eng.params.data_ptr(), eng.ws.data_ptr(), eng.ws.numel(), 4096, 4096, 2, 0, st))
torch.cuda.synchronize()
print("ok")
