"""In-stream duration of every kernel of the train step (LCN_TRACE=1: CUDA events around each launch, eager mode,
warm caches as in the real step -- an ncu launch list flushes the caches before every kernel)."""
import os, sys
os.environ["LCN_TRACE"] = "1"
os.environ.setdefault("LCN_DISABLE_PDL", "1")        # event pairs measure whole kernels only without overlap
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.gpu_helpers import make_pair, synth_xy, dev
eng, cfg, p = make_pair(L=3, knn=3, path='bf16')
x, y = synth_xy(4096)
xd, yd = dev(x), dev(y)
for _ in range(5): eng.train_step(xd, yd, dropout=0.25)
torch.cuda.synchronize()
eng.lib.lcn_debug_trace_dump()
print("---- 20 steps ----", flush=True)
for _ in range(20): eng.train_step(xd, yd, dropout=0.25)
eng.lib.lcn_debug_trace_dump()
