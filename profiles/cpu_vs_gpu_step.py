"""Is the train step CPU-launch-bound?  Wall time to ENQUEUE n steps vs time until the GPU has finished them."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.gpu_helpers import make_pair, synth_xy, dev
eng, cfg, p = make_pair(L=3, knn=3, path='bf16')
x, y = synth_xy(4096)
xd, yd = dev(x), dev(y)
for _ in range(5): eng.train_step(xd, yd, dropout=0.25)
torch.cuda.synchronize()
n = 50
for part in ("forward", "backward", "adam", "step"):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        if part in ("forward", "step"): eng.forward(xd, bn_group=4096, training=True, dropout=0.25)
        if part in ("backward", "step"): eng.backward(xd, yd, 0.25)
        if part in ("adam", "step"): eng.adam()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{part:9s}: enqueue {1e6 * (t1 - t0) / n:8.1f} us/iter   until GPU done {1e6 * (t2 - t0) / n:8.1f} us/iter")
