"""Eager vs CUDA-graph train steps (and eager vs eager): per-tensor count / size of parameter differences after 4 steps.
Arrival-order fp32 reductions in the weight-gradient kernels make single elements with a near-cancelling gradient take
a different Adam direction; anything beyond a handful of such elements would be a real ordering bug."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.gpu_helpers import make_pair, synth_xy, dev
n = 384
x, y = synth_xy(n)
xd, yd = dev(x), dev(y)
for path in ("bf16", "fp32"):
    for mode in ("eager-eager", "eager-graph", "graph-graph"):
        a, _, _ = make_pair(L=1, knn=2, path=path)
        b, _, _ = make_pair(L=1, knn=2, path=path)
        fa = a.train_step if mode.startswith("eager") else a.train_step_graph
        fb = b.train_step if mode.endswith("eager") else b.train_step_graph
        for s in range(4):
            la = float(fa(xd, yd, dropout=0.25)[0].item())
            lb = float(fb(xd, yd, dropout=0.25)[0].item())
            if s == 0:
                ga, gb = a.grads_raw.clone(), b.grads_raw.clone()
        d = (a.params - b.params).abs()
        tol = 1e-6 + 1e-4 * b.params.abs()
        bad = d > tol
        dg = (ga - gb).abs()
        print(f"{path} {mode}: loss {la:.8f} {lb:.8f}  params: max diff {d.max().item():.3e}, {int(bad.sum())} of {d.numel()} beyond "
              f"tol; step-1 raw grads: max diff {dg.max().item():.3e} (max |g| {ga.abs().max().item():.3e}), "
              f"{int((dg > 1e-5 * ga.abs().max()).sum())} elements differ by more than 1e-5 of max")
        if bad.any():
            idx = bad.nonzero().flatten()[:8].tolist()
            for i in idx:
                print(f"   elem {i}: a {a.params[i].item():.6e} b {b.params[i].item():.6e} g1a {ga[i].item():.3e} g1b {gb[i].item():.3e}")
