"""Per-layer / per-group / per-joint error map of the fused inference kernel against the oracle (debug aid)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import lcn_oracle as O
from tests.gpu_helpers import dev, make_pair, synth_xy
from tests.test_gpu_fused import _layer_refs
L, knn, n, bn = 3, 3, int(sys.argv[1]) if len(sys.argv) > 1 else 512, 256
eng, cfg, p = make_pair(L=L, knn=knn, path="bf16")
x, _ = synth_xy(n)
out, taps = eng.forward_taps(dev(x), bn)
taps = taps.cpu().numpy().astype(np.float64)
ng = (n + bn - 1) // bn
xpad = np.zeros((ng * bn, 34)); xpad[:n] = x
a_prev = xpad
for l in range(1 + 2 * L):
    res = taps[l - 2] if (l >= 2 and l % 2 == 0) else None
    ref = _layer_refs(cfg, p, a_prev, l, bn, res)
    d = np.abs(taps[l] - ref).reshape(ng, bn, 17, 64)
    sc = np.abs(ref).max()
    print(f"layer {l}: max rel {d.max() / sc:.4f}")
    for g in range(ng):
        per_j = d[g].max(axis=(0, 2)) / sc
        rows = d[g].max(axis=(1, 2)) / sc
        if per_j.max() > 1e-2:
            print(f"   group {g}: bad joints {np.nonzero(per_j > 1e-2)[0].tolist()}  bad rows {int((rows > 1e-2).sum())} first {np.nonzero(rows > 1e-2)[0][:8].tolist()}")
    a_prev = taps[l]
# affine fit per (joint, channel) for layer 0 group 1: is the error a per-column scale/shift (statistics) or random (GEMM)?
ref = _layer_refs(cfg, p, xpad, 0, bn, None).reshape(ng, bn, 17, 64)
got = taps[0].reshape(ng, bn, 17, 64)
for g in range(ng):
    for j in (0, 5, 16):
        for f in (0, 1, 33):
            r_, g_ = ref[g, :, j, f], got[g, :, j, f]
            neg = r_ < 0
            # undo leaky relu to compare the BN output
            r2 = np.where(neg, r_ / 0.2, r_); g2 = np.where(g_ < 0, g_ / 0.2, g_)
            A = np.vstack([r2, np.ones_like(r2)]).T
            sl, ic = np.linalg.lstsq(A, g2, rcond=None)[0]
            resid = np.abs(g2 - (sl * r2 + ic)).max()
            print(f"g{g} j{j} f{f}: slope {sl:.4f} icpt {ic:+.4f} resid {resid:.4f}  first rows got {g_[:3]} ref {r_[:3]}")
# which reference row does each produced row of group 1 match?
g = 1
R = ref[g].reshape(bn, -1); Gt = got[g].reshape(bn, -1)
d2 = ((Gt[:, None, :] - R[None, :, :]) ** 2).sum(-1)
match = d2.argmin(1)
print("row -> matched ref row:", match[:16].tolist(), "...", match[60:70].tolist(), "...", match[126:134].tolist(), "...", match[250:256].tolist())
print("fraction identity", float((match == np.arange(bn)).mean()), " match^row unique:", np.unique(match ^ np.arange(bn))[:10].tolist())
