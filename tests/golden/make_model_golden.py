"""Model-level golden vectors of the oracle ("generate with fixed seeds and commit as .npz").

    python tests/golden/make_model_golden.py

TensorFlow is not installable here, so these are outputs of the RESTATED reference (oracle/lcn_oracle.py), not of
the reference itself: their job is to detect drift of the oracle -- every GPU parity claim is made against it -- not to
pin it to TensorFlow ("parity unpinned" for the model).  Cases: the three mask configurations the
benchmarks use (knn=3 locally connected L=3 = configs[1]; exponential L=3 = configs[3]; knn=1 F=128 L=1 as the small
stand-in of configs[4]), forward + loss + every gradient + three TF1-Adam steps with dropout 0.25 from an injected keep
mask and (one case) regularization 5e-4.  Stored compactly: full predictions, and for every gradient / updated
parameter tensor its L2 norm, its sum and 64 entries at seeded positions.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import lcn_oracle as O  # noqa: E402

CASES = {
    "knn3_L3": dict(F=64, L=3, knn=3, mask_type="locally_connected", n=96, reg=0.0),
    "exp_L3": dict(F=64, L=3, knn=1, mask_type="exponential", n=64, reg=0.0),
    "knn1_F128_L1_reg": dict(F=128, L=1, knn=1, mask_type="locally_connected", n=48, reg=5e-4),
}


def synth_xy(n, seed=1234):
    rng = np.random.default_rng(seed)
    root = rng.uniform(-0.5, 0.5, (n, 1, 2))
    x = np.clip(root + rng.normal(0, 0.15, (n, 17, 2)), -1, 1)
    y = np.concatenate([x + rng.normal(0, 0.02, (n, 17, 2)), rng.normal(0, 0.1, (n, 17, 1))], axis=2)
    return x.reshape(n, 34), y.reshape(n, 51)


def fingerprint(a, seed):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    idx = np.random.default_rng(seed).integers(0, a.size, 64)
    return np.concatenate([[np.sqrt((a * a).sum()), a.sum()], a[idx]])


def run_case(c):
    cfg = O.LcnConfig(F=c["F"], num_layers=c["L"], mask_type=c["mask_type"],
                      neighbour_matrix=O.get_neighbour_matrix_by_hand(knn=c["knn"]), regularization=c["reg"])
    p = {k: v.astype(np.float64) for k, v in O.init_params(cfg, seed=7, dtype=np.float32).items()}
    rng = np.random.default_rng(11)
    for k in p:
        if k.endswith("gamma") or k.endswith("beta") or k == "mask":
            p[k] = (p[k] + rng.normal(0, 0.1, p[k].shape)).astype(np.float32).astype(np.float64)
    x, y = synth_xy(c["n"])
    rate = 0.25
    keep = [rng.random((c["n"], 17 * c["F"])) >= rate for _ in range(1 + 2 * c["L"])]
    out = {}
    out["pred"] = O.forward(cfg, p, x)[0]
    out["pred_dropout"] = O.forward(cfg, p, x, rate, keep)[0]
    loss, grads = O.loss_and_grads(cfg, p, x, y, rate, keep)
    out["loss"] = np.float64(loss)
    for i, k in enumerate(sorted(grads)):
        out["grad/" + k] = fingerprint(grads[k], 100 + i)
    st = O.AdamState()
    losses = []
    for _ in range(3):
        l, lr, _ = O.train_step(cfg, p, st, x, y, rate, keep)
        losses.append(l)
    out["adam_losses"] = np.array(losses)
    for i, k in enumerate(sorted(p)):
        out["param3/" + k] = fingerprint(p[k], 200 + i)
    return out


def main():
    blob = {}
    for name, c in CASES.items():
        for k, v in run_case(c).items():
            blob[name + "/" + k] = v
    path = os.path.join(HERE, "model_oracle.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, len(blob), "arrays,", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
