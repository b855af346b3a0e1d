"""Shared helpers of the GPU parity tests: build an (engine, oracle) pair on the same seeded inputs."""
import numpy as np
import torch

from lcn_pose_b200.engine import LcnEngine
from oracle import lcn_oracle as O


def make_pair(F=64, L=1, knn=3, mask_type="locally_connected", path="fp32", seed=0, perturb=True, reg=0.0):
    nm = O.get_neighbour_matrix_by_hand(knn=knn)
    cfg = O.LcnConfig(F=F, num_layers=L, mask_type=mask_type, neighbour_matrix=nm, regularization=reg)
    p = O.init_params(cfg, seed=seed, dtype=np.float32)     # fp32-representable parameters on both sides
    rng = np.random.default_rng(seed + 100)
    if perturb:
        for k in p:
            if k.endswith("gamma") or k.endswith("beta") or k == "mask":
                p[k] = (p[k] + rng.normal(0, 0.1, p[k].shape)).astype(np.float32)
    eng = LcnEngine(F=F, in_F=2, num_layers=L, mask_type=mask_type, neighbour_matrix=nm, path=path,
                    regularization=reg)
    eng.set_params(p)
    p64 = {k: v.astype(np.float64) for k, v in p.items()}
    return eng, cfg, p64


def synth_xy(n, seed=1234):
    """H36M-shaped synthetic 2D inputs / 3D labels."""
    rng = np.random.default_rng(seed)
    root = rng.uniform(-0.5, 0.5, (n, 1, 2))
    x = np.clip(root + rng.normal(0, 0.15, (n, 17, 2)), -1, 1)
    y = np.concatenate([x + rng.normal(0, 0.02, (n, 17, 2)), rng.normal(0, 0.1, (n, 17, 1))], axis=2)
    return x.reshape(n, 34).astype(np.float32), y.reshape(n, 51).astype(np.float32)


def rel_err(a, b):
    """max |a-b| / max |b|"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def elem_rel_err(a, b, floor=1e-3):
    """Elementwise relative error next to the max-norm rel_err: max_i |a_i - b_i| / (|b_i| + floor * max|b|).  The floor
    keeps entries that are zero up to rounding (|b_i| << max|b|) from dominating; with floor = 1e-3 an entry 1000 times
    smaller than the largest one is still held to the stated relative tolerance within a factor of two."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float((np.abs(a - b) / (np.abs(b) + floor * max(np.abs(b).max(), 1e-300))).max())


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()
