"""Cross-check the oracle's hand-derived backward and TF1 Adam against torch autograd."""
import math

import numpy as np
import pytest
import torch

from oracle import lcn_oracle as O
from oracle import torch_restatement as T

CASES = [
    dict(mask_type="locally_connected", knn=3, F=8, L=2, B=12, drop=0.0, reg=0.0),
    dict(mask_type="locally_connected", knn=1, F=4, L=1, B=7, drop=0.25, reg=0.0),
    dict(mask_type="exponential", knn=2, F=8, L=1, B=9, drop=0.0, reg=5e-4),
    dict(mask_type="locally_connected", knn=2, F=16, L=3, B=5, drop=0.3, reg=None),
]


def _setup(c, seed=0):
    cfg = O.LcnConfig(F=c["F"], num_layers=c["L"], mask_type=c["mask_type"],
                      neighbour_matrix=O.get_neighbour_matrix_by_hand(knn=c["knn"]), regularization=c["reg"])
    p = O.init_params(cfg, seed=seed)
    rng = np.random.default_rng(seed + 1)
    # perturb so that nothing sits at a symmetric point (gamma=1, beta=0, var_mask=0/1)
    for k in p:
        if k.endswith("gamma") or k.endswith("beta") or k == "mask":
            p[k] = p[k] + rng.normal(0, 0.1, p[k].shape)
    x = rng.normal(0, 0.3, (c["B"], 34))
    y = rng.normal(0, 0.3, (c["B"], 51))
    keep = None
    if c["drop"] > 0:
        keep = [rng.uniform(size=(c["B"], 17 * c["F"])) >= c["drop"] for _ in range(1 + 2 * c["L"])]
    return cfg, p, x, y, keep


@pytest.mark.parametrize("c", CASES)
def test_backward_matches_autograd(c):
    cfg, p, x, y, keep = _setup(c)
    loss, g = O.loss_and_grads(cfg, p, x, y, c["drop"], keep)
    tp = T.build_params(p)
    tk = None if keep is None else [torch.tensor(k, dtype=torch.float64) for k in keep]
    tl, tout = T.loss_fn(cfg, tp, torch.tensor(x), torch.tensor(y), c["drop"], tk)
    tl.backward()
    assert abs(loss - tl.item()) < 1e-13 * max(1, abs(loss))
    out, _ = O.forward(cfg, p, x, c["drop"], keep)
    np.testing.assert_allclose(out, tout.detach().numpy(), atol=1e-12)
    assert set(g) == set(tp)
    for k in g:
        ref = tp[k].grad.numpy()
        scale = max(np.abs(ref).max(), 1e-300)
        assert np.abs(g[k] - ref).max() / scale < 1e-9, k


def test_clip_norm_gradient_leaks_to_masked_out_weights():
    """masked-out entries get a (tiny) gradient through the norm."""
    cfg, p, x, y, _ = _setup(CASES[0])
    _, g = O.loss_and_grads(cfg, p, x, y)
    name = O.weight_names(cfg)[1]
    sup = np.kron(cfg.support(), np.ones((cfg.F, cfg.F))) != 0
    assert np.abs(g[name][~sup]).max() > 0
    assert np.abs(g[name][~sup]).mean() < np.abs(g[name][sup]).mean()


def test_adam_matches_tf1_formula():
    cfg, p, x, y, _ = _setup(CASES[0])
    st = O.AdamState()
    p0 = {k: v.copy() for k, v in p.items()}
    _, g1 = O.loss_and_grads(cfg, p, x, y)
    O.adam_step(cfg, p, g1, st)
    # step 1 closed form: m = .1 g, v = .001 g^2, lr_t = lr*sqrt(.001)/.1
    lr_t = cfg.learning_rate * math.sqrt(1 - 0.999) / (1 - 0.9)
    for k in p:
        exp = p0[k] - lr_t * 0.1 * g1[k] / (np.sqrt(0.001 * g1[k] ** 2) + 1e-8)
        np.testing.assert_allclose(p[k], exp, rtol=1e-12, atol=1e-18)
    # lr schedule uses global_step = t-1
    assert O.learning_rate_at(cfg, 1) == cfg.learning_rate
    assert abs(O.learning_rate_at(cfg, 32001) - cfg.learning_rate * 0.96) < 1e-18
    # a second step moves again and stays finite
    _, g2 = O.loss_and_grads(cfg, p, x, y)
    O.adam_step(cfg, p, g2, st)
    assert st.t == 2 and all(np.isfinite(v).all() for v in p.values())
