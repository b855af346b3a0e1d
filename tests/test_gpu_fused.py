"""GPU parity tests of the fused inference kernel (lcn_pose_b200/csrc/lcn_stack_tc.cu): one thread-block cluster per
BatchNorm group runs the whole layer stack of cgcnn._inference_lcn (network/models_att.py:707-775) the way
base_model.predict feeds it (:79-132: batches of batch_size poses, last one zero padded, dropout 0).
Checked through the C ABI against the float64 oracle: every layer output (parity tap) against the oracle layer
applied to the GPU's own previous activation (1e-2 relative: the bf16 tolerance of BASELINE.json north_star),
the predictions end to end, and against the per-layer kernels of the training layout."""
import numpy as np
import pytest

from oracle import lcn_oracle as O
from tests.gpu_helpers import dev, make_pair, rel_err, synth_xy

pytestmark = pytest.mark.gpu


def _layer_refs(cfg, p, a_prev, l, bn_group, res):
    mask = O.mask_values(cfg, p)
    wn, bn_, bnn = O.weight_names(cfg), O.bias_names(cfg), O.bn_names(cfg)
    wm, _, _ = O.effective_weight(cfg, p[wn[l]], mask)
    z = a_prev @ wm + p[bn_[l]]
    a = np.empty_like(z)
    for g in range(z.shape[0] // bn_group):
        sl = slice(g * bn_group, (g + 1) * bn_group)
        a[sl] = O._bn_forward(z[sl], p[bnn[l] + "/gamma"], p[bnn[l] + "/beta"], cfg.F)[0]
    a = np.where(a > 0, a, 0.2 * a)
    if res is not None:
        a = a + res
    return a


@pytest.mark.parametrize("L,knn,mask_type,n,bn_group", [
    (3, 3, "locally_connected", 512, 256),     # two full groups of two row tiles
    (1, 3, "locally_connected", 1000, 200),    # reference default batch_size 200: padded tiles + zero-padded last batch
    (2, 2, "locally_connected", 300, 128),     # one row tile per group
    (2, 1, "locally_connected", 77, 100),      # single partial group
    (1, 2, "exponential", 600, 256),           # dense constant mask: every block present
])
def test_fused_inference_per_layer_and_end_to_end(L, knn, mask_type, n, bn_group):
    eng, cfg, p = make_pair(L=L, knn=knn, mask_type=mask_type, path="bf16")
    x, _ = synth_xy(n)
    out, taps = eng.forward_taps(dev(x), bn_group)
    out, taps = out.cpu().numpy(), taps.cpu().numpy().astype(np.float64)
    ng = (n + bn_group - 1) // bn_group
    xpad = np.zeros((ng * bn_group, x.shape[1]))
    xpad[:n] = x
    a_prev = xpad
    for l in range(1 + 2 * L):
        res = taps[l - 2] if (l >= 2 and l % 2 == 0) else None
        ref = _layer_refs(cfg, p, a_prev, l, bn_group, res)
        e = rel_err(taps[l], ref)
        assert e < 1e-2, f"A[{l}] rel err {e}"
        a_prev = taps[l]
    # head on the GPU's own last activation
    mask = O.mask_values(cfg, p)
    wm, _, _ = O.effective_weight(cfg, p[O.weight_names(cfg)[-1]], mask)
    y = a_prev @ wm + p[O.bias_names(cfg)[-1]]
    y = y.reshape(-1, 17, 3)
    y[:, :, :2] += xpad.reshape(-1, 17, 2)
    assert rel_err(out, y.reshape(-1, 51)[:n]) < 1e-2
    # end to end against the oracle's predict() and against the per-layer kernels
    ref = O.predict(cfg, p, x.astype(np.float64), bn_group)
    assert rel_err(out, ref) < 5e-2
    out_plain = eng.forward(dev(x), bn_group=bn_group, training=False).cpu().numpy()
    assert np.array_equal(out, out_plain)          # the tap does not change the arithmetic
    out_layers = eng.forward(dev(x), bn_group=bn_group, training=True).cpu().numpy()
    assert rel_err(out, out_layers) < 3e-2


def test_fused_inference_many_groups_is_group_local():
    """More groups than clusters: every group's result equals the result of running that group alone."""
    eng, cfg, p = make_pair(L=1, knn=3, path="bf16")
    n, bn = 256 * 70 + 19, 256
    x, _ = synth_xy(n)
    full = eng.forward(dev(x), bn_group=bn, training=False).cpu().numpy()
    assert np.isfinite(full).all()
    for g in (0, 33, 69, 70):
        part = eng.forward(dev(x[g * bn:(g + 1) * bn]), bn_group=bn, training=False).cpu().numpy()
        assert np.array_equal(full[g * bn:(g + 1) * bn], part), g
