"""GPU tests of the reference's method-level layer / mask API on the mirror (mask_weights,
batch_normalization_warp, kaiming, two_linear, _inference_lcn, loss, training) and of tools.procrustes with every option,
each against the float64 oracle (or the reference's own golden output) on seeded inputs, through the C ABI."""
import numpy as np
import pytest
import torch

from lcn_pose_b200.network import models_att
from lcn_pose_b200.tools import data as D
from lcn_pose_b200.tools import params_help, tools
from oracle import lcn_oracle as O
from tests.gpu_helpers import dev, elem_rel_err, make_pair, rel_err, synth_xy

pytestmark = pytest.mark.gpu


def _net(path="fp32", L=2, knn=2, reg=0, **kw):
    params = params_help.get_params(is_training=True)
    params.update(neighbour_matrix=params_help.get_neighbour_matrix_by_hand(params_help.filter_hub.neighbour_dict_set[0], knn=knn),
                  knn=knn, num_layers=L, init_type="same", regularization=reg, **kw)
    net = models_att.cgcnn(**params, path=path, seed=5)
    p = {k: v.astype(np.float64) for k, v in net.engine.get_params().items()}
    cfg = O.LcnConfig(F=params["F"], num_layers=L, neighbour_matrix=params["neighbour_matrix"],
                      regularization=reg or 0.0)
    return net, cfg, p


def test_mask_weights_takes_a_weight_tensor_like_the_reference():
    net, cfg, p = _net()
    w = p["linear_model/two_linear_0/w2_0"]
    got = net.mask_weights(w.astype(np.float32))                       # models_att.py:576-586: no clip inside
    ref = O.mask_weights(w, O.mask_values(cfg, p))
    assert got.shape == ref.shape and rel_err(got, ref) < 2e-6
    sup = np.kron(cfg.support(), np.ones((64, 64))) != 0
    assert np.all(got[~sup] == 0)
    w1 = p["linear_model/w1"]                                          # [34, 1088]: in_F != out_F
    assert rel_err(net.mask_weights(w1.astype(np.float32)), O.mask_weights(w1, O.mask_values(cfg, p))) < 2e-6
    t = net.mask_weights(dev(w.astype(np.float32)))                    # CUDA tensor in -> CUDA tensor out
    assert t.is_cuda and rel_err(t.cpu().numpy(), ref) < 2e-6
    assert rel_err(net.mask, O.mask_values(cfg, p)) < 2e-6


def test_batch_normalization_warp_is_keras_bn_with_batch_statistics():
    net, cfg, p = _net()
    rng = np.random.default_rng(0)
    y = (rng.normal(0, 2.0, (200, 17 * 64)) + rng.normal(0, 1, (1, 17 * 64))).astype(np.float32)
    name = "batch_normalization10"                                     # models_att.py:668 (first BN of block 0)
    g = np.float64(p["linear_model/two_linear_0/batch_normalization10/gamma"])
    b = np.float64(p["linear_model/two_linear_0/batch_normalization10/beta"])
    got = net.batch_normalization_warp(y, True, name)
    ref, _, _, _ = O._bn_forward(y.astype(np.float64), g, b, 64)
    assert rel_err(got, ref) < 1e-5
    with pytest.raises(KeyError):
        net.batch_normalization_warp(y, True, "no_such_layer")


def test_kaiming_is_truncated_normal_times_sqrt_2_over_fan_in():
    net, _, _ = _net()
    w = net.kaiming([1088, 1088], np.float32)
    assert w.shape == (1088, 1088) and w.dtype == np.float32
    s = np.sqrt(2.0 / 1088)
    assert np.abs(w).max() <= 2 * s * (1 + 1e-6)                        # truncated at 2 sigma
    assert abs(w.std() / s - 0.8796) < 0.01                             # std of a 2-sigma truncated normal
    assert net.kaiming([51], np.float32).shape == (51,)                 # biases use the same initializer (9-Q13)


@pytest.mark.parametrize("path,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_two_linear_block_on_an_injected_activation(path, tol):
    net, cfg, p = _net(path=path)
    rng = np.random.default_rng(1)
    n, idx = 200, 1
    xin = rng.normal(0, 1.0, (n, 17 * 64)).astype(np.float32)
    got = net.two_linear(xin, 0.0, idx)                                # models_att.py:630-705
    mask = O.mask_values(cfg, p)
    y = xin.astype(np.float64)
    for which in ("2", "3"):
        w = p[f"linear_model/two_linear_{idx}/w{which}_{idx}"]
        bb = p[f"linear_model/two_linear_{idx}/b{which}_{idx}"]
        bn = f"linear_model/two_linear_{idx}/batch_normalization{int(which) - 1}{idx}"
        wm, _, _ = O.effective_weight(cfg, w, mask)
        y, _, _, _ = O._bn_forward(y @ wm + bb, p[bn + "/gamma"], p[bn + "/beta"], 64)
        y = np.where(y > 0, y, 0.2 * y)
    ref = xin.astype(np.float64) + y
    assert got.shape == ref.shape and rel_err(got, ref) < tol


def test_inference_loss_training_sequence_is_one_reference_train_step():
    """_inference_lcn -> loss -> training in the order build_graph wires them (models_att.py:309-311) equals one
    oracle train step: loss value, the zero-debiased EMA, the learning rate op_train returns and the updated parameters."""
    net, cfg, p = _net(L=1, knn=3, reg=5e-4)
    x, y = synth_xy(256)
    st = O.AdamState()
    ema = models_att.DebiasedEma(0.9)
    for step in range(3):
        before = {k: v.copy() for k, v in p.items()}
        logits = net._inference_lcn(x, 0.0)
        ref_logits = O.forward(cfg, p, x.astype(np.float64))[0]
        assert rel_err(logits, ref_logits) < 5e-4
        loss, loss_avg = net.loss(logits, y)
        lr = net.training(loss, net.learning_rate, net.decay_type, net.decay_params)
        ref_loss, ref_lr, _ = O.train_step(cfg, p, st, x.astype(np.float64), y.astype(np.float64))
        assert abs(loss - ref_loss) < 2e-4 * ref_loss                   # includes regularization * sum l2_loss (:362-365)
        assert abs(loss_avg - ema.update(ref_loss)) < 2e-4 * ref_loss
        assert abs(lr - ref_lr) < 1e-12
        got = net.engine.get_params()
        for k in p:                                   # the step each element took, against the oracle's (|step| ~ lr)
            d_ref, d_got = p[k] - before[k], got[k].astype(np.float64) - before[k]
            bad = np.abs(d_got - d_ref) > 5e-3 * np.abs(d_ref).max() + 1e-7 * np.abs(before[k]).max()
            assert bad.mean() < 1e-4, (k, bad.mean())
        net.engine.set_params({k: v.astype(np.float32) for k, v in p.items()})      # keep both on the oracle's trajectory
        for k, (o, r, c) in net.engine.tensors.items():
            net.engine.adam_m[o:o + r * c].copy_(torch.as_tensor(st.m[k].astype(np.float32).reshape(-1)))
            net.engine.adam_v[o:o + r * c].copy_(torch.as_tensor(st.v[k].astype(np.float32).reshape(-1)))
    assert net.engine.step == 3                                          # global_step, incremented by apply_gradients
    with pytest.raises(AssertionError):
        net.training(0.0, 1e-3, "step", net.decay_params)               # models_att.py:400-401


@pytest.mark.parametrize("tag,kw", [("best", {}), ("noscale", {"scaling": False}), ("norefl", {"reflection": False}),
                                    ("refl", {"reflection": True}),
                                    ("noscale_norefl", {"scaling": False, "reflection": False})])
def test_procrustes_every_option_against_the_reference_output(golden, tag, kw):
    A, B = golden["proc_A"], golden["proc_B"]
    d, Z, tf = tools.procrustes_batch(A, B, **kw)
    assert np.abs(Z - golden[f"proc_{tag}_Z"]).max() < 0.05            # mm, float32 kernel vs the reference's float64
    assert np.abs(tf["rotation"] - golden[f"proc_{tag}_R"]).max() < 2e-4
    assert np.abs(tf["scale"] - golden[f"proc_{tag}_scale"]).max() < 2e-4
    assert np.abs(tf["translation"] - golden[f"proc_{tag}_t"]).max() < 2.0   # |B_bar| ~ 5e3 mm times the rotation error
    assert np.abs(d - golden[f"proc_{tag}_d"]).max() < 2e-5
    d1, Z1, tf1 = tools.procrustes(A[3], B[3], **kw)                    # the reference's per-pose signature
    assert np.abs(Z1 - golden[f"proc_{tag}_Z"][3]).max() < 0.05 and set(tf1) == {"rotation", "scale", "translation"}


def test_datareader_matches_the_reference_datareader(golden):
    cams = lambda split: [{"joint_3d_image": j, "camera_param": {"name": str(c)}, "cameraid": 0, "videoid": i, "subject": 1,
                           "action": int(golden["dr_denorm_action"][i]) if split == "test" else 2}
                          for i, (j, c) in enumerate(zip(golden[f"dr_{split}_j3d"], golden[f"dr_{split}_cam"]))]
    dr = D.DataReader()
    x_tr, x_te = dr.read_2d(cams("train"), cams("test"))
    y_tr, y_te = dr.read_3d()
    assert x_tr.dtype == np.float64 and x_tr.shape == golden["dr_x_train"].shape
    for got, key in ((x_tr, "dr_x_train"), (x_te, "dr_x_test"), (y_tr, "dr_y_train"), (y_te, "dr_y_test")):
        assert np.abs(got - golden[key]).max() < 2e-6, key             # float32 on the device, values in [-1, 1]
    res = dr.denormalize(golden["dr_y_test"])
    assert [r["action"] for r in res] == golden["dr_denorm_action"].tolist()
    assert set(res[0]) == {"cameraid", "videoid", "subject", "action", "result"}
    den = np.array([r["result"] for r in res])
    assert np.abs(den - golden["dr_denorm"]).max() < 1e-3              # pixels / mm at ~1e3: float32 resolution


def test_gather_rows_is_fancy_indexing():
    eng, _, _ = make_pair(L=1)
    rng = np.random.default_rng(2)
    xs, ys = rng.normal(size=(1000, 34)).astype(np.float32), rng.normal(size=(1000, 51)).astype(np.float32)
    idx = rng.permutation(1000)[:257].astype(np.int64)
    bx, by = torch.empty((257, 34), device="cuda"), torch.empty((257, 51), device="cuda")
    eng.gather_rows(dev(xs), bx, dev(idx), dev(ys), by)
    assert np.array_equal(bx.cpu().numpy(), xs[idx]) and np.array_equal(by.cpu().numpy(), ys[idx])
    eng.gather_rows(dev(xs), bx, dev(idx))                              # one row set
    assert np.array_equal(bx.cpu().numpy(), xs[idx])


def test_loss_ema_advances_every_step_like_op_loss_average():
    """The reference runs op_loss_average in every sess.run of fit (models_att.py:210-212): the 0.9 EMA must advance per
    step, not per epoch.  Device EMA inside the graph-replayed step against the host restatement fed the same losses."""
    eng, cfg, p = make_pair(L=1, knn=2, path="bf16")
    x, y = synth_xy(256)
    xd, yd = dev(x), dev(y)
    ema = models_att.DebiasedEma(0.9)
    want = 0.0
    for _ in range(7):
        loss, _ = eng.train_step_graph(xd, yd, dropout=0.25, track_ema=True)
        want = ema.update(float(loss.item()))
    assert abs(eng.loss_average() - want) < 1e-5 * want
