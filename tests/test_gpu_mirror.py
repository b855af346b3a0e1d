"""GPU test of the drop-in mirror of network/models_att.py (class cgcnn): the call sequence of the reference's train.py
(:107-113: get_params -> update_parameters -> cgcnn(**params) -> fit) and inference.py (:99-106: predict), on a small
synthetic set.  Checks the public contract: fit returns (losses, t_step), writes the reference's files, the loss goes
down, predict returns float64 [N, 51] with the zero-padded last batch, get_var serves the TF variable names, a
checkpoint restores the same predictions, and predict equals the oracle on the trained parameters."""
import json
import os
import shutil

import numpy as np
import pytest

from lcn_pose_b200.network import models_att
from lcn_pose_b200.tools import params_help
from oracle import lcn_oracle as O
from tests.gpu_helpers import rel_err, synth_xy

pytestmark = pytest.mark.gpu


class _Args:            # the argparse namespace of train.py:15-34 (defaults shortened for the test)
    test_indices = "pytest_mirror"; knn = 2; layers = 1; dropout = 0.0; channels = 64; checkpoints = "final"
    mask_type = "locally_connected"; init_type = "same"; epochs = 3; batch_size = 128
    learning_rate = 1e-3; regularization = None


def test_fit_predict_roundtrip_like_train_py(tmp_path, monkeypatch):
    monkeypatch.setattr(models_att, "ROOT_PATH", str(tmp_path))
    params = params_help.get_params(is_training=True)
    params_help.update_parameters(_Args, params)
    params["eval_frequency"] = 1
    net = models_att.cgcnn(**params, path="fp32", seed=3)
    x, y = synth_xy(128 * 6 + 37)
    xv, yv = synth_xy(300, seed=99)
    out_dir = str(tmp_path / "out")
    losses, t_step = net.fit(x, y, xv, yv, out_dir)
    assert len(losses) == 3 and all(np.isfinite(losses)) and t_step > 0
    assert losses[-1] < losses[0]                                  # three epochs of Adam reduce the validation loss
    assert json.load(open(os.path.join(out_dir, "training_error.json")))
    ck = os.path.join(str(tmp_path), "experiment", params["dir_name"], "checkpoints", "final")
    assert any(f.startswith("model-") for f in os.listdir(ck))
    pred = net.predict(xv)
    assert pred.dtype == np.float64 and pred.shape == (300, 51)
    # oracle on the trained parameters, batches of 128 with the zero-padded last one (models_att.py:92-97)
    p = {k: v.astype(np.float64) for k, v in net.engine.get_params().items()}
    cfg = O.LcnConfig(F=64, num_layers=1, neighbour_matrix=params["neighbour_matrix"])
    ref = O.predict(cfg, p, xv.astype(np.float64), 128)
    assert rel_err(pred, ref) < 5e-4
    # TF variable names and checkpoint restore
    assert net.get_var("linear_model/w1").shape == (34, 17 * 64)
    net2 = models_att.cgcnn(**params, path="fp32", seed=4)
    pred2 = net2.predict(xv)                                       # restores experiment/<dir>/checkpoints/final
    assert rel_err(pred2, pred) < 1e-6          # same checkpoint, same kernels (last bit: accumulation in arrival order)
    string, loss = net2.evaluate(xv, yv)
    assert "loss" in string and abs(loss - losses[-1]) < 1e-6 * max(1.0, abs(losses[-1]))


def test_invalid_init_type_raises_like_reference():
    params = params_help.get_params(is_training=True)
    params["init_type"] = "ones"
    with pytest.raises(ValueError, match="Unknown init_type"):
        models_att.cgcnn(**params)


def test_fit_with_a_validation_set_larger_than_a_predict_chunk(tmp_path, monkeypatch):
    """ADVICE (high): evaluate() inside fit() runs predict chunks of 32768 rows, whose workspace layout is larger than the
    train step's; the captured train-step graph must not end up replaying on a reallocated workspace.  fit() reserves
    the maximum of both layouts up front -- checked here: the workspace pointer never changes, the graph is captured once,
    training keeps converging after the first validation, and the validation loss equals the oracle's on the trained
    parameters."""
    monkeypatch.setattr(models_att, "ROOT_PATH", str(tmp_path))
    params = params_help.get_params(is_training=True)
    params_help.update_parameters(_Args, params)
    params["num_epochs"] = 4
    net = models_att.cgcnn(**params, path="bf16", seed=3)
    x, y = synth_xy(128 * 8)
    xv, yv = synth_xy(40000, seed=99)                               # > 32768: two predict chunks
    ptrs = []
    orig = net.engine.train_step_graph

    def spy(*a, **k):
        r = orig(*a, **k)
        ptrs.append((net.engine.ws.data_ptr(), len(net.engine._graphs)))
        return r
    monkeypatch.setattr(net.engine, "train_step_graph", spy)
    losses, _ = net.fit(x, y, xv, yv)
    assert len(losses) == 4 and all(np.isfinite(losses))
    assert len(set(p for p, _ in ptrs)) == 1, "the workspace was reallocated under the captured graph"
    assert all(n_graphs == 1 for _, n_graphs in ptrs)
    assert losses[-1] < losses[0]
    p = {k: v.astype(np.float64) for k, v in net.engine.get_params().items()}
    cfg = O.LcnConfig(F=64, num_layers=1, neighbour_matrix=params["neighbour_matrix"])
    _, ref_loss = O.predict(cfg, p, xv[:1280].astype(np.float64), 128, yv[:1280].astype(np.float64))
    _, got_loss = net.predict(xv[:1280], yv[:1280], sess=True)
    assert abs(got_loss - ref_loss) < 5e-2 * ref_loss              # bf16 path


def test_resume_keeps_optimizer_state_and_evaluate_restores_the_checkpoint(tmp_path, monkeypatch):
    """fit(starting_checkpoint=...) continues with the saved Adam slots and global_step (ADVICE: medium), and
    evaluate(data, labels) on a freshly built model restores the latest checkpoint like the reference's
    _get_session(None) instead of scoring the random initial parameters (ADVICE: low)."""
    monkeypatch.setattr(models_att, "ROOT_PATH", str(tmp_path))
    params = params_help.get_params(is_training=True)
    params_help.update_parameters(_Args, params)
    x, y = synth_xy(128 * 6)
    xv, yv = synth_xy(300, seed=99)
    net = models_att.cgcnn(**params, path="fp32", seed=3)
    losses, _ = net.fit(x, y, xv, yv)                                # 3 epochs x 6 steps
    assert net.engine.step == 18
    ck = os.path.join(str(tmp_path), "experiment", params["dir_name"], "checkpoints", "final")
    assert sorted(os.listdir(ck)) == ["checkpoint", "model-18.data-00000-of-00001", "model-18.index"]
    fresh = models_att.cgcnn(**params, path="fp32", seed=11)
    string, loss = fresh.evaluate(xv, yv)                            # sess=None -> restores model-18
    assert abs(loss - losses[-1]) < 1e-6 * max(1.0, abs(losses[-1])) and "time:" in string
    params2 = dict(params, num_epochs=5)
    net2 = models_att.cgcnn(**params2, path="fp32", seed=12)
    keep = str(tmp_path / "resume_from")
    shutil.copytree(ck, keep)
    losses2, _ = net2.fit(x, y, xv, yv, starting_checkpoint=keep)
    assert net2.engine.step > 18 and float(net2.engine.adam_v.abs().sum()) > 0
    assert losses2[0] < losses[0]                                    # it continued from the trained state, not from scratch
