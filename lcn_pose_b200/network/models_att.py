"""Drop-in mirror of the reference's network/models_att.py for the LCN hot path.

Same public surface as the reference (class cgcnn with the constructor kwargs of models_att.py:478-506,
fit / predict / evaluate with the signatures of :79-286, get_exponential_matrix), so train.py / inference.py
keep working against it, but nothing here builds a TensorFlow graph: every arithmetic step is a kernel of
liblcn_b200.so driven through lcn_pose_b200.engine.LcnEngine.  There is no CPU fallback.

Differences from the reference that are deliberate:
 * Q10 -- resuming restores the checkpoint and keeps it (the reference re-initialises after restoring);
 * checkpoints are .npz files named experiment/<dir>/checkpoints/{final,best}/model-<step>.npz whose keys are
   the reference's TF variable names;
 * TensorBoard summaries are not written.
"""
import collections
import glob
import json
import math
import os
import re
import shutil
import time

import numpy as np
import torch

from .. import _lib
from ..engine import LcnEngine

ROOT_PATH = os.path.join(os.path.dirname(os.path.realpath(__file__)), "..", "..")


def get_exponential_matrix():
    """network/models_att.py:14-69 -- float32 [17,17] of 1/2**hop_distance, bit exact (C ABI)."""
    return _lib.exponential_matrix()


class PermutationSampler:
    """The epoch-free sampler of base_model.fit (models_att.py:194-198): a deque refilled with a fresh
    np.random.permutation(N) whenever fewer than batch_size indices are left."""

    def __init__(self, n, batch_size, rng=None):
        self.n, self.batch_size = n, batch_size
        self.rng = rng if rng is not None else np.random
        self.indices = collections.deque()

    def next(self):
        if len(self.indices) < self.batch_size:
            self.indices.extend(self.rng.permutation(self.n))
        return np.fromiter((self.indices.popleft() for _ in range(self.batch_size)), dtype=np.int64,
                           count=self.batch_size)


def schedule(num_epochs, n_train, batch_size):
    """num_steps / eval_frequency of base_model.fit (models_att.py:185-186)."""
    num_steps = int(num_epochs * n_train / batch_size)
    eval_frequency = num_steps // num_epochs
    return num_steps, eval_frequency


class DebiasedEma:
    """tf.train.ExponentialMovingAverage(0.9).apply on a tensor (models_att.py:370-379): zero-initialised
    shadow with zero-debias [TF-sem]."""

    def __init__(self, decay=0.9):
        self.decay, self.biased, self.t = decay, 0.0, 0

    def update(self, x):
        self.t += 1
        self.biased = self.decay * self.biased + (1 - self.decay) * x
        return self.biased / (1 - self.decay ** self.t)


class base_model(object):
    def __init__(self):
        self.regularizers = []
        self.checkpoints = "final"
        self.writer = None

    # ---- models_att.py:79-132 ---------------------------------------------------------------------
    def predict(self, data, labels=None, sess=None):
        """Batches of batch_size poses, the last one zero padded (the zero rows take part in the BatchNorm
        statistics); dropout 0.  Returns float64 [N, 51] (and loss * batch_size / N when labels are given)."""
        if sess is None and not self._restored:
            self._restore_latest()
        data = np.asarray(data.toarray() if hasattr(data, "toarray") else data)
        size = data.shape[0]
        if labels is None:
            return self.engine.predict(data, self.batch_size)
        # with labels the reference also feeds zero-padded labels for the last batch and averages the batch
        # losses (models_att.py:107-116,130): run the padded rows explicitly so their predictions exist
        pad = (-size) % self.batch_size
        data_p = np.concatenate([data, np.zeros((pad,) + data.shape[1:], data.dtype)]) if pad else data
        labels_p = np.asarray(labels, dtype=np.float64)
        labels_p = np.concatenate([labels_p, np.zeros((pad,) + labels_p.shape[1:])]) if pad else labels_p
        preds_p = self.engine.predict(data_p, self.batch_size)
        per_batch = ((preds_p - labels_p) ** 2).reshape(-1, self.batch_size * preds_p.shape[1]).mean(axis=1)
        return preds_p[:size], float(per_batch.sum()) * self.batch_size / size

    # ---- models_att.py:134-146 --------------------------------------------------------------------
    def evaluate(self, data, labels, sess=None):
        t_process, t_wall = time.process_time(), time.time()
        predictions, loss = self.predict(data, labels, sess if sess is not None else True)
        string = "loss: {:.4e}".format(loss)
        if sess is None:
            string += "\ntime: {:.0f}s (wall {:.0f}s)".format(time.process_time() - t_process, time.time() - t_wall)
        return string, loss

    # ---- models_att.py:148-286 --------------------------------------------------------------------
    def fit(self, train_data, train_labels, val_data, val_labels, output_dir=None, starting_checkpoint=None):
        t_process, t_wall = time.process_time(), time.time()
        eng = self.engine
        starting_step = 1
        path = os.path.join(self._get_path("checkpoints"), "final")
        best_path = os.path.join(self._get_path("checkpoints"), "best")
        if starting_checkpoint is None:
            shutil.rmtree(self._get_path("checkpoints"), ignore_errors=True)
            os.makedirs(path, exist_ok=True)
            os.makedirs(best_path, exist_ok=True)
        else:
            for file in os.listdir(starting_checkpoint):
                match = re.search(r"model-(\d+)", file)
                if match:
                    starting_step = int(match.group(1))
                    self._load(os.path.join(starting_checkpoint, file))
                    print(f"Resuming from step {starting_step}")
        n = train_data.shape[0]
        num_steps, eval_frequency = schedule(self.num_epochs, n, self.batch_size)
        print(f"Total steps to be done to complete all the epochs: {num_steps}")
        # dataset resident on the device; per-step gather is a device index_select
        xd = torch.as_tensor(np.ascontiguousarray(train_data, dtype=np.float32)).to(eng.device)
        yd = torch.as_tensor(np.ascontiguousarray(train_labels, dtype=np.float32)).to(eng.device)
        sampler = PermutationSampler(n, self.batch_size)
        # fixed batch buffers: the step is a CUDA-graph replay (LcnEngine.train_step_graph) fed by a device gather
        bx = torch.empty((self.batch_size, xd.shape[1]), dtype=torch.float32, device=eng.device)
        by = torch.empty((self.batch_size, yd.shape[1]), dtype=torch.float32, device=eng.device)
        ema = DebiasedEma(0.9)
        losses, training_error, validation_error = [], [], []
        min_loss = 10000
        self._restored = True
        for step in range(starting_step, num_steps + 1):
            idx = torch.as_tensor(sampler.next()).to(eng.device, non_blocking=True)
            torch.index_select(xd, 0, idx, out=bx)
            torch.index_select(yd, 0, idx, out=by)
            loss_dev, learning_rate = eng.train_step_graph(bx, by, dropout=self.dropout)
            if eval_frequency > 0 and step % eval_frequency == 0:
                loss_average = ema.update(float(loss_dev.item()))
                epoch = step * self.batch_size / n
                print("step {} / {} (epoch {:.2f} / {}):".format(step, num_steps, epoch, self.num_epochs))
                print("  learning_rate = {:.2e}, loss_average = {:.4e}".format(learning_rate, loss_average))
                training_error.append([time.time(), step, loss_average])
                string, loss = self.evaluate(val_data, val_labels, sess=True)
                losses.append(loss)
                print("validation {}".format(string))
                print("time: {:.0f}s (wall {:.0f}s)".format(time.process_time() - t_process, time.time() - t_wall))
                validation_error.append([0, step, loss])
                self._save(path, step)
                if loss < min_loss:
                    min_loss = loss
                    self._save(best_path, step)
        print("validation loss: trough = {:.4f}, mean = {:.2f}".format(min_loss, np.mean(losses[-10:]) if losses else float("nan")))
        t_step = (time.time() - t_wall) / max(num_steps, 1)
        if output_dir is not None:
            os.makedirs(output_dir, exist_ok=True)
            with open(output_dir + "/training_error.json", "w") as f:
                f.write(json.dumps([training_error], indent=4))
            with open(output_dir + "/validation_error.json", "w") as f:
                f.write(json.dumps([validation_error], indent=4))
        return losses, t_step

    # ---- helpers ------------------------------------------------------------------------------------
    def get_var(self, name):
        """models_att.py:424-429: value of a variable of the latest checkpoint, by TF variable name."""
        if not self._restored:
            self._restore_latest()
        return self.engine.get_params()[name]

    def _get_path(self, folder):
        return os.path.join(ROOT_PATH, "experiment", self.dir_name, folder)

    def _save(self, directory, step):
        os.makedirs(directory, exist_ok=True)
        for old in glob.glob(os.path.join(directory, "model-*.npz")):   # Saver(max_to_keep=1)
            os.remove(old)
        p = self.engine.get_params()
        p["global_step"] = np.asarray(step)
        np.savez(os.path.join(directory, f"model-{step}.npz"), **p)

    def _load(self, file):
        ck = np.load(file)
        self.engine.set_params({k: ck[k] for k in ck.files if k != "global_step"})
        self._restored = True

    def _restore_latest(self):
        d = os.path.join(self._get_path("checkpoints"), self.checkpoints)
        files = sorted(glob.glob(os.path.join(d, "model-*.npz")), key=lambda f: int(re.search(r"model-(\d+)", f).group(1)))
        if not files:
            raise FileNotFoundError("no checkpoint in %s" % d)
        print("restore from %s" % files[-1])
        self._load(files[-1])


class cgcnn(base_model):
    """Locally connected network of the reference (models_att.py:475-775) on the B200 engine."""

    def __init__(self, F=64, mask_type="locally_connected", init_type="ones", neighbour_matrix=None, in_joints=17,
                 out_joints=17, in_F=2, num_layers=2, residual=True, batch_norm=True, max_norm=True, num_epochs=200,
                 learning_rate=0.001, decay_type="exp", decay_params=None, regularization=0.0, dropout=0,
                 batch_size=200, eval_frequency=200, dir_name="", checkpoints="final", is_training=True, knn=1,
                 path="bf16", device="cuda:0", seed=None):
        super().__init__()
        assert neighbour_matrix.shape[0] == neighbour_matrix.shape[1]
        assert neighbour_matrix.shape[0] == in_joints
        assert in_joints == 17 and out_joints == 17
        if decay_type != "exp":
            assert 0, "not implemented lr decay types!"                         # models_att.py:400-401
        self.F, self.mask_type, self.init_type = F, mask_type, init_type
        self.neighbour_matrix = neighbour_matrix
        self.in_joints, self.out_joints, self.num_layers = in_joints, out_joints, num_layers
        self.residual, self.batch_norm, self.max_norm = residual, batch_norm, max_norm
        self.num_epochs, self.learning_rate = num_epochs, learning_rate
        self.decay_type = decay_type
        self.decay_params = decay_params or {"decay_steps": 32000, "decay_rate": 0.96}
        self.regularization, self.dropout = regularization, dropout
        self.batch_size, self.eval_frequency = batch_size, eval_frequency
        self.dir_name, self.checkpoints = dir_name, checkpoints
        self.in_F, self.is_training, self.knn = in_F, is_training, knn
        self._restored = False
        self.build_graph(in_joints, self.in_F, path=path, device=device, seed=seed)

    def build_graph(self, M_0, in_F, path="bf16", device="cuda:0", seed=None):
        """models_att.py:288-333: creates the mask and every variable (here: the engine and its flat buffers)."""
        self._initialize_mask()
        self.engine = LcnEngine(F=self.F, in_F=in_F, num_layers=self.num_layers, mask_type=self.mask_type,
                                neighbour_matrix=self.neighbour_matrix, residual=self.residual,
                                batch_norm=self.batch_norm, max_norm=self.max_norm, path=path, device=device,
                                learning_rate=self.learning_rate, decay_steps=self.decay_params["decay_steps"],
                                decay_rate=self.decay_params["decay_rate"], regularization=self.regularization)
        self.engine.init_params(seed=np.random.randint(1 << 31) if seed is None else seed)   # op_init

    def _initialize_mask(self):
        """models_att.py:534-574: only init_type 'same' is valid for the trainable mask."""
        if "locally_connected" in self.mask_type:
            assert self.neighbour_matrix is not None
            if self.init_type != "same":
                raise ValueError("Unknown init_type: {}".format(self.init_type))

    initialize_mask = _initialize_mask

    def mask_weights(self, name):
        """models_att.py:576-586 (after clip_by_norm, :659): the dense effective weight of variable `name`,
        computed on the device."""
        names = [n for n in self.engine.tensors if n.rsplit("/", 1)[-1].startswith("w")]
        layer = names.index(name)
        self.engine.prepare()
        return self.engine.read_tensor(2, layer, 128, 128).cpu().numpy()

    def mask_values(self):
        """softmax(var, axis=0) * support, or the exponential constant (models_att.py:569-574)."""
        self.engine.prepare()
        return self.engine.read_tensor(3, 0, 128, 128).cpu().numpy()

    def inference(self, data, dropout=0.0):
        """cgcnn._inference_lcn (models_att.py:707-775) on one batch: data [B, 34] -> [B, 51]."""
        x = torch.as_tensor(np.ascontiguousarray(data, dtype=np.float32)).to(self.engine.device)
        return self.engine.forward(x, bn_group=x.shape[0], training=False, dropout=dropout).cpu().numpy()

    _inference_lcn = inference

    def prediction(self, logits):
        return logits
