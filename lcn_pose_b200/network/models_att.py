"""Drop-in mirror of the reference's network/models_att.py for the LCN hot path.

Same public surface as the reference: class cgcnn with the constructor kwargs of models_att.py:478-506, fit / predict /
evaluate with the signatures of :79-286, the layer / mask methods of SURVEY 8(b) (_initialize_mask, mask_weights,
batch_normalization_warp, kaiming, two_linear, _inference_lcn, inference, prediction, loss, training, build_graph,
get_var) and the module function get_exponential_matrix.  train.py / inference.py run against it unchanged (put
`shims/` on PYTHONPATH).  Nothing here builds a TensorFlow graph: the reference's graph-building methods
take symbolic tensors and return symbolic tensors; their mirrors take arrays and run the same op eagerly as kernels of
liblcn_b200.so (through lcn_pose_b200.engine.LcnEngine).  There is no CPU fallback.

Differences from the reference that are deliberate:
 * Q10 -- resuming restores the checkpoint (variables, Adam slots, global_step) and keeps it; the reference
   re-initialises every variable after restoring;
 * checkpoints are TensorFlow `Saver` V2 files (TensorBundle: model-<step>.index / .data-00000-of-00001 plus the
   `checkpoint` state file) written and read by lcn_pose_b200/tools/tf_checkpoint.py with the reference's variable names,
   so weights trained by the reference load here and vice versa; no .meta graph is written;
 * TensorBoard summaries are not written.
"""
import collections
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
from ..tools import tf_checkpoint

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
    """Host restatement of tf.train.ExponentialMovingAverage(0.9).apply on a tensor (models_att.py:370-379):
    zero-initialised shadow with zero-debias [TF-sem].  fit() keeps this state on the device
    (LcnEngine.update_loss_ema, every step); this class is what the CPU tests compare it with."""

    def __init__(self, decay=0.9):
        self.decay, self.biased, self.t = decay, 0.0, 0

    def update(self, x):
        self.t += 1
        self.biased = self.decay * self.biased + (1 - self.decay) * x
        return self.biased / (1 - self.decay ** self.t)


def _as_dev(a, device):
    if torch.is_tensor(a):
        return a.to(device=device, dtype=torch.float32).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).to(device)


class base_model(object):
    def __init__(self):
        self.regularizers = []
        self.checkpoints = "final"
        self.writer = None

    # ---- models_att.py:79-132 ---------------------------------------------------------------------
    def predict(self, data, labels=None, sess=None):
        """Batches of batch_size poses, the last one zero padded (the zero rows take part in the BatchNorm
        statistics); dropout 0.  Returns float64 [N, 51] (and loss * batch_size / N when labels are given, the loss of
        a batch being op_loss = mse + regularization * sum l2_loss, models_att.py:107-116,130,362-365)."""
        if sess is None and not self._restored:
            self._restore_latest()                                            # _get_session(None), :445-463
        data = np.asarray(data.toarray() if hasattr(data, "toarray") else data)
        size = data.shape[0]
        if labels is None:
            return self.engine.predict(data, self.batch_size)
        # with labels the reference also feeds zero-padded labels for the last batch and sums the batch
        # losses: run the padded rows explicitly so their predictions exist
        pad = (-size) % self.batch_size
        data_p = np.concatenate([data, np.zeros((pad,) + data.shape[1:], data.dtype)]) if pad else data
        labels_p = np.asarray(labels, dtype=np.float64)
        labels_p = np.concatenate([labels_p, np.zeros((pad,) + labels_p.shape[1:])]) if pad else labels_p
        preds_p = self.engine.predict(data_p, self.batch_size)
        per_batch = ((preds_p - labels_p) ** 2).reshape(-1, self.batch_size * preds_p.shape[1]).mean(axis=1)
        loss = float(per_batch.sum())
        if self.regularization:                                               # `!= 0 and is not None`, :362
            loss += len(per_batch) * float(self.regularization) * self.engine.l2_regularizer()
        return preds_p[:size], loss * self.batch_size / size

    # ---- models_att.py:134-146 --------------------------------------------------------------------
    def evaluate(self, data, labels, sess=None):
        t_process, t_wall = time.process_time(), time.time()
        predictions, loss = self.predict(data, labels, sess)   # sess=None restores the latest checkpoint, like the reference
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
            for file in sorted(os.listdir(starting_checkpoint)):
                match = re.search(r"model-(\d+)\.index$", file)
                if match:
                    starting_step = int(match.group(1))
                    self._load(os.path.join(starting_checkpoint, file[: -len(".index")]))
                    print(f"Resuming from step {starting_step}")
        n = train_data.shape[0]
        num_steps, eval_frequency = schedule(self.num_epochs, n, self.batch_size)
        print(f"Total steps to be done to complete all the epochs: {num_steps}")
        # dataset resident on the device; the per-step gather (train_data[idx], :200) is one kernel over both sets
        xd = _as_dev(train_data, eng.device)
        yd = _as_dev(train_labels, eng.device)
        sampler = PermutationSampler(n, self.batch_size)
        # fixed batch buffers: the step is a CUDA-graph replay (LcnEngine.train_step_graph) fed by the device gather.
        # The workspace is sized once for the train step AND the validation predict chunks, so that evaluate() inside
        # the loop can never reallocate it under the captured graph.
        bx = torch.empty((self.batch_size, xd.shape[1]), dtype=torch.float32, device=eng.device)
        by = torch.empty((self.batch_size, yd.shape[1]), dtype=torch.float32, device=eng.device)
        val_chunk = max(1, 32768 // self.batch_size) * self.batch_size
        n_val = int(np.shape(val_data)[0])
        eng.reserve_ws((self.batch_size, self.batch_size, True),
                       (min(val_chunk, max(n_val + (-n_val) % self.batch_size, 1)), self.batch_size, False))
        idx_pin = [torch.empty(self.batch_size, dtype=torch.int64).pin_memory() for _ in range(2)]
        idx_dev = [torch.empty(self.batch_size, dtype=torch.int64, device=eng.device) for _ in range(2)]
        idx_ev = [None, None]
        losses, training_error, validation_error = [], [], []
        min_loss = 10000
        self._restored = True
        for step in range(starting_step, num_steps + 1):
            b = step & 1
            if idx_ev[b] is not None:
                idx_ev[b].synchronize()                    # the copy that last read this pinned index buffer is done
            idx_pin[b].copy_(torch.from_numpy(sampler.next()))
            idx_dev[b].copy_(idx_pin[b], non_blocking=True)
            idx_ev[b] = torch.cuda.Event()
            idx_ev[b].record(torch.cuda.current_stream(eng.device))
            eng.gather_rows(xd, bx, idx_dev[b], yd, by)
            loss_dev, learning_rate = eng.train_step_graph(bx, by, dropout=self.dropout, track_ema=True)
            if eval_frequency > 0 and step % eval_frequency == 0:
                loss_average = eng.loss_average()          # EMA advanced on the device EVERY step (:210-212,370-379)
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

    # ---- models_att.py:335-350: thin graph-construction wrappers, eager here ---------------------------
    def initialize_mask(self):
        self._initialize_mask()

    def inference(self, data, dropout):
        return self._inference_lcn(data, data_dropout=dropout)

    def prediction(self, logits):
        return logits

    def loss(self, logits, labels):
        """models_att.py:352-380: returns (loss, loss_average) -- mse (+ regularization * sum l2_loss) and its 0.9 EMA
        with zero-debias, which this call advances by one step exactly like evaluating op_loss_average does."""
        eng = self.engine
        lg, lb = _as_dev(logits, eng.device), _as_dev(labels, eng.device)
        self._last_labels = lb
        eng.loss_dev.copy_(eng.mse_loss(lg, lb, with_reg=False))
        eng.update_loss_ema()
        mse = float(eng.loss_dev.item())
        total = mse + (float(self.regularization) * eng.l2_regularizer() if self.regularization else 0.0)
        return total, eng.loss_average()

    def training(self, loss, learning_rate, decay_type, decay_params):
        """models_att.py:382-421: one optimizer step for the batch of the last _inference_lcn / loss pair --
        compute_gradients + TF1 Adam apply_gradients + global_step += 1.  Returns the learning rate used, which is what
        op_train evaluates to (:418-420)."""
        if decay_type != "exp":
            assert 0, "not implemented lr decay types!"                         # :400-401
        eng = self.engine
        eng.learning_rate = learning_rate
        eng.decay_steps, eng.decay_rate = decay_params["decay_steps"], decay_params["decay_rate"]
        if getattr(self, "_last_x", None) is None or getattr(self, "_last_labels", None) is None:
            raise _lib.LcnError("training() needs a preceding _inference_lcn(x, dropout) and loss(logits, labels)")
        eng.backward(self._last_x, self._last_labels, self._last_dropout)
        return eng.adam()

    # ---- helpers ------------------------------------------------------------------------------------
    def get_var(self, name):
        """models_att.py:424-429: value of a variable of the latest checkpoint, by TF variable name."""
        if not self._restored:
            self._restore_latest()
        if name == "global_step":
            return np.asarray(self.engine.step, dtype=np.int32)
        return self.engine.get_params()[name]

    def _get_path(self, folder):
        return os.path.join(ROOT_PATH, "experiment", self.dir_name, folder)

    def _save(self, directory, step):
        """op_saver.save(sess, path, global_step=step) (models_att.py:256-260): Saver(max_to_keep=1) semantics."""
        tf_checkpoint.save_model(directory, step, self.engine.get_params(), self.engine.get_state(),
                                 self.engine.tensors)

    def _load(self, prefix):
        params, state = tf_checkpoint.load_model(prefix, self.engine.tensors, self.engine.n_params)
        self.engine.set_params(params)
        if state is not None:
            self.engine.set_state(state)
        self._restored = True

    def _restore_latest(self):
        d = os.path.join(self._get_path("checkpoints"), self.checkpoints)
        prefix = tf_checkpoint.latest_checkpoint(d)                          # tf.train.latest_checkpoint, :456-458
        if prefix is None:
            raise FileNotFoundError("no checkpoint in %s" % d)
        print("restore from %s" % prefix)
        self._load(prefix)


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
        self.activation = "leaky_relu(alpha=0.2)"                                # tf.nn.leaky_relu, :526
        self._restored = False
        self._last_x = self._last_labels = None
        self._last_dropout = 0.0
        self.build_graph(in_joints, self.in_F, path=path, device=device, seed=seed)

    def build_graph(self, M_0, in_F, path="bf16", device="cuda:0", seed=None):
        """models_att.py:288-333: creates the mask and every variable (here: the engine and its flat buffers)."""
        self.initialize_mask()
        self.engine = LcnEngine(F=self.F, in_F=in_F, num_layers=self.num_layers, mask_type=self.mask_type,
                                neighbour_matrix=self.neighbour_matrix, residual=self.residual,
                                batch_norm=self.batch_norm, max_norm=self.max_norm, path=path, device=device,
                                learning_rate=self.learning_rate, decay_steps=self.decay_params["decay_steps"],
                                decay_rate=self.decay_params["decay_rate"], regularization=self.regularization)
        self.engine.init_params(seed=np.random.randint(1 << 31) if seed is None else seed)   # op_init

    # ---- mask ---------------------------------------------------------------------------------------
    def _initialize_mask(self):
        """models_att.py:534-574: only init_type 'same' is valid for the trainable mask."""
        if "locally_connected" in self.mask_type:
            assert self.neighbour_matrix is not None
            L = self.neighbour_matrix.T
            assert L.shape == (self.in_joints, self.in_joints)
            if self.init_type != "same":
                raise ValueError("Unknown init_type: {}".format(self.init_type))

    @property
    def mask(self):
        """self.mask of the reference (:571,574): softmax(var, axis=0) * support or the exponential constant,
        float32 [in joint, out joint], evaluated on the device from the current variable."""
        self.engine.prepare()
        return self.engine.read_tensor(3, 0, 128, 128).cpu().numpy()

    mask_values = mask.fget

    def mask_weights(self, weights):
        """models_att.py:576-586: reshape(weights, [17, Fi, 17, Fo]) * mask[17, 1, 17, 1] for a [17*Fi, 17*Fo] array
        (NumPy in -> NumPy out, CUDA tensor in -> CUDA tensor out), with the current mask values."""
        eng = self.engine
        is_np = not torch.is_tensor(weights)
        w = _as_dev(weights, eng.device)
        assert w.shape[0] % self.in_joints == 0 and w.shape[1] % self.in_joints == 0
        out = eng.mask_weights(w)
        return out.cpu().numpy() if is_np else out

    # ---- layers ---------------------------------------------------------------------------------------
    def _bn_tensor_prefix(self, name):
        for k in self.engine.tensors:
            if k.endswith("/" + name + "/gamma"):
                return k[: -len("/gamma")]
        raise KeyError("no BatchNormalization layer named %r" % name)

    def batch_normalization_warp(self, y, training, name):
        """models_att.py:588-612: Keras BatchNormalization(axis=-1, name=name) on reshape(y, [-1, 17, F]).  `training`
        is accepted for signature parity; like in the reference it cannot switch to moving statistics -- every entry
        point passes is_training=True and the moving averages are never updated nor read."""
        is_np = not torch.is_tensor(y)
        out = self.engine.batch_norm(_as_dev(y, self.engine.device), self._bn_tensor_prefix(name))
        return out.cpu().numpy() if is_np else out

    def kaiming(self, shape, dtype=np.float32, partition_info=None):
        """models_att.py:614-628: truncated_normal(shape) * sqrt(2 / shape[0]) (used for weights AND biases, 9-Q13).
        An initializer, not on the hot path: drawn on the host."""
        out = np.random.standard_normal(tuple(shape))
        bad = np.abs(out) > 2
        while bad.any():
            out[bad] = np.random.standard_normal(int(bad.sum()))
            bad = np.abs(out) > 2
        return (out * math.sqrt(2.0 / float(shape[0]))).astype(dtype)

    def two_linear(self, xin, data_dropout, idx):
        """models_att.py:630-705: residual block idx on an activation xin [B, 17*F] with the block's own variables
        (w2_idx, b2_idx, w3_idx, b3_idx and its two BatchNorm layers): two LCN layers, each followed by BN, LeakyReLU
        and dropout, then xin + y.  Runs lcn_model_forward_layers on the injected activation."""
        eng = self.engine
        assert 0 <= idx < self.num_layers
        is_np = not torch.is_tensor(xin)
        a = _as_dev(xin, eng.device)
        n = a.shape[0]
        eng.write_activation(2 * idx, a)
        eng.forward_layers(2 * idx + 1, 2 * idx + 3, n, dropout=float(data_dropout))
        out = eng.read_tensor(1, 2 * idx + 2, n, n)
        return out.cpu().numpy() if is_np else out

    def _inference_lcn(self, x, data_dropout=0.0):
        """models_att.py:707-775 on one batch: x [B, 17*in_F] -> logits [B, 51].  The batch is one BatchNorm group and
        the activations stay in the workspace, so loss() / training() can follow like in the reference's graph."""
        eng = self.engine
        is_np = not torch.is_tensor(x)
        xd = _as_dev(x, eng.device)
        self._last_x, self._last_dropout = xd, float(data_dropout)
        out = eng.forward(xd, bn_group=xd.shape[0], training=True, dropout=float(data_dropout))
        return out.cpu().numpy() if is_np else out
