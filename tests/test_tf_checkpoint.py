"""TensorFlow Saver-V2 (TensorBundle) reader / writer, lcn_pose_b200/tools/tf_checkpoint.py (the
reference saves / restores with tf.compat.v1.train.Saver, network/models_att.py:256-260,445-463).  CPU only."""
import os
import struct

import numpy as np
import pytest

from lcn_pose_b200.tools import tf_checkpoint as T


def test_crc32c_known_answers_and_native_equals_python():
    assert T.crc32c(b"123456789") == 0xE3069283                  # the standard CRC-32C check value
    assert T.crc32c(bytes(32)) == 0x8A9136AA                     # TensorFlow's crc32c_test.cc vectors
    assert T.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    assert T.crc32c(bytes(range(32))) == 0x46DD794E
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 8, 9, 63, 1000, 4099):
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert T.crc32c(b) == T.crc32c_py(b), n
    assert T.crc32c(b"56789", T.crc32c(b"1234")) == 0xE3069283   # incremental
    assert T.unmask_crc(T.mask_crc(0xDEADBEEF)) == 0xDEADBEEF


def _tensors():
    rng = np.random.default_rng(1)
    t = {"mask": rng.normal(size=(17, 17)).astype(np.float32), "global_step": np.asarray(1234, np.int32),
         "linear_model/w1": rng.normal(size=(34, 1088)).astype(np.float32), "steps64": np.arange(3, dtype=np.int64)}
    for i in range(70):                                          # > 4 restart intervals, shared key prefixes
        t["linear_model/two_linear_%d/w2_%d/Adam" % (i, i)] = rng.normal(size=(i + 1, 3)).astype(np.float32)
    return t


def test_bundle_round_trip_and_file_structure(tmp_path):
    t = _tensors()
    prefix = str(tmp_path / "model-1234")
    T.write_bundle(prefix, t)
    assert sorted(os.listdir(tmp_path)) == ["model-1234.data-00000-of-00001", "model-1234.index"]
    raw = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xDB4775248B80FB57          # LevelDB table magic
    assert os.path.getsize(prefix + ".data-00000-of-00001") == sum(v.nbytes for v in t.values())
    table = T.read_table(prefix + ".index")
    assert [k for k, _ in table] == sorted(k for k, _ in table) and table[0][0] == b""   # header first, keys sorted
    back = T.read_bundle(prefix)
    assert set(back) == set(t)
    for k in t:
        assert back[k].dtype == t[k].dtype and back[k].shape == t[k].shape and np.array_equal(back[k], t[k]), k
    assert T.list_bundle(prefix)["linear_model/w1"] == (T.DT_FLOAT, (34, 1088))
    assert T.list_bundle(prefix)["global_step"] == (T.DT_INT32, ())


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / "model-1")
    T.write_bundle(prefix, _tensors())
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[100] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    with pytest.raises(ValueError, match="checksum"):
        T.read_bundle(prefix)
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[20] ^= 0x01
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(ValueError):
        T.read_table(prefix + ".index")


def test_save_model_uses_reference_variable_names_and_latest_checkpoint(tmp_path):
    table = {"mask": (0, 17, 17), "linear_model/w1": (292, 34, 64), "linear_model/b1": (292 + 34 * 64, 1, 64),
             "linear_model/batch_normalization/gamma": (292 + 35 * 64, 1, 64),
             "linear_model/batch_normalization/beta": (292 + 36 * 64, 1, 64)}
    n = 292 + 37 * 64
    rng = np.random.default_rng(2)
    params = {k: rng.normal(size=(r, c) if r > 1 else (c,)).astype(np.float32) for k, (o, r, c) in table.items()}
    state = {"adam_m": rng.normal(size=n).astype(np.float32), "adam_v": rng.random(n).astype(np.float32),
             "global_step": np.asarray(77), "loss_ema": np.array([0.5, 77.0], np.float32)}
    d = str(tmp_path / "final")
    T.save_model(d, 70, params, state, table)
    p2 = T.save_model(d, 77, params, state, table)                       # max_to_keep = 1: step 70 is gone
    assert sorted(os.listdir(d)) == ["checkpoint", "model-77.data-00000-of-00001", "model-77.index"]
    assert T.latest_checkpoint(d) == p2
    names = T.list_bundle(p2)
    for want in ("mask", "linear_model/w1", "linear_model/w1/Adam", "linear_model/w1/Adam_1", "global_step",
                 "training/beta1_power", "linear_model/batch_normalization/moving_mean",
                 "linear_model/batch_normalization/moving_variance"):
        assert want in names, want
    back, st = T.load_model(p2, table, n)
    for k in params:
        assert np.array_equal(back[k], params[k])
    assert st["global_step"] == 77
    for k, (o, r, c) in table.items():
        assert np.array_equal(st["adam_m"][o:o + r * c], state["adam_m"][o:o + r * c])
        assert np.array_equal(st["adam_v"][o:o + r * c], state["adam_v"][o:o + r * c])
    # a weights-only bundle (e.g. a stripped reference checkpoint) restores the variables and no optimizer state
    T.write_bundle(str(tmp_path / "weights"), params)
    back2, st2 = T.load_model(str(tmp_path / "weights"), table, n)
    assert st2 is None and all(np.array_equal(back2[k], params[k]) for k in params)
    os.remove(os.path.join(d, "checkpoint"))
    assert T.latest_checkpoint(d) == p2                                  # falls back to the highest step present
