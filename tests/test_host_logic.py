"""CPU tests of the host-side logic that mirrors base_model.fit / predict bookkeeping (no GPU compute)."""
import numpy as np

from lcn_pose_b200.network.models_att import DebiasedEma, PermutationSampler, schedule
from lcn_pose_b200.tools import params_help, filter_hub
from oracle import lcn_oracle as O


def test_schedule_matches_reference_formulas():
    # models_att.py:185-186 with the Kaggle H36M run: 200 epochs x 7798 steps
    num_steps, eval_freq = schedule(200, 1559752, 200)
    assert num_steps == 1559752 and eval_freq == 7798
    assert schedule(3, 1000, 256) == (11, 3)


def test_permutation_sampler_covers_every_index_once_per_refill():
    rng = np.random.RandomState(0)
    s = PermutationSampler(10, 4, rng)
    seen = np.concatenate([s.next() for _ in range(5)])      # 20 indices = two permutations
    assert sorted(seen[:10].tolist()) == list(range(10))
    assert sorted(seen[10:].tolist()) == list(range(10))
    assert len(s.indices) == 0


def test_debiased_ema_equals_tf_zero_debias():
    e = DebiasedEma(0.9)
    xs = [3.0, 1.0, 2.0]
    out = [e.update(x) for x in xs]
    assert abs(out[0] - 3.0) < 1e-12
    b = 0.0
    for t, x in enumerate(xs, 1):
        b = 0.9 * b + 0.1 * x
    assert abs(out[-1] - b / (1 - 0.9 ** 3)) < 1e-12


def test_params_help_mirror():
    p = params_help.get_params(is_training=True)
    assert p["batch_size"] == 200 and p["dropout"] == 0.25 and p["F"] == 64 and p["in_F"] == 2
    assert p["decay_params"] == {"decay_steps": 32000, "decay_rate": 0.96}
    assert p["neighbour_matrix"].tobytes() == O.get_neighbour_matrix_by_hand(knn=1).tobytes()

    class A:
        test_indices = "7"; knn = 3; layers = 5; dropout = 0.1; channels = 128; checkpoints = "best"
        mask_type = "exponential"; init_type = "same"; epochs = 2; batch_size = 512
        learning_rate = 5e-4; regularization = None
    params_help.update_parameters(A, p)
    assert p["dir_name"] == "test7/" and p["num_layers"] == 5 and p["F"] == 128 and p["batch_size"] == 512
    assert p["neighbour_matrix"].tobytes() == O.get_neighbour_matrix_by_hand(knn=3).tobytes()
    assert p["regularization"] is None and p["in_F"] == 2          # --in-F never reaches params
    other = {k: list(v) for k, v in filter_hub.neighbour_dict_set[0].items()}
    other[3] = [2, 6]; other[6] = [5, 3]
    m = params_help.get_neighbour_matrix_by_hand(other, knn=2)
    ref = O.get_neighbour_matrix_by_hand(other, knn=2)
    assert m.tobytes() == ref.tobytes()
