"""Mirror of the reference's tools/params_help.py API (the constructor-kwargs contract of cgcnn).

get_neighbour_matrix_by_hand (params_help.py:8-20) is bit exact; the default skeleton goes through the
C ABI (lcn_neighbour_matrix), any other dictionary through the same integer arithmetic in NumPy.
"""
import numpy as np

from . import filter_hub
from .. import _lib


def get_neighbour_matrix_by_hand(neighbour_dict, knn=1):
    assert len(neighbour_dict) == 17
    if neighbour_dict == filter_hub.neighbour_dict_set[0]:
        return _lib.neighbour_matrix(max(int(knn), 1))
    a = np.zeros((17, 17), dtype=np.float32)
    for idx in range(17):
        a[idx, [idx] + list(neighbour_dict[idx])] = 1
    if knn >= 2:
        a = np.array(np.linalg.matrix_power(a, knn) != 0, dtype=np.float32)
    return a


def get_params(is_training, gt_dataset=None):
    """Default parameter dictionary, same keys and values as params_help.py:122-173."""
    return {
        "dir_name": "test1/", "num_epochs": 200, "batch_size": 200, "decay_type": "exp",
        "decay_params": {"decay_steps": 32000, "decay_rate": 0.96}, "F": 64,
        "mask_type": "locally_connected", "init_type": "random",
        "neighbour_matrix": get_neighbour_matrix_by_hand(filter_hub.neighbour_dict_set[0], knn=1),
        "in_joints": 17, "out_joints": 17, "num_layers": 3, "in_F": 2, "residual": True, "max_norm": True,
        "batch_norm": True, "regularization": 0, "dropout": 0.25 if is_training else 0, "learning_rate": 1e-3,
        "checkpoints": "final", "is_training": is_training, "knn": 1,
    }


def update_parameters(args, params):
    """CLI overrides, params_help.py:92-119 (note: --in-F is parsed by the scripts but never copied, 9-Q9)."""
    if getattr(args, "test_indices", None):
        params["dir_name"] = "test" + args.test_indices + "/"
    if getattr(args, "knn", None):
        params["knn"] = args.knn
        params["neighbour_matrix"] = get_neighbour_matrix_by_hand(filter_hub.neighbour_dict_set[0], knn=args.knn)
    if getattr(args, "layers", None) is not None:
        params["num_layers"] = args.layers
    if getattr(args, "dropout", None) is not None:
        params["dropout"] = args.dropout
    if getattr(args, "channels", None):
        params["F"] = args.channels
    if getattr(args, "checkpoints", None):
        params["checkpoints"] = args.checkpoints
    if getattr(args, "mask_type", None):
        params["mask_type"] = args.mask_type
    if getattr(args, "init_type", None):
        params["init_type"] = args.init_type
    if getattr(args, "epochs", None):
        params["num_epochs"] = args.epochs
    if getattr(args, "batch_size", None):
        params["batch_size"] = args.batch_size
    if hasattr(args, "learning_rate"):
        params["learning_rate"] = args.learning_rate
    if hasattr(args, "regularization"):
        params["regularization"] = args.regularization
