"""`from network import models_att` of the reference's scripts -> lcn_pose_b200.network.models_att."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _root import root_path  # noqa: E402
from lcn_pose_b200.network import models_att as _m  # noqa: E402
from lcn_pose_b200.network.models_att import *  # noqa: F401,F403,E402
from lcn_pose_b200.network.models_att import base_model, cgcnn, get_exponential_matrix  # noqa: F401,E402

_m.ROOT_PATH = root_path()
