"""`from tools import data` of train.py / inference.py -> lcn_pose_b200.tools.data."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _root import root_path  # noqa: E402
from lcn_pose_b200.tools import data as _d  # noqa: E402
from lcn_pose_b200.tools.data import *  # noqa: F401,F403,E402
from lcn_pose_b200.tools.data import (DataReader, flip_data, get_subset, rotate_data, translation_data,  # noqa: F401,E402
                                      undo)

_d.ROOT_PATH = root_path()
