"""`from tools import filter_hub` -> lcn_pose_b200.tools.filter_hub."""
from lcn_pose_b200.tools.filter_hub import *  # noqa: F401,F403
from lcn_pose_b200.tools.filter_hub import neighbour_dict_set  # noqa: F401
