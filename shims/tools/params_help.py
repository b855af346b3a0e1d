"""`from tools import params_help` -> lcn_pose_b200.tools.params_help."""
from lcn_pose_b200.tools.params_help import *  # noqa: F401,F403
from lcn_pose_b200.tools.params_help import get_neighbour_matrix_by_hand, get_params, update_parameters  # noqa: F401
