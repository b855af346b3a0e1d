"""`from tools import tools` of evaluate.py -> lcn_pose_b200.tools.tools."""
from lcn_pose_b200.tools.tools import *  # noqa: F401,F403
from lcn_pose_b200.tools.tools import align_to_gt, evaluate_batch, image_to_camera_frame, procrustes  # noqa: F401
