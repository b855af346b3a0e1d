"""lcn_pose_b200 -- B200 (sm_100a) implementation of the adgx/lcn-pose LCN hot path.

Python host side: a ctypes binding of the C ABI in include/lcn_b200.h (`_lib`), the engine that owns
device buffers through torch (`engine`), and mirrors of the reference modules the hot path lives behind
(`network.models_att`, `tools.tools`, `tools.params_help`, `tools.filter_hub`, `tools.data`).
There is no CPU fallback: importing `_lib` fails loudly when liblcn_b200.so is missing.
"""
__version__ = "0.1.0"
