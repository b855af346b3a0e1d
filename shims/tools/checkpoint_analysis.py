"""tools/checkpoint_analysis.py of the reference prints the tensors of a checkpoint with TensorFlow's inspector;
this one lists them with the TensorBundle reader of lcn_pose_b200.tools.tf_checkpoint.

    python -m tools.checkpoint_analysis experiment/test1/checkpoints/final
"""
import sys

from lcn_pose_b200.tools import tf_checkpoint


def main(directory):
    prefix = tf_checkpoint.latest_checkpoint(directory) or directory
    for name, (dtype, shape) in sorted(tf_checkpoint.list_bundle(prefix).items()):
        print("tensor_name: ", name, " dtype:", {1: "float32", 3: "int32", 9: "int64"}.get(dtype, dtype), " shape:", shape)


if __name__ == "__main__":
    main(sys.argv[1])
