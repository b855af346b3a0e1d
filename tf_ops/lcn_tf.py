"""Python glue of the TensorFlow binding (tf_ops/lcn_ops.cc): loads the op library, registers the gradient of LcnLoss,
and shows the three graph-construction methods of the reference's network/models_att.py rewritten on the ops.

Needs TensorFlow (the reference pins 2.13, requirements.txt:1).  It is NOT importable in this repository's image -- the
tests and bench.py drive the same C ABI through ctypes (lcn_pose_b200/_lib.py) instead.  Kept next to the C++ source so
that a maintainer of the reference has the complete binding in one place (INTEGRATION.md walks through it).
"""
import os

import numpy as np
import tensorflow as tf

_HERE = os.path.dirname(os.path.abspath(__file__))
ops = tf.load_op_library(os.path.join(_HERE, "liblcn_tf_ops.so"))


@tf.RegisterGradient("LcnLoss")
def _lcn_loss_grad(op, d_loss, d_logits, d_grads):
    """d loss / d params is the op's own third output (the kernels fuse loss and backward, lcn_model_backward +
    lcn_model_finalize_grads); the data / labels / dropout / step inputs get no gradient, as in the reference where they
    are placeholders."""
    return [d_loss * op.outputs[2], None, None, None, None]


def model_attrs(model):
    """The attribute set every Lcn* op takes, from a cgcnn instance (constructor kwargs of models_att.py:478-506)."""
    L = np.asarray(model.neighbour_matrix, np.float32).T
    exp = "exponential" in model.mask_type
    return dict(F=model.F, in_F=model.in_F, num_layers=model.num_layers, mask_type=model.mask_type,
                residual=bool(model.residual), max_norm=bool(model.max_norm), path="bf16",
                support=(np.ones(289) if exp else (L != 0).astype(np.float32).reshape(-1)).tolist(),
                const_mask=model.exponential_matrix.reshape(-1).tolist() if exp else [])


# ---- what changes in network/models_att.py -------------------------------------------------------------------------
# build_graph (:288-333) keeps its placeholders; the per-tensor tf.compat.v1.get_variable calls of _inference_lcn /
# two_linear / _initialize_mask are replaced by ONE flat variable (initial value = the same initialisers concatenated in
# lcn_model_tensor_info order), and:
#
#   def _inference_lcn(self, x, data_dropout):                                    # replaces :707-775
#       return ops.lcn_forward(self.params, x, data_dropout, self.global_step + 1, **model_attrs(self))
#
#   def loss(self, logits, labels):                                                # replaces :352-380
#       loss, _, _ = ops.lcn_loss(self.params, self.ph_data, labels, self.ph_dropout, self.global_step + 1,
#                                 **model_attrs(self))
#       ...EMA / summaries unchanged...
#
#   def training(self, loss, learning_rate, decay_type, decay_params):             # replaces :382-421
#       either keep tf.compat.v1.train.AdamOptimizer: compute_gradients(loss) now returns [(grads, self.params)]
#       through the registered gradient above, or use the fused step:
#       lr = tf.compat.v1.train.exponential_decay(learning_rate, global_step, **decay_params)
#       return ops.lcn_adam(self.params.handle, self.adam_m.handle, self.adam_v.handle, self.ph_data, self.ph_labels,
#                           self.ph_dropout, global_step + 1, lr, regularization=self.regularization or 0.0,
#                           **model_attrs(self))
#
# evaluate.py:53-61 becomes one call: err, sums = ops.lcn_eval(pred, gt, box, cam, root_depth, protocol2=...).
