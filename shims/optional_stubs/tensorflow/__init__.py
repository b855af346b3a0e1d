"""Placeholder so that `import tensorflow as tf` at the top of the reference's inference.py succeeds on a box without
TensorFlow.  The script never touches `tf`; the model it builds comes from shims/network/models_att.py.  Any attribute
access fails loudly: nothing computes through this module."""


def __getattr__(name):
    raise AttributeError("tensorflow is not installed: shims/optional_stubs/tensorflow is an import placeholder only "
                         "(attribute %r requested)" % name)
