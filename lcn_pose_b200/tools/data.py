"""Device versions of the augmentation functions of the reference's tools/data.py (same names and argument meaning;
the arithmetic runs in liblcn_b200.so: lcn_augment / lcn_tta_undo / lcn_denormalize).

    flip_data(data)                       tools/data.py:10-25
    translation_data(data, t)             tools/data.py:27-54   (scalar factor; the ndarray branch of the reference is
                                                                  broken -- it adds the whole array -- and unused)
    rotate_data(data, angle=180)          tools/data.py:289-322
    undo(data, op_ord, number_actions, angle, translation)       tools/data.py:269-287

Inputs may be NumPy arrays (copied to the device, result returned as NumPy float64 like the reference) or CUDA
float32 tensors (result is a CUDA tensor: no host round trip)."""
import ctypes as C

import numpy as np
import torch

from .. import _lib as L


def _run_aug(data, op, angle=0.0, t=0.0):
    lib = L.load()
    is_np = not torch.is_tensor(data)
    src = torch.as_tensor(np.ascontiguousarray(data, dtype=np.float32)).cuda() if is_np else data.contiguous()
    assert src.is_cuda and src.dtype == torch.float32
    n = src.shape[0]
    k = src.numel() // (n * 17)
    dst = torch.empty_like(src)
    L.check(lib.lcn_augment(C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()), n, int(k), op, float(angle), float(t),
                            C.c_void_p(torch.cuda.current_stream(src.device).cuda_stream)))
    return dst.cpu().numpy().astype(np.float64).reshape(np.shape(data)) if is_np else dst


def flip_data(data):
    return _run_aug(data, L.LCN_AUG_FLIP)


def translation_data(data, translation_factor=0.5):
    if isinstance(translation_factor, np.ndarray):
        raise ValueError("per-item translation factors are not supported (the reference's ndarray branch is broken)")
    return _run_aug(data, L.LCN_AUG_TRANSLATE, t=translation_factor)


def rotate_data(data, angle=180):
    return _run_aug(data, L.LCN_AUG_ROTATE, angle=angle)


def undo(data, op_ord, number_actions=2, angle=180, translation=0.5):
    """data: [(number_actions+1)*N, 51] predictions of the original + augmented inputs, stacked; op_ord maps
    'f' / 'r' / 't' to the slice index of that operation.  Returns [N, 51]."""
    lib = L.load()
    is_np = not torch.is_tensor(data)
    src = torch.as_tensor(np.ascontiguousarray(data, dtype=np.float32)).cuda() if is_np else data.contiguous()
    total = src.numel() // 51
    n = total // (number_actions + 1)
    assert n * (number_actions + 1) == total
    out = torch.empty((n, 51), dtype=torch.float32, device=src.device)
    L.check(lib.lcn_tta_undo(C.c_void_p(src.data_ptr()), C.c_void_p(out.data_ptr()), n, int(number_actions),
                             int(op_ord.get("f", -1)), int(op_ord.get("r", -1)), int(op_ord.get("t", -1)), float(angle),
                             float(translation), C.c_void_p(torch.cuda.current_stream(src.device).cuda_stream)))
    return out.cpu().numpy().astype(np.float64) if is_np else out
