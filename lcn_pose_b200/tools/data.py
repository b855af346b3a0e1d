"""Device versions of the augmentation functions of the reference's tools/data.py (same names and argument meaning;
the arithmetic runs in liblcn_b200.so: lcn_augment / lcn_tta_undo / lcn_denormalize).

    flip_data(data)                       tools/data.py:10-25
    translation_data(data, t)             tools/data.py:27-54   (scalar factor; the ndarray branch of the reference is
                                                                  broken -- it adds the whole array -- and unused)
    rotate_data(data, angle=180)          tools/data.py:289-322
    undo(data, op_ord, number_actions, angle, translation)       tools/data.py:269-287

DataReader (tools/data.py:324-489) keeps the reference's methods -- real_read, read_2d, read_3d, denormalize -- with
the per-item Python loops replaced by one kernel each (lcn_normalize / lcn_denormalize); `*_device` variants return
CUDA tensors so that predict -> denormalize -> evaluate never leaves the GPU.

Inputs may be NumPy arrays (copied to the device, result returned as NumPy float64 like the reference) or CUDA
float32 tensors (result is a CUDA tensor: no host round trip)."""
import ctypes as C
import os
import pickle

import numpy as np
import torch

from .. import _lib as L

ROOT_PATH = os.path.join(os.path.dirname(os.path.realpath(__file__)), "..", "..")


def camera_resolution(camera_name):
    """The per-camera image size of DataReader (tools/data.py:358-369, repeated at :374-385, :407-417, :460-469)."""
    camera_name = str(camera_name)
    if camera_name in ("54138969", "60457274"):
        return 1000, 1002
    if camera_name in ("55011271", "58860488", "50591643", "65906101"):
        return 1000, 1000
    if camera_name.find("cam_") != -1:
        return 2048, 2048
    if int(camera_name) >= 0:
        return 2048, 2048
    assert 0, "data item has an invalid camera name %s" % camera_name


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


class DataReader(object):
    """tools/data.py:324-489.  Same methods and return types (float64 NumPy, list of dicts); the arithmetic runs on
    the device."""

    def __init__(self):
        self.gt_trainset = None
        self.gt_testset = None
        self.dt_dataset = None

    def real_read(self, subset, type):
        file_name = "%s_%s.pkl" % (subset, type)
        print("loading %s" % file_name)
        with open(os.path.join(ROOT_PATH, "dataset", file_name), "rb") as f:
            return pickle.load(f)

    @staticmethod
    def _stack(items):
        """joint_3d_image [n,17,3] float32 and (res_w, res_h) [n,2] of a list of dataitems, on the device."""
        j = np.empty((len(items), 17, 3), dtype=np.float32)
        r = np.empty((len(items), 2), dtype=np.float32)
        for idx, item in enumerate(items):
            j[idx] = item["joint_3d_image"]
            r[idx] = camera_resolution(item["camera_param"]["name"])
        return torch.as_tensor(j).cuda(), torch.as_tensor(r).cuda()

    @staticmethod
    def normalize_device(joint_3d_image, res, want_2d=True, want_3d=True):
        """read_2d / read_3d 'scale' normalisation (tools/data.py:355-371,404-424) of device tensors:
        -> (x2d [n,34] or None, y3d [n,51] or None)."""
        n = joint_3d_image.shape[0]
        x2d = torch.empty((n, 34), dtype=torch.float32, device=joint_3d_image.device) if want_2d else None
        y3d = torch.empty((n, 51), dtype=torch.float32, device=joint_3d_image.device) if want_3d else None
        if n:
            L.check(L.load().lcn_normalize(C.c_void_p(joint_3d_image.data_ptr()), C.c_void_p(res.data_ptr()), n,
                                           C.c_void_p(x2d.data_ptr() if want_2d else 0),
                                           C.c_void_p(y3d.data_ptr() if want_3d else 0), _stream(joint_3d_image)))
        return x2d, y3d

    def read_2d(self, gt_trainset, gt_testset, which="scale"):
        if self.gt_trainset is None:
            self.gt_trainset = gt_trainset
        if self.gt_testset is None:
            self.gt_testset = gt_testset
        if which != "scale":
            assert 0, "not support normalize type %s" % which
        out = []
        for items in (self.gt_trainset, self.gt_testset):
            j, r = self._stack(items)
            out.append(self.normalize_device(j, r, True, False)[0].cpu().numpy().astype(np.float64))
        return out[0], out[1]

    def read_3d(self, which="scale"):
        if self.gt_trainset is None:
            self.gt_trainset = self.real_read("train")        # (the reference's call, missing an argument, :396)
        if self.gt_testset is None:
            self.gt_testset = self.real_read("test")
        if which != "scale":
            assert 0, "not support normalize type %s" % which
        out = []
        for items in (self.gt_trainset, self.gt_testset):
            j, r = self._stack(items)
            out.append(self.normalize_device(j, r, False, True)[1].cpu().numpy().astype(np.float64))
        return out[0], out[1]

    def denormalize_device(self, data):
        """[n,51] or [n,17,3] CUDA float32 (normalised predictions) -> [n,17,3] image-frame poses, on the device, with the
        resolutions of self.gt_testset (tools/data.py:471-472)."""
        pose = data.reshape(-1, 17, 3).clone().contiguous()
        r = np.empty((len(self.gt_testset), 2), dtype=np.float32)
        for idx, item in enumerate(self.gt_testset):
            r[idx] = camera_resolution(item["camera_param"]["name"])
        n = min(pose.shape[0], len(self.gt_testset))
        res = torch.as_tensor(r[:n]).to(pose.device)
        if n:
            L.check(L.load().lcn_denormalize(C.c_void_p(pose.data_ptr()), C.c_void_p(res.data_ptr()), n, _stream(pose)))
        return pose

    def denormalize(self, data, which="scale"):
        """tools/data.py:447-489: list of {'cameraid', 'videoid', 'subject', 'action', 'result': [17,3] float64}."""
        if self.gt_testset is None:
            self.gt_testset = self.real_read("test")
        if which != "scale":
            assert 0
        d = torch.as_tensor(np.ascontiguousarray(np.asarray(data).reshape(-1, 17, 3), dtype=np.float32)).cuda()
        pose = self.denormalize_device(d).cpu().numpy().astype(np.float64)
        res = []
        for idx, item in enumerate(self.gt_testset):
            res.append({"cameraid": item["cameraid"], "videoid": item["videoid"], "subject": item["subject"],
                        "action": item["action"], "result": pose[idx]})
        return res


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
                             float(translation) if translation is not None else 0.0,     # inference.py passes None without --translate_data
                             C.c_void_p(torch.cuda.current_stream(src.device).cuda_stream)))
    return out.cpu().numpy().astype(np.float64) if is_np else out


# ---- dataset subsetting (tools/data.py:56-201): host list bookkeeping used by train.py / inference.py --subset ----
_SUBSET_KEYS = {"camera": ("cameraid",), "action": ("action",), "subject": ("subject",),
                "camera_action": ("cameraid", "action"), "camera_subject": ("cameraid", "subject"),
                "action_camera_subject": ("action", "cameraid", "subject")}


def _first_per_group(gt_dataset, key, subset_size):
    """The first subset_size // n_groups items of every distinct value of item[key], grouped in order of first
    appearance of the group in a set() walk like the reference (its uniform(0, 1) draw is truthy with probability 1)."""
    if not all(key in item for item in gt_dataset):
        raise ValueError("The dataset must contain '%s' key in each item." % key)
    groups = {name: [] for name in set(item[key] for item in gt_dataset)}
    quota = subset_size // len(groups)
    for item in gt_dataset:
        np.random.uniform(0, 1)                      # keeps the global NumPy stream in step with the reference
        if len(groups[item[key]]) < quota:
            groups[item[key]].append(item)
    return [it for g in groups.values() for it in g]


def get_subset_by_camera(gt_dataset, subset_size=1000):
    return _first_per_group(gt_dataset, "cameraid", subset_size)


def get_subset_by_action(gt_dataset, subset_size=1000):
    return _first_per_group(gt_dataset, "action", subset_size)


def get_subset_by_subject(gt_dataset, subset_size=1000):
    return _first_per_group(gt_dataset, "subject", subset_size)


def get_subset(gt_dataset, subset_size=1000, mode="camera"):
    """tools/data.py:142-201.  Multi-key modes intersect the single-key subsets by item identity (the reference
    builds set()s of dicts, which raises TypeError for its list-of-dict datasets; identity is what it means)."""
    if subset_size is None:
        return gt_dataset
    keys = _SUBSET_KEYS[mode]
    subsets = [_first_per_group(gt_dataset, k, subset_size) for k in keys]
    if len(subsets) == 1:
        return subsets[0]
    common = set(map(id, subsets[0]))
    for s in subsets[1:]:
        common &= set(map(id, s))
    combined = [it for it in subsets[0] if id(it) in common]
    if len(combined) < subset_size:
        return combined
    pick = np.random.choice(len(combined), size=subset_size, replace=False)
    return [combined[i] for i in pick]
