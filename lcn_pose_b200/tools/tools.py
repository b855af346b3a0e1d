"""Mirror of the metric functions of the reference's tools/tools.py, backed by the batched CUDA evaluator.

evaluate.py calls these per pose (evaluate.py:53-59); the per-pose signatures are kept for drop-in use and
delegate to the same kernel with n = 1.  `evaluate_batch` is the entry a user should call: it runs the whole
evaluate._eval loop (un-projection, optional Procrustes, per-joint error, per-joint / per-action sums, PCK) in
one launch.  No CPU fallback: everything goes through lcn_eval_mpjpe.
"""
import numpy as np
import torch

from ..engine import eval_mpjpe

THRESHOLD = 50  # mm, evaluate.py:10


def _dev(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).cuda()


def evaluate_batch(preds, gts, boxes, cams, root_depths, protocol2=False, actions=None, n_actions=0,
                   want_err=True):
    """preds [n,17,3] image frame, gts [n,17,3] camera frame (mm), boxes [n,4], cams [n,4]=(fx,fy,cx,cy),
    root_depths [n]; numpy or CUDA tensors.  Returns dict(err [n,17] float32 numpy or None,
    per_joint [17] MPJPE, mpjpe, pck, per_action [n_actions] or None)."""
    t = [x if torch.is_tensor(x) else _dev(x) for x in (preds, gts, boxes, cams, root_depths)]
    act = None
    if actions is not None:
        act = actions if torch.is_tensor(actions) else _dev(actions, torch.int32)
    err, sums = eval_mpjpe(t[0].reshape(-1, 17, 3), t[1].reshape(-1, 17, 3), t[2], t[3], t[4], protocol2, act,
                           n_actions, want_err)
    s = sums.cpu().numpy()
    allr = s[-1]
    out = {"err": err.cpu().numpy() if err is not None else None,
           "per_joint": allr[:17] / allr[17], "mpjpe": float((allr[:17] / allr[17]).mean()),
           "pck": float(allr[18] / (allr[17] * 17) * 100), "per_action": None, "per_action_pck": None}
    if act is not None and n_actions > 0:
        cnt = np.maximum(s[:n_actions, 17], 1)
        out["per_action"] = s[:n_actions, :17].sum(1) / (cnt * 17)
        out["per_action_pck"] = s[:n_actions, 18] / (cnt * 17) * 100
    return out


def image_to_camera_frame(pose3d_image_frame, box, camera, rootIdx, root_depth):
    """tools/tools.py:183-194 for one pose [17,3]; evaluated by the batched kernel with n = 1 (float32 on the
    device, returned as float64 like the reference).  Use evaluate_batch for throughput."""
    p = np.asarray(pose3d_image_frame, np.float32).reshape(1, 17, 3)
    cam = np.array([[camera["fx"], camera["fy"], camera["cx"], camera["cy"]]], dtype=np.float32)
    _, _, pose = eval_mpjpe(_dev(p), _dev(np.zeros_like(p)), _dev(np.asarray(box, np.float32)[None]), _dev(cam),
                            _dev(np.asarray([root_depth], np.float32)), False, want_err=False, want_pose=True)
    return pose[0].cpu().numpy().astype(np.float64)


def align_to_gt(pose, pose_gt):
    """tools/tools.py:197-202: procrustes(pose_gt, pose)[1] (scaling, reflections allowed) for one pose."""
    p = np.asarray(pose, np.float32).reshape(1, 17, 3)
    g = np.asarray(pose_gt, np.float32).reshape(1, 17, 3)
    _, _, z = eval_mpjpe(_dev(p), _dev(g), None, None, None, True, want_err=False, want_pose=True, camera_frame=True)
    return z[0].cpu().numpy().astype(np.float64)


def procrustes(A, B, scaling=True, reflection='best'):
    """tools/tools.py:96-181 for one pose pair (A = target [17,3], B = input [17,3]): returns (d, Z, tform) with
    tform = {'rotation', 'scale', 'translation'} exactly like the reference, every option of its signature
    (lcn_procrustes; float32 on the device, float64 out).  procrustes_batch takes [n,17,3] stacks."""
    d, Z, tf = procrustes_batch(np.asarray(A)[None], np.asarray(B)[None], scaling, reflection)
    return float(d[0]), Z[0], {"rotation": tf["rotation"][0], "scale": float(tf["scale"][0]),
                               "translation": tf["translation"][0]}


def procrustes_batch(A, B, scaling=True, reflection='best'):
    """n pose pairs at once: A, B [n,17,3] (NumPy or CUDA float32).  Returns (d [n], Z [n,17,3], tform dict of
    stacked 'rotation' [n,3,3], 'scale' [n], 'translation' [n,3]) as float64 NumPy."""
    import ctypes as C
    from .. import _lib as L
    a = A if torch.is_tensor(A) else _dev(np.asarray(A, np.float32))
    b = B if torch.is_tensor(B) else _dev(np.asarray(B, np.float32))
    a, b = a.reshape(-1, 17, 3).contiguous(), b.reshape(-1, 17, 3).contiguous()
    assert a.shape == b.shape
    n = a.shape[0]
    z = torch.empty_like(a)
    tf = torch.empty((n, 14), dtype=torch.float32, device=a.device)
    refl = 0 if isinstance(reflection, str) else (2 if reflection else 1)
    L.check(L.load().lcn_procrustes(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), n, int(bool(scaling)), refl,
                                    C.c_void_p(z.data_ptr()), C.c_void_p(tf.data_ptr()),
                                    C.c_void_p(torch.cuda.current_stream(a.device).cuda_stream)))
    t = tf.cpu().numpy().astype(np.float64)
    return t[:, 13], z.cpu().numpy().astype(np.float64), {"rotation": t[:, :9].reshape(n, 3, 3), "scale": t[:, 9],
                                                           "translation": t[:, 10:13]}


def pose_errors(pred_image_frame, gt, box, camera, root_depth, protocol2=False):
    """One evaluate.py:54-61 iteration on the device (n = 1): returns the 17 per-joint errors in mm."""
    cam = np.array([[camera["fx"], camera["fy"], camera["cx"], camera["cy"]]], dtype=np.float32)
    r = evaluate_batch(np.asarray(pred_image_frame, np.float32)[None], np.asarray(gt, np.float32)[None],
                       np.asarray(box, np.float32)[None], cam, np.asarray([root_depth], np.float32), protocol2)
    return r["err"][0]
