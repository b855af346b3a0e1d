"""End-to-end parity of the whole path (BASELINE.json north_star: "end-to-end MPJPE and P-MPJPE within 0.1 mm"):
inference (predict batching, models_att.py:79-132) -> denormalize (tools/data.py:471-472) -> image_to_camera_frame ->
[Procrustes] -> per-joint error (evaluate.py:53-61), GPU kernels through the C ABI against the float64 oracle on the
same seeded H36M-shaped inputs and the same parameters.  BOTH arithmetic paths hold the north_star's 0.1 mm on the MPJPE
and P-MPJPE means and on every per-joint mean (measured on the B200, profiles/r2/parity_mpjpe.json: bf16 0.001 mm on the
means, 0.07 mm worst per-joint mean; the fp32-parity path 2e-6 mm)."""
import numpy as np
import pytest
import torch

from lcn_pose_b200 import _lib as L
from lcn_pose_b200.engine import eval_mpjpe
from oracle import lcn_oracle as O
from tests.gpu_helpers import dev, make_pair

pytestmark = pytest.mark.gpu

FX, FY, CX, CY, RES_W, RES_H = 1145.05, 1143.78, 512.54, 515.45, 1000.0, 1002.0


def _synthetic_set(n, seed=1234):
    """camera 54138969-like intrinsics, roots 3-6 m away, gt = root + N(0, 300) mm, box from the root
    +-1000 mm projected, 2D inputs = projected gt normalised to [-1, 1] (tools/data.py:355-371)."""
    rng = np.random.default_rng(seed)
    root = np.stack([rng.normal(0, 500, n), rng.normal(0, 500, n), rng.uniform(3000, 6000, n)], 1)
    gt = root[:, None, :] + rng.normal(0, 300, (n, 17, 3))
    gt[:, 0] = root
    tl = np.stack([(root[:, 0] - 1000) / root[:, 2] * FX + CX, (root[:, 1] - 1000) / root[:, 2] * FY + CY], 1)
    br = np.stack([(root[:, 0] + 1000) / root[:, 2] * FX + CX, (root[:, 1] + 1000) / root[:, 2] * FY + CY], 1)
    box = np.concatenate([tl, br], 1)
    u = gt[:, :, 0] / gt[:, :, 2] * FX + CX
    v = gt[:, :, 1] / gt[:, :, 2] * FY + CY
    x2d = np.stack([u / RES_W * 2 - 1, v / RES_W * 2 - RES_H / RES_W], 2).reshape(n, 34)
    cam = np.tile(np.array([FX, FY, CX, CY]), (n, 1))
    return x2d.astype(np.float32), gt, box, cam, root[:, 2].copy()


@pytest.mark.parametrize("path,tol_mm", [("fp32", 0.1), ("bf16", 0.1)])
def test_mpjpe_and_pmpjpe_end_to_end(path, tol_mm):
    n, bs = 1000, 256                       # four batches, the last one zero padded
    x2d, gt, box, cam, rd = _synthetic_set(n)
    eng, cfg, p = make_pair(L=2, knn=3, path=path)
    # ---- oracle, float64 ----
    pred_ref = O.predict(cfg, p, x2d.astype(np.float64), bs).reshape(n, 17, 3)
    den_ref = O.denormalize(pred_ref, np.full(n, RES_W), np.full(n, RES_H))
    e1_ref = O.eval_errors(den_ref, gt, box, cam, rd, False)
    e2_ref = O.eval_errors(den_ref, gt, box, cam, rd, True)
    # ---- GPU ----
    out = eng.forward(dev(x2d), bn_group=bs, training=False)
    pose = out.view(n, 17, 3).contiguous()
    res = dev(np.tile(np.array([RES_W, RES_H], np.float32), (n, 1)))
    L.check(L.load().lcn_denormalize(pose.data_ptr(), res.data_ptr(), n, torch.cuda.current_stream().cuda_stream))
    f32 = lambda a: dev(np.asarray(a, dtype=np.float32))
    e1, _ = eval_mpjpe(pose, f32(gt), f32(box), f32(cam), f32(rd), False)
    e2, _ = eval_mpjpe(pose, f32(gt), f32(box), f32(cam), f32(rd), True)
    e1, e2 = e1.cpu().numpy().astype(np.float64), e2.cpu().numpy().astype(np.float64)
    # what was measured goes on file (gpurun_out/parity_mpjpe.json -> profiles/r2/), not only pass / fail
    import json
    import os
    rec_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_mpjpe.json")
    os.makedirs(os.path.dirname(rec_path), exist_ok=True)
    blob = json.load(open(rec_path)) if os.path.exists(rec_path) else {}
    blob[path] = {"tolerance_mm": tol_mm, "poses": n, "mpjpe_oracle_mm": float(e1_ref.mean()), "pmpjpe_oracle_mm": float(e2_ref.mean()),
                  "mpjpe_delta_mm": float(abs(e1.mean() - e1_ref.mean())), "pmpjpe_delta_mm": float(abs(e2.mean() - e2_ref.mean())),
                  "per_joint_mpjpe_delta_max_mm": float(np.abs(e1.mean(0) - e1_ref.mean(0)).max()),
                  "per_joint_pmpjpe_delta_max_mm": float(np.abs(e2.mean(0) - e2_ref.mean(0)).max()),
                  "per_pose_joint_error_delta_max_mm": float(np.abs(e1 - e1_ref).max())}
    json.dump(blob, open(rec_path, "w"), indent=1, sort_keys=True)
    print("MPJPE parity", path, blob[path])
    assert abs(e1.mean() - e1_ref.mean()) < tol_mm, (path, e1.mean(), e1_ref.mean())
    assert abs(e2.mean() - e2_ref.mean()) < tol_mm, (path, e2.mean(), e2_ref.mean())
    # per-joint means (the table evaluate.py:102-106 prints)
    assert np.abs(e1.mean(0) - e1_ref.mean(0)).max() < tol_mm
    assert np.abs(e2.mean(0) - e2_ref.mean(0)).max() < tol_mm
