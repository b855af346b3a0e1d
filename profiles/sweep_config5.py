"""BASELINE.json configs[4]: wide/deep LCN sweep -- layers=5, 128 channels, knn in {1,2,3,full} at batch 16384:
train-step and mid-layer-GEMM throughput vs mask density (fraction of the bf16 tensor peak on the nonzero blocks)."""
import ctypes as C, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcn_pose_b200.engine import LcnEngine
from lcn_pose_b200 import _lib as L
B, F, LAYERS = int(os.environ.get("SWEEP_B", 16384)), 128, 5
peak = 1652.9e12
rng = np.random.default_rng(0)
x = torch.as_tensor((rng.random((B, 34)) - 0.5).astype(np.float32)).cuda()
y = torch.as_tensor(rng.normal(0, 0.1, (B, 51)).astype(np.float32)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for knn in (1, 2, 3, 9):
    nm = L.neighbour_matrix(knn)
    nnz = int((nm != 0).sum())
    eng = LcnEngine(F=F, in_F=2, num_layers=LAYERS, neighbour_matrix=nm, path="bf16")
    eng.init_params(42)
    for _ in range(3):
        eng.train_step(x, y, dropout=0.25)
    torch.cuda.synchronize()
    reps = 10
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush.zero_()
        a.record(); eng.train_step(x, y, dropout=0.25); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)[reps // 2]
    fwd_flop = 2 * nnz * (2 * F + 2 * LAYERS * F * F + 3 * F)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for a, b in ev:
        flush.zero_()
        a.record()
        L.check(eng.lib.lcn_layer_gemm(eng.h, eng.params.data_ptr(), eng.ws.data_ptr(), eng.ws.numel(), B, B, 2, 0, st))
        b.record()
    torch.cuda.synchronize()
    gms = sorted(a.elapsed_time(b) for a, b in ev)[reps // 2]
    gflop = 2.0 * nnz * F * F * B
    print(json.dumps({"knn": knn, "nnz_blocks_of_289": nnz, "batch": B, "train_ms": ms, "train_poses_per_s": B / ms * 1e3,
                      "train_tensor_frac_burst": B / ms * 1e3 * 3 * fwd_flop / peak, "mid_gemm_ms": gms,
                      "mid_gemm_tflops": gflop / gms / 1e9, "mid_gemm_tensor_frac_burst": gflop / gms / 1e9 / 1652.9}), flush=True)
    del eng
    torch.cuda.empty_cache()
