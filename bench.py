#!/usr/bin/env python
"""bench.py -- poses/s of the LCN train step (BASELINE.json configs[1]: knn=3, layers=3, F=64,
locally_connected mask, batch 4096 per GPU, masked TF1 Adam) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (restated)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

KNN, LAYERS, F, BATCH = 3, 3, 64, 4096
NNZ = 175                                     # nonzero joint-pair blocks of the knn=3 mask
FWD_FLOP_PER_POSE = 2 * NNZ * (2 * F + 2 * LAYERS * F * F + 3 * F)     # 8 713 600
TRAIN_FLOP_PER_POSE = 3 * FWD_FLOP_PER_POSE
WORKLOAD = f"LCN knn={KNN} layers={LAYERS} F={F} locally_connected, train step (fwd+bwd+masked Adam), batch {BATCH}/GPU"


def synth_xy(n, seed=1234):
    rng = np.random.default_rng(seed)
    root = rng.uniform(-0.5, 0.5, (n, 1, 2))
    x = np.clip(root + rng.normal(0, 0.15, (n, 17, 2)), -1, 1)
    y = np.concatenate([x + rng.normal(0, 0.02, (n, 17, 2)), rng.normal(0, 0.1, (n, 17, 1))], axis=2)
    return x.reshape(n, 34).astype(np.float32), y.reshape(n, 51).astype(np.float32)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture (profiles/roofline_traffic.json names the report); None when absent."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    return d.get("dram_bytes_read", 0) + d.get("dram_bytes_write", 0)


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled during the timed region: NVML (about 1 ms per sample) when pynvml is
    importable, else nvidia-smi (about 100 ms per sample)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def run(self):
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.index)
            mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
            bits = [(N.nvmlClocksEventReasonHwSlowdown if hasattr(N, "nvmlClocksEventReasonHwSlowdown") else N.nvmlClocksThrottleReasonHwSlowdown),
                    (N.nvmlClocksEventReasonHwThermalSlowdown if hasattr(N, "nvmlClocksEventReasonHwThermalSlowdown") else N.nvmlClocksThrottleReasonHwThermalSlowdown),
                    (N.nvmlClocksEventReasonSwThermalSlowdown if hasattr(N, "nvmlClocksEventReasonSwThermalSlowdown") else N.nvmlClocksThrottleReasonSwThermalSlowdown),
                    (N.nvmlClocksEventReasonSwPowerCap if hasattr(N, "nvmlClocksEventReasonSwPowerCap") else N.nvmlClocksThrottleReasonSwPowerCap)]
            get = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop_flag:
                sm = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
                r = get(h)
                self.rows.append([str(sm), str(mx)] + ["Active" if (r & b) else "Not Active" for b in bits])
                time.sleep(0.002)
            return
        except Exception:
            pass
        self.run_smi()

    def run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference executes this path as dense fp32 matmuls on masked weights (TF stock ops,
# network/models_att.py).  TensorFlow is not installable here, so the restated graph
# (oracle/torch_restatement.py, autograd) + the oracle's TF1 Adam is timed on all host cores.
# --------------------------------------------------------------------------------------------------
def cpu_train_steps(batch, steps, warmup):
    import torch
    from oracle import lcn_oracle as O
    from oracle import torch_restatement as T
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.LcnConfig(F=F, num_layers=LAYERS, neighbour_matrix=O.get_neighbour_matrix_by_hand(knn=KNN))
    p = T.build_params(O.init_params(cfg, seed=42, dtype=np.float32), dtype=torch.float32)
    x, y = synth_xy(batch)
    xt, yt = torch.tensor(x), torch.tensor(y)
    m = {k: torch.zeros_like(v) for k, v in p.items()}
    v2 = {k: torch.zeros_like(v) for k, v in p.items()}
    t = 0

    def step():
        nonlocal t
        t += 1
        for v in p.values():
            v.grad = None
        loss, _ = T.loss_fn(cfg, p, xt, yt)
        loss.backward()
        lr_t = O.learning_rate_at(cfg, t) * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        with torch.no_grad():
            for k, w in p.items():
                m[k].mul_(0.9).add_(w.grad, alpha=0.1)
                v2[k].mul_(0.999).addcmul_(w.grad, w.grad, value=0.001)
                w.sub_(lr_t * m[k] / (v2[k].sqrt() + 1e-8))
        return float(loss.detach())
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return dt, torch.get_num_threads()


def run_reference(args, rank, world):
    if rank != 0:
        return
    batch = 1024          # bounded sample of the batch-4096 step: same graph, 1/4 of the rows
    dt, threads = cpu_train_steps(batch, args.steps, args.warmup)
    value = batch * args.steps / dt
    line = {"impl": "reference", "metric": "poses/sec (LCN train step)", "value": value, "unit": "poses/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "l2": "n/a (CPU)"},
            "cpu_baseline": {"value": value, "unit": "poses/s", "cores": threads, "kind": "port",
                             "sample": f"{args.steps} train steps at batch {batch} (dense fp32 restatement of the "
                                       f"TF graph, torch-CPU autograd; TensorFlow 2.13 not installable offline)"},
            "e2e": {"value": value, "unit": "poses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=JSON_OUT, flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def launches_per_step(n_bn):
    """Kernels of this library per train step on the bf16 path (counted from the launch list, profiles/r1):
    forward: (GEMM | first layer) + bn_act per BN layer (the BatchNorm statistics are accumulated by the GEMM itself),
    head GEMM; backward: loss, head dgrad, last-layer wgrad, BN backward (reduce + apply) per BN layer, dgrad + wgrad per mid layer,
    first-layer wgrad, bias-gradient reduce; optimizer: pairdot, mask gradient, Adam, mask scalars, 5 weight packs."""
    fwd = 2 * n_bn + 1
    bwd = 3 + 2 * n_bn + 2 * (n_bn - 1) + 2
    opt = 9
    return fwd + bwd + opt


def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from lcn_pose_b200.engine import LcnEngine
    from oracle import lcn_oracle as O   # only for the neighbour matrix of the workload and the CPU baseline leg
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = LcnEngine(F=F, in_F=2, num_layers=LAYERS, mask_type="locally_connected",
                    neighbour_matrix=O.get_neighbour_matrix_by_hand(knn=KNN), path=args.path, device=f"cuda:{local_rank}")
    eng.init_params(seed=42)
    x, y = synth_xy(BATCH, seed=1234 + rank)
    xd, yd = torch.as_tensor(x).to(dev), torch.as_tensor(y).to(dev)
    x_pin, y_pin = torch.as_tensor(x).pin_memory(), torch.as_tensor(y).pin_memory()
    xe, ye = torch.empty_like(xd), torch.empty_like(yd)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    n_bn = 1 + 2 * LAYERS

    # one bucket: the raw gradients, packed to the nonzero joint-pair blocks (LCN_DP_PACKED=0: the full parameter-layout vector)
    from lcn_pose_b200.dist import average_gradient_bucket as allreduce
    packed = os.environ.get("LCN_DP_PACKED", "1") != "0"

    def step(xx, yy):
        if args.no_graph:
            eng.forward(xx, bn_group=BATCH, training=True, dropout=args.dropout)
            eng.backward(xx, yy, args.dropout)
            if world > 1:
                if packed:
                    allreduce(eng.pack_grads())
                    eng.unpack_grads()
                else:
                    allreduce(eng.grads_raw)
            eng.adam()
        else:
            # the same launches, replayed from a CUDA graph (LcnEngine.train_step_graph)
            eng.train_step_graph(xx, yy, args.dropout, allreduce if world > 1 else None, packed=packed)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(xd, yd)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    # ---- device-resident timing: K steps, L2 flushed between steps, CUDA events per step ----
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in evs:
        flush.zero_()
        a.record()
        step(xd, yd)
        b.record()
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    # ---- end to end through the public API: every step copies ITS inputs from pinned host memory to the device,
    # runs the step and reads the loss back into pinned host memory.  The input copy of step i+1 runs on a copy
    # stream into the other of two device buffers while step i computes (what an input pipeline does); the host
    # waits for the GPU once, after the last loss has landed. ----
    main_s = torch.cuda.current_stream()
    copy_s = torch.cuda.Stream(device=dev)
    bufs = [(xe, ye), (torch.empty_like(xd), torch.empty_like(yd))]
    losses_pin = torch.zeros(args.steps).pin_memory()
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    for b in range(2):                               # graph capture / warm-up of both buffer sets, untimed
        bufs[b][0].copy_(x_pin); bufs[b][1].copy_(y_pin)
        step(*bufs[b])
        consumed[b].record(main_s)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        b = i & 1
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(consumed[b])           # the step that last read this buffer pair has finished
            bufs[b][0].copy_(x_pin, non_blocking=True)
            bufs[b][1].copy_(y_pin, non_blocking=True)
            copied[b].record(copy_s)
        main_s.wait_event(copied[b])
        step(*bufs[b])
        consumed[b].record(main_s)
        losses_pin[i:i + 1].copy_(eng.loss_dev, non_blocking=True)
    barrier()
    e2e_s = time.perf_counter() - t0
    assert bool(torch.isfinite(losses_pin).all()), "non-finite loss in the end-to-end loop"
    # ---- dominant kernel alone: forward mid-layer GEMM (block-sparse X*(W.M)) ----
    import ctypes as C
    from lcn_pose_b200 import _lib as L
    gemm_ms = None
    if rank == 0:
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        reps = 20
        for _ in range(3):
            L.check(eng.lib.lcn_layer_gemm(eng.h, eng.params.data_ptr(), eng.ws.data_ptr(), eng.ws.numel(), BATCH, BATCH, 2, 0, st))
        ge = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in ge:
            flush.zero_()
            a.record()
            L.check(eng.lib.lcn_layer_gemm(eng.h, eng.params.data_ptr(), eng.ws.data_ptr(), eng.ws.numel(), BATCH, BATCH, 2, 0, st))
            b.record()
        torch.cuda.synchronize()
        gemm_ms = sum(a.elapsed_time(b) for a, b in ge) / reps
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    # ---- BASELINE.json configs[2] (secondary numbers, not the headline): inference at BN group 256 + Protocol-1/2
    # evaluation, pose batch sharded over the ranks with no communication on the data path ----
    inf = None
    if args.infer_poses > 0:
        from lcn_pose_b200.engine import eval_mpjpe
        n_inf = (args.infer_poses // 256) * 256
        eng2 = LcnEngine(F=F, in_F=2, num_layers=LAYERS, mask_type="locally_connected",
                         neighbour_matrix=O.get_neighbour_matrix_by_hand(knn=KNN), path=args.path, device=f"cuda:{local_rank}")
        eng2.init_params(seed=42)
        gen = torch.Generator(device=dev).manual_seed(1234 + rank)
        xi = torch.rand((n_inf, 34), device=dev, generator=gen) - 0.5
        oi = torch.empty((n_inf, 51), device=dev)
        gt = torch.randn((n_inf, 17, 3), device=dev, generator=gen) * 300 + torch.tensor([0., 0., 4500.], device=dev)
        box = torch.tensor([0., 0., 999., 999.], device=dev).repeat(n_inf, 1)
        cam = torch.tensor([1145.05, 1143.78, 512.54, 515.45], device=dev).repeat(n_inf, 1)
        rd = gt[:, 0, 2].contiguous()

        def best_ms(fn, reps=3):
            fn()
            barrier()
            evs2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
            for a, b in evs2:
                a.record(); fn(); b.record()
            barrier()
            return min(a.elapsed_time(b) for a, b in evs2)
        f_ms = best_ms(lambda: eng2.forward(xi, bn_group=256, out=oi))
        pred = oi.view(n_inf, 17, 3)
        p1_ms = best_ms(lambda: eval_mpjpe(pred, gt, box, cam, rd, False, want_err=False))
        p2_ms = best_ms(lambda: eval_mpjpe(pred, gt, box, cam, rd, True, want_err=False))
        inf = [f_ms, p1_ms, p2_ms]
    # ---- max over ranks ----
    t = torch.tensor([dev_ms, e2e_s] + (inf or [0, 0, 0]), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s, f_ms, p1_ms, p2_ms = t.tolist()
    if rank == 0:
        tf_burst, tf_sust, hbm, how = measured_peaks()
        value = world * BATCH * args.steps / (dev_ms * 1e-3)
        gemm_flop = 2.0 * NNZ * 64 * 64 * BATCH
        ach = gemm_flop / (gemm_ms * 1e-3) / 1e12
        line = {"metric": "poses/sec (LCN train step)", "value": value, "unit": "poses/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if args.path == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "l2": "flushed between steps (256 MiB memset)",
                           "dropout": args.dropout, "path": args.path, "launch": "eager" if args.no_graph else "cuda-graph replay",
                           "parallelism": f"dp{world}: per-GPU BatchNorm statistics, NCCL allreduce of the " + ("packed nonzero-block" if packed else "parameter-layout") + " gradient bucket"},
                "e2e": {"value": world * BATCH * args.steps / e2e_s, "unit": "poses/s",
                        "h2d_bytes_per_step": int(x_pin.numel() * 4 + y_pin.numel() * 4), "d2h_bytes_per_step": 4},
                "gpu_launches": launches_per_step(n_bn) * args.steps,
                "step_tensor_frac": value / world * TRAIN_FLOP_PER_POSE / (tf_sust * 1e12),
                "roofline": {"bound": "tensor", "kernel": "mid-layer forward GEMM (block-sparse, 175 nonzero 64x64 blocks)",
                             "achieved": ach, "peak": tf_burst, "unit": "TFLOP/s", "frac": ach / tf_burst,
                             "traffic": ncu_traffic(), "peak_source": how, "ms_per_launch": gemm_ms},
                "clocks": sampler.summary() if sampler else None}
        if inf is not None:
            tot = world * n_inf
            line["inference"] = {
                "workload": f"configs[2] shape: LCN knn={KNN} layers={LAYERS} F={F} inference at BN group 256 + Protocol-1/2 "
                            f"evaluation, {n_inf} synthetic poses per GPU resident in HBM, sharded by BN group, no collective",
                "forward_poses_per_s": tot / (f_ms * 1e-3),
                "forward_tensor_frac_burst": n_inf / (f_ms * 1e-3) * FWD_FLOP_PER_POSE / (tf_burst * 1e12),
                "eval_p1_poses_per_s": tot / (p1_ms * 1e-3), "eval_p1_hbm_frac": n_inf * 444 / (p1_ms * 1e-3) / (hbm * 1e9),
                "eval_p2_poses_per_s": tot / (p2_ms * 1e-3), "eval_p2_hbm_frac": n_inf * 444 / (p2_ms * 1e-3) / (hbm * 1e9)}
        # CPU baseline: bounded sample on this box's host cores (rank 0, N=1 only)
        if world == 1 and not args.no_cpu_baseline:
            cb, cs = 1024, 8
            dt, threads = cpu_train_steps(cb, cs, 1)
            line["cpu_baseline"] = {"value": cb * cs / dt, "unit": "poses/s", "cores": threads, "kind": "port",
                                    "sample": f"{cs} train steps at batch {cb}: dense fp32 restatement of the TF graph "
                                              f"(torch-CPU autograd + TF1 Adam); TensorFlow 2.13 not installable offline"}
        print(json.dumps(line), file=JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


JSON_OUT = sys.stdout


def main():
    # The contract is ONE JSON line on stdout.  NCCL prints its version banner with a C-level printf to stdout when the
    # first communicator is created (seen on the GPU boxes: "NCCL version 2.28.9+cuda12.9" in front of the line), so the
    # process's fd 1 is pointed at stderr and the JSON line goes to a duplicate of the original stdout.
    global JSON_OUT
    sys.stdout.flush()
    JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--path", default=os.environ.get("LCN_BENCH_PATH", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--dropout", type=float, default=0.25)     # params_help.py:166 training default
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--infer-poses", type=int, default=1 << 22, help="poses per GPU of the secondary inference+eval leg (0: skip)")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
