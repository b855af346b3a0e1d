"""Per-kernel in-stream durations of one data-parallel train step (eager launches, LCN_TRACE=1 -> CUDA events around every
lcn_launch), to see what the gradient exchange costs kernel by kernel.
    LCN_TRACE=1 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 profiles/trace_dp_step.py [p2p|packed]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from lcn_pose_b200 import dist as lcn_dist  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
mode = sys.argv[1] if len(sys.argv) > 1 else "p2p"
eng = bench.make_engine(2, "bf16", local)
if mode == "p2p":
    lcn_dist.init_native_dp(eng)
x, y = bench.synth_xy(4096, seed=1234 + rank)
xd, yd = torch.as_tensor(x).to(dev), torch.as_tensor(y).to(dev)
for _ in range(5):
    lcn_dist.dp_train_step(eng, xd, yd, 0.25, mode=mode, graph=False)
torch.cuda.synchronize()
dist.barrier()
eng.lib.lcn_debug_trace_dump()          # drop the warm-up records
for _ in range(20):
    lcn_dist.dp_train_step(eng, xd, yd, 0.25, mode=mode, graph=False)
torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    sys.stdout.flush()
    eng.lib.lcn_debug_trace_dump()
eng.close()
dist.barrier()
dist.destroy_process_group()
