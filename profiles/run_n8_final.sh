#!/bin/bash
# round-2 final multi-GPU record: config 2 with the streamed and the end-of-backward exchange, config 4 streamed
bash profiles/run_dp_n.sh 8 n8 p2p p2p-end
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 \
  bench.py --gpus 8 --config 4 --steps 100 --warmup 10 --dp-mode p2p --no-cpu-baseline --infer-poses 0 \
  > gpurun_out/n8/bench_n8_config4.json 2> gpurun_out/n8/bench_n8_config4.err
echo "config4 rc=$?"; cat gpurun_out/n8/bench_n8_config4.json | head -c 600
