#!/bin/bash
# usage: profiles/run_dp_n.sh N tag [modes...]   -- data-parallel bench at N GPUs for each exchange mode (bounded by timeouts)
N=$1; TAG=$2; shift 2
mkdir -p gpurun_out/$TAG
for mode in "$@"; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 200 --warmup 10 --dp-mode $mode --no-cpu-baseline --infer-poses 0 \
    > gpurun_out/$TAG/bench_n${N}_${mode}.json 2> gpurun_out/$TAG/bench_n${N}_${mode}.err
  echo "mode $mode rc=$?"; python -c "
import json,sys
try:
  d=json.load(open('gpurun_out/$TAG/bench_n${N}_${mode}.json')); print('$mode', d['ms_per_step'], d['value'], d['e2e']['value'])
except Exception as e: print('no json', e)"
done
