#!/bin/bash
# sweep of the fused-kernel configurations (env overrides of stack_config) at BN group 256 and 128
for cfg in "LCN_STACK_NV=1" "LCN_STACK_NV=1 LCN_STACK_NS1=5" "LCN_STACK_NV=2 LCN_STACK_MC=1" "LCN_STACK_NV=2 LCN_STACK_MC=0" "LCN_STACK_NV=1 LCN_STACK_MC=0"; do
  for bn in 256 128; do
    echo "== $cfg bn=$bn"
    env $cfg timeout 120 python profiles/probe_infer.py 1048576 $bn 2>&1 | head -1
  done
done
