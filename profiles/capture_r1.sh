#!/bin/bash
# Round-1 evidence run (one GPU): tests, bench, launch list of the bench command, ncu --set full of the dominant kernels.
# Everything lands in gpurun_out/r1/ ; summaries are then copied to profiles/r1/ by hand (profiles/ncu_summary.py).
set -x
mkdir -p gpurun_out/r1
python -m pytest tests -m gpu -q > gpurun_out/r1/tests_gpu.log 2>&1; echo "tests rc=$?"
python bench.py > gpurun_out/r1/bench.json 2> gpurun_out/r1/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r1/bench_reference.json 2> gpurun_out/r1/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --infer-poses 0 > gpurun_out/r1/ncu_launches.log 2>&1
# the launch bench.py's roofline block times (mid-layer forward GEMM, L2 flushed before it)
ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 3 -c 2 -o gpurun_out/r1/full_tc_gemm_mid -f \
    python profiles/run_gemm_once.py > gpurun_out/r1/ncu_full_gemm_mid.log 2>&1
# the same kernel and the weight-gradient kernel inside a train step (eager launches)
ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 20 -c 2 -o gpurun_out/r1/full_tc_gemm -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --infer-poses 0 > gpurun_out/r1/ncu_full_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tc_wgrad -s 10 -c 1 -o gpurun_out/r1/full_tc_wgrad -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --infer-poses 0 > gpurun_out/r1/ncu_full_wgrad.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lcn_stack -s 2 -c 1 -o gpurun_out/r1/full_lcn_stack -f \
    python profiles/run_fused_once.py > gpurun_out/r1/ncu_full_stack.log 2>&1
ncu --set full --clock-control none -k regex:k_eval -c 2 -o gpurun_out/r1/full_eval -f \
    python profiles/probe_infer.py 4194304 > gpurun_out/r1/ncu_full_eval.log 2>&1
python profiles/probe_infer.py 4194304 > gpurun_out/r1/probe_infer.log 2>&1
ls -la gpurun_out/r1
