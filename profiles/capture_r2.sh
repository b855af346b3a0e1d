#!/bin/bash
# Round-2 evidence run (one GPU): bench lines of every config, launch list of the bench command, ncu --set full of the
# dominant kernels (bf16 and split-bf16 GEMM, weight gradient, fused inference).  Everything lands in gpurun_out/r2/ ;
# summaries are then copied to profiles/r2/ (profiles/ncu_summary.py, profiles/summarize_launches.py).
set -x
mkdir -p gpurun_out/r2
python bench.py > gpurun_out/r2/bench.json 2> gpurun_out/r2/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/r2/bench_reference.json 2> gpurun_out/r2/bench_reference.err
python bench.py --path x3 --steps 30 --warmup 5 --no-cpu-baseline --infer-poses 0 > gpurun_out/r2/bench_x3.json 2> gpurun_out/r2/bench_x3.err
python bench.py --config 3 --steps 2 --total-poses 67108864 > gpurun_out/r2/bench_config3.json 2> gpurun_out/r2/bench_config3.err; echo "config3 rc=$?"
python bench.py --config 4 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2/bench_config4.json 2> gpurun_out/r2/bench_config4.err; echo "config4 rc=$?"
python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/r2/bench_config5.json 2> gpurun_out/r2/bench_config5.err; echo "config5 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --infer-poses 0 > gpurun_out/r2/ncu_launches.log 2>&1
# the launch bench.py's roofline block times (mid-layer forward GEMM, L2 flushed before it)
ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 3 -c 2 -o gpurun_out/r2/full_tc_gemm_mid -f \
    python profiles/run_gemm_once.py > gpurun_out/r2/ncu_full_gemm_mid.log 2>&1
LCN_BENCH_PATH=x3 ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 3 -c 2 -o gpurun_out/r2/full_tc_gemm_mid_x3 -f \
    python profiles/run_gemm_once.py > gpurun_out/r2/ncu_full_gemm_mid_x3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tc_wgrad -s 10 -c 1 -o gpurun_out/r2/full_tc_wgrad -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --infer-poses 0 > gpurun_out/r2/ncu_full_wgrad.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lcn_stack -s 2 -c 1 -o gpurun_out/r2/full_lcn_stack -f \
    python profiles/run_fused_once.py > gpurun_out/r2/ncu_full_stack.log 2>&1
ls -la gpurun_out/r2
