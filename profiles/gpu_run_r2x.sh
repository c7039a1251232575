#!/bin/bash
# Round 2, GPU call X: coalesced loaders / pair-store epilogue of the training GEMM -- the training tests, the three GEMM shapes in
# isolation, the train_step bench line; and the ray-sharded sequence loop with two frames of host slack (N = 1 part: tests only).
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_8_train.py -q -m gpu > gpurun_out/r2x_tests.log 2>&1
timeout 200 python profiles/bench_gemm.py > gpurun_out/r2x_gemm.log 2>&1
(timeout 200 python bench.py --workload train_step --steps 20 --no-extras --no-cpu-baseline 2> gpurun_out/r2x_b1.err | tail -1) > gpurun_out/r2x_bench_train.json
tail -4 gpurun_out/r2x_tests.log; grep -v Warn gpurun_out/r2x_gemm.log | tail -12; cut -c1-330 gpurun_out/r2x_bench_train.json; python -c "
import json; d=json.loads(open('gpurun_out/r2x_bench_train.json').read()); print(d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['roofline']['avg_launch_ms'])"
