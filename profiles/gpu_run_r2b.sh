#!/bin/bash
# Round 2, GPU call B: the training-step kernels (strided tcgen05 GEMM, loss backward, Adam), the whole step against the golden
# vectors, the recalibrated reduced-precision gates with their full printout, and the train_step bench line.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_8_train.py -q -s > gpurun_out/r2b_train_tests.log 2>&1
timeout 600 python -m pytest tests -q -m gpu -s --deselect tests/test_gpu_8_train.py > gpurun_out/r2b_tests.log 2>&1
(timeout 300 python bench.py --workload train_step --no-extras 2> gpurun_out/r2b_bench_train.err | tail -1) > gpurun_out/r2b_bench_train.json
(timeout 300 python bench.py --workload train_step --precision bf16 --no-extras 2>> gpurun_out/r2b_bench_train.err | tail -1) > gpurun_out/r2b_bench_train_bf16.json
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_8_train.py -q -k "gemm and bf16x3 or loss_backward or colsum" > gpurun_out/r2b_sanitizer.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/r2b_sanitizer.log
grep -E "gemm case|loss_bwd|training step|2048 rays|passed|failed|Error|error" gpurun_out/r2b_train_tests.log | cut -c1-260 | tail -80
grep -E "restatement|operand program|passed|failed" gpurun_out/r2b_tests.log | grep -v "print(" | cut -c1-330
tail -5 gpurun_out/r2b_sanitizer.log; cut -c1-1500 gpurun_out/r2b_bench_train.json; tail -3 gpurun_out/r2b_bench_train.err
