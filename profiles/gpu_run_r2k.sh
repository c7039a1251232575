#!/bin/bash
# Round 2, GPU call K: early staging in the split schedule (mlp_pp.cu) -- A/B against the round-1 schedule (bit-identity + timing),
# the x3-related GPU tests, the bench frame in the two parity modes.
mkdir -p gpurun_out
timeout 200 python profiles/ab_pp_flags.py > gpurun_out/r2k_ab.log 2>&1
timeout 500 python -m pytest tests/test_gpu_2_mlp.py tests/test_gpu_3_render.py tests/test_gpu_5_decoder_tc.py tests/test_gpu_4_decoder.py -q -m gpu -x > gpurun_out/r2k_tests.log 2>&1
(timeout 200 python bench.py --precision fp16x3m --steps 5 --no-extras --no-cpu-baseline 2> gpurun_out/r2k_b1.err | tail -1) > gpurun_out/r2k_bench_x3m.json
(timeout 200 python bench.py --precision bf16x3 --steps 5 --no-extras --no-cpu-baseline 2> gpurun_out/r2k_b2.err | tail -1) > gpurun_out/r2k_bench_x3.json
cat gpurun_out/r2k_ab.log; tail -5 gpurun_out/r2k_tests.log; cut -c1-330 gpurun_out/r2k_bench_x3m.json; echo; cut -c1-330 gpurun_out/r2k_bench_x3.json
