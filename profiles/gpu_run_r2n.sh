#!/bin/bash
# Round 2, GPU call N: alpha_linear in the last trunk layer's epilogue (fp16x3m, TC_F_DOT_ALPHA) -- A/B, x3 tests, precision table, bench frame.
mkdir -p gpurun_out
timeout 300 python profiles/ab_pp_flags.py > gpurun_out/r2n_ab.log 2>&1
timeout 400 python -m pytest tests/test_gpu_2_mlp.py tests/test_gpu_3_render.py -q -m gpu > gpurun_out/r2n_tests.log 2>&1
timeout 200 python profiles/teacher_forced_precisions.py > gpurun_out/r2n_tf.log 2>&1
(timeout 200 python bench.py --precision fp16x3m --steps 5 --no-extras --no-cpu-baseline 2> gpurun_out/r2n_b1.err | tail -1) > gpurun_out/r2n_bench_x3m.json
cat gpurun_out/r2n_ab.log; tail -8 gpurun_out/r2n_tests.log; tail -12 gpurun_out/r2n_tf.log; cut -c1-330 gpurun_out/r2n_bench_x3m.json
