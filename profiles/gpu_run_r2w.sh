#!/bin/bash
# Round 2, GPU call W: per-ray bias rows staged in shared memory for the pair kernel's view layer (flag 4): bit-identity tests, A/B
# of the 20,000-ray query (flags 3 vs 7 through dfn_debug_set_impl), the bench frame.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_2_mlp.py tests/test_gpu_3_render.py -q -m gpu > gpurun_out/r2w_tests.log 2>&1
timeout 200 python profiles/ab_impl.py 20000 1,67,131,67,131 > gpurun_out/r2w_ab.log 2>&1
(timeout 200 python bench.py --steps 10 --no-extras --no-cpu-baseline 2> gpurun_out/r2w_b1.err | tail -1) > gpurun_out/r2w_bench.json
tail -4 gpurun_out/r2w_tests.log; grep -v Warn gpurun_out/r2w_ab.log | tail -12; cut -c1-330 gpurun_out/r2w_bench.json
