#!/bin/bash
# Round 2, GPU call H: after the CTA-pair kernel was instantiated per program family (FaceNeRF/NeRF | Decoder head | Decoder torso) --
# parity tests, the default bench line and the launch list of the frame.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_3_render.py tests/test_gpu_2_mlp.py tests/test_gpu_1_stages.py tests/test_gpu_4_decoder.py tests/test_gpu_5_decoder_tc.py -q -m gpu > gpurun_out/r2h_tests.log 2>&1
(timeout 300 python bench.py --no-extras 2> gpurun_out/r2h_bench.err | tail -1) > gpurun_out/r2h_bench.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r02_h.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2h_l.log 2>&1
tail -3 gpurun_out/r2h_tests.log; cut -c1-700 gpurun_out/r2h_bench.json
