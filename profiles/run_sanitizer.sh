#!/bin/bash
# compute-sanitizer memcheck over the GPU parity tests (run on the B200 box: gpurun -- bash profiles/run_sanitizer.sh).
# Covers the HBM-bound stage kernels, the encoders, the fp32 building blocks and the tcgen05 kernels (FaceNeRF / NeRF
# programs in bf16 and bf16x3, Decoder head + torso programs, the fused render paths).
set -e
cd "$(dirname "$0")/.."
compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest \
    tests/test_gpu_1_stages.py tests/test_gpu_6_encoders.py tests/test_gpu_4_decoder.py tests/test_gpu_5_decoder_tc.py \
    "tests/test_gpu_2_mlp.py::test_query_points_bf16x3_vs_fp32_oracle" "tests/test_gpu_2_mlp.py::test_query_points_bf16_vs_bf16_oracle" \
    tests/test_gpu_3_render.py -x -q
