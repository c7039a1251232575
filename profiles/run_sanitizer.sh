#!/bin/bash
# compute-sanitizer memcheck over the GPU parity tests (run on the B200 box: gpurun -- bash profiles/run_sanitizer.sh).
# Covers the HBM-bound stage kernels (incl. the fused coarse->fine and preparation launches), the encoders, the fp32 building blocks,
# the tcgen05 kernels (FaceNeRF / NeRF programs on the CTA-pair kernel and on mlp_pp.cu, the Decoder head + torso programs on both,
# the mixed-precision mode), the fused render paths and the training step's kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest \
    tests/test_gpu_1_stages.py tests/test_gpu_6_encoders.py tests/test_gpu_4_decoder.py tests/test_gpu_5_decoder_tc.py \
    "tests/test_gpu_2_mlp.py::test_query_points_parity_modes_vs_fp32_oracle" "tests/test_gpu_2_mlp.py::test_query_points_single_pass_vs_quantized_oracle" \
    "tests/test_gpu_2_mlp.py::test_kernel_variants_agree" "tests/test_gpu_2_mlp.py::test_split_schedule_switches" \
    "tests/test_gpu_2_mlp.py::test_query_points_mixed_mode_at_scale_and_nerf" tests/test_gpu_3_render.py \
    "tests/test_gpu_8_train.py::test_gemm_against_fp64" "tests/test_gpu_8_train.py::test_two_training_steps_match_the_reference" \
    -q 2>&1 | tail -25
echo "sanitizer rc=${PIPESTATUS[0]}"
