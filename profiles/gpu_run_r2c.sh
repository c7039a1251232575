#!/bin/bash
# Round 2, GPU call C: fused coarse->fine + preparation kernels (bit-exactness vs the stage chain, whole-frame parity), the training
# step again, bench lines, launch lists and ncu captures of the new kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -s > gpurun_out/r2c_tests.log 2>&1
(timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -4) > gpurun_out/r2c_smoke.log 2>&1
(timeout 400 python bench.py 2> gpurun_out/r2c_bench.err | tail -1) > gpurun_out/r2c_bench.json
(timeout 300 python bench.py --workload train_step --no-extras 2> gpurun_out/r2c_bench_train.err | tail -1) > gpurun_out/r2c_bench_train.json
(timeout 300 python bench.py --workload train_step --precision bf16 --no-extras 2>> gpurun_out/r2c_bench_train.err | tail -1) > gpurun_out/r2c_bench_train_bf16.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_fused.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2c_l1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r02_train.csv \
    python bench.py --workload train_step --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2c_l2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"coarse_to_fine|render_prep|volume_weights" -s 9 -c 3 -f -o gpurun_out/prof_r02_stages \
    python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2c_p1.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/r2c_tests.log | tail -20
grep -E "training step|2048 rays" gpurun_out/r2c_tests.log | cut -c1-300
cat gpurun_out/r2c_smoke.log; cut -c1-700 gpurun_out/r2c_bench.json; echo; tail -3 gpurun_out/r2c_bench.err
cut -c1-2500 gpurun_out/r2c_bench_train.json; echo; cut -c1-600 gpurun_out/r2c_bench_train_bf16.json; tail -3 gpurun_out/r2c_bench_train.err
