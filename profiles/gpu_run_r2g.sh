#!/bin/bash
# Round 2, GPU call G: after the Decoder programs moved to the CTA-pair kernel -- whole GPU suite, head_torso bench, launch list and
# ncu --set full captures of its launches (single-pass and split precision).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2g_tests.log 2>&1
(timeout 300 python bench.py --workload head_torso --no-extras 2> gpurun_out/r2g_bench_ht.err | tail -1) > gpurun_out/r2g_bench_ht.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r02_ht.csv \
    python bench.py --workload head_torso --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2g_l2.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:mlp_pair_kernel -s 6 -c 2 -f -o gpurun_out/prof_r02_bench_dec \
    python bench.py --workload head_torso --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2g_p3.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:mlp_pp_kernel -s 6 -c 2 -f -o gpurun_out/prof_r02_bench_dec_x3 \
    python bench.py --workload head_torso --precision bf16x3 --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2g_p4.log 2>&1
tail -3 gpurun_out/r2g_tests.log; cut -c1-500 gpurun_out/r2g_bench_ht.json
