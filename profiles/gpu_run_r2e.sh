#!/bin/bash
# Round 2, GPU call E: ncu evidence for the CTA-pair kernel and the fused stage kernels (launch lists + --set full captures of the
# bench's own launches), and the other workloads' captures after the epilogue change.
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2e_l1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_ht.csv \
    python bench.py --workload head_torso --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2e_l2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_pair_kernel -s 6 -c 2 -f -o gpurun_out/prof_r02_bench_pair \
    python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2e_p1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"coarse_to_fine|render_prep|volume_weights" -s 9 -c 3 -f -o gpurun_out/prof_r02_stages \
    python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2e_p2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_pp_kernel -s 6 -c 2 -f -o gpurun_out/prof_r02_bench_dec \
    python bench.py --workload head_torso --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2e_p3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_pp_kernel -s 6 -c 2 -f -o gpurun_out/prof_r02_bench_x3 \
    python bench.py --precision bf16x3 --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2e_p4.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r02*.csv | tail -12
