#!/bin/bash
# Round 2, GPU call I: checkpoint after the per-family instantiation of the CTA-pair kernel -- whole GPU suite, smoke, the default bench
# line (modes, parity, extras), the head_torso line, the reference arm, launch lists and ncu --set full captures of the bench's launches.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2i_tests.log 2>&1
(timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -4) > gpurun_out/r2i_smoke.log 2>&1
(timeout 600 python bench.py 2> gpurun_out/r2i_bench.err | tail -1) > gpurun_out/r2i_bench.json
(timeout 300 python bench.py --workload head_torso 2> gpurun_out/r2i_bench_ht.err | tail -1) > gpurun_out/r2i_bench_ht.json
(timeout 200 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/r2i_ref.err | tail -1) > gpurun_out/r2i_ref.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2i_l1.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_ht.csv \
    python bench.py --workload head_torso --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2i_l2.log 2>&1
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r02_train.csv \
    python bench.py --workload train_step --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2i_l3.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:mlp_pair_kernel -s 6 -c 2 -f -o gpurun_out/prof_r02_bench_pair \
    python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2i_p1.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:mlp_pair_kernel -s 6 -c 2 -f -o gpurun_out/prof_r02_bench_dec \
    python bench.py --workload head_torso --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2i_p2.log 2>&1
tail -3 gpurun_out/r2i_tests.log; cat gpurun_out/r2i_smoke.log; cut -c1-400 gpurun_out/r2i_bench.json; echo; cut -c1-300 gpurun_out/r2i_bench_ht.json; echo; cut -c1-300 gpurun_out/r2i_ref.json
