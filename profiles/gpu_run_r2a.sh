#!/bin/bash
# Round 2, GPU call A: parity tests (incl. the new reduced-precision gates), the bench line with modes / parity / extras, the CPU
# reference arm, launch lists and ncu --set full captures of the bench's own tcgen05 launches.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -q -m gpu -s 2>&1 | tail -250) > gpurun_out/r2a_tests.log 2>&1
(timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -4) > gpurun_out/r2a_smoke.log 2>&1
(timeout 400 python bench.py 2> gpurun_out/r2a_bench.err | tail -1) > gpurun_out/r2a_bench.json
(timeout 200 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/r2a_ref.err | tail -1) > gpurun_out/r2a_ref.json
(timeout 200 python bench.py --workload head_torso --no-extras 2> gpurun_out/r2a_bench_ht.err | tail -1) > gpurun_out/r2a_bench_ht.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2a_l1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 6 -c 2 -f -o gpurun_out/prof_r02_bench_bf16 \
    python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2a_p1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_pp_kernel -s 6 -c 2 -f -o gpurun_out/prof_r02_bench_x3 \
    python bench.py --precision bf16x3 --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2a_p2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_pp_kernel -s 6 -c 2 -f -o gpurun_out/prof_r02_bench_dec \
    python bench.py --workload head_torso --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2a_p3.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:embed_kernel -s 3 -c 1 -f -o gpurun_out/prof_r02_embed \
    python profiles/prof_embed.py > gpurun_out/r2a_p4.log 2>&1
(timeout 100 python profiles/prof_embed.py 2>&1 | tail -2) > gpurun_out/r2a_embed.log
(timeout 200 python profiles/teacher_forced_precisions.py 2>&1 | tail -5) > gpurun_out/r2a_tf.log
tail -60 gpurun_out/r2a_tests.log; cat gpurun_out/r2a_smoke.log; cut -c1-3000 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err; cat gpurun_out/r2a_embed.log gpurun_out/r2a_tf.log
