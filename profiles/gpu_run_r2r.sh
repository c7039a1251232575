#!/bin/bash
# Round 2, GPU call R (final N=1 evidence): whole GPU suite, smoke, default bench line (modes / parity / extras), head_torso line,
# reference arm, launch lists, ncu --set full of the pair kernel and of the fp16x3m / bf16x3 / Decoder-bf16x3 split kernels in the
# bench frames, timelines of the split kernel with the round-1 schedule and the present one.
mkdir -p gpurun_out
timeout 700 python -m pytest tests -q -m gpu > gpurun_out/r2r_tests.log 2>&1
(timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -4) > gpurun_out/r2r_smoke.log 2>&1
(timeout 200 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/r2r_ref.err | tail -1) > gpurun_out/r2r_ref.json
(timeout 500 python bench.py 2> gpurun_out/r2r_bench.err | tail -1) > gpurun_out/r2r_bench.json
(timeout 300 python bench.py --workload head_torso 2> gpurun_out/r2r_bench_ht.err | tail -1) > gpurun_out/r2r_bench_ht.json
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02b.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2r_l1.log 2>&1
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02b_ht.csv \
    python bench.py --workload head_torso --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2r_l2.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:mlp_pair_kernel -s 6 -c 2 -f -o gpurun_out/prof_r02b_bench_pair \
    python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2r_p1.log 2>&1
timeout 250 ncu --set full --clock-control none --import-source on -k regex:mlp_pp_kernel -s 6 -c 2 -f -o gpurun_out/prof_r02b_bench_x3m \
    python bench.py --precision fp16x3m --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2r_p2.log 2>&1
timeout 250 ncu --set full --clock-control none --import-source on -k regex:mlp_pp_kernel -s 6 -c 2 -f -o gpurun_out/prof_r02b_bench_x3 \
    python bench.py --precision bf16x3 --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2r_p3.log 2>&1
timeout 250 ncu --set full --clock-control none --import-source on -k regex:mlp_pp_kernel -s 6 -c 2 -f -o gpurun_out/prof_r02b_bench_dec_x3 \
    python bench.py --workload head_torso --precision bf16x3 --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2r_p4.log 2>&1
for f in 0 7; do for m in fp16x3m bf16x3; do DFN_PP_FLAGS=$f timeout 120 python profiles/trace_query.py $m > gpurun_out/trace_pp_${m}_f$f.txt 2>&1; done; done
tail -3 gpurun_out/r2r_tests.log; cat gpurun_out/r2r_smoke.log; cut -c1-400 gpurun_out/r2r_bench.json; echo; cut -c1-300 gpurun_out/r2r_bench_ht.json; echo; cut -c1-300 gpurun_out/r2r_ref.json
