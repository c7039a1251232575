#!/bin/bash
# Round 2, GPU call J: the pipelined ray-shard sink in bench.py, the register-resident render_prep_kernel, the SEG-sized cdf loops of
# coarse_to_fine_kernel and the pair-aware workspace sizing -- whole GPU suite (with durations), smoke, the default bench line, the
# launch list of the frame and ncu --set full of the three stage kernels of the bench frame.
mkdir -p gpurun_out
timeout 700 python -m pytest tests -q -m gpu --durations=12 > gpurun_out/r2j_tests.log 2>&1
(timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -4) > gpurun_out/r2j_smoke.log 2>&1
(timeout 420 python bench.py 2> gpurun_out/r2j_bench.err | tail -1) > gpurun_out/r2j_bench.json
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02j.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2j_l1.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"render_prep_kernel|coarse_to_fine_kernel|volume_weights_kernel" -s 9 -c 3 -f \
    -o gpurun_out/prof_r02j_stages python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2j_p1.log 2>&1
tail -15 gpurun_out/r2j_tests.log; cat gpurun_out/r2j_smoke.log; cut -c1-700 gpurun_out/r2j_bench.json; echo; tail -3 gpurun_out/r2j_bench.err
