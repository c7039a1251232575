#!/bin/bash
# Round 2, GPU call Y (final N = 1 check of the committed tree): whole GPU suite, smoke, default bench line, head_torso line, reference arm.
mkdir -p gpurun_out
timeout 700 python -m pytest tests -q -m gpu > gpurun_out/r2y_tests.log 2>&1
(timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -4) > gpurun_out/r2y_smoke.log 2>&1
(timeout 500 python bench.py 2> gpurun_out/r2y_bench.err | tail -1) > gpurun_out/r2y_bench.json
(timeout 300 python bench.py --workload head_torso 2> gpurun_out/r2y_bench_ht.err | tail -1) > gpurun_out/r2y_bench_ht.json
(timeout 200 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/r2y_ref.err | tail -1) > gpurun_out/r2y_ref.json
tail -3 gpurun_out/r2y_tests.log; cat gpurun_out/r2y_smoke.log; cut -c1-300 gpurun_out/r2y_bench.json; echo; cut -c1-300 gpurun_out/r2y_bench_ht.json; echo; cut -c1-200 gpurun_out/r2y_ref.json
