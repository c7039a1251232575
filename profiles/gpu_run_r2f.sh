#!/bin/bash
# Round 2, GPU call F: checkpoint -- whole GPU suite, smoke, the default bench line (modes incl. fp16x3m, parity, extras), reference arm.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2f_tests.log 2>&1
(timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -4) > gpurun_out/r2f_smoke.log 2>&1
(timeout 600 python bench.py 2> gpurun_out/r2f_bench.err | tail -1) > gpurun_out/r2f_bench.json
(timeout 200 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/r2f_ref.err | tail -1) > gpurun_out/r2f_ref.json
tail -3 gpurun_out/r2f_tests.log; cat gpurun_out/r2f_smoke.log; cut -c1-600 gpurun_out/r2f_bench.json; echo; tail -2 gpurun_out/r2f_bench.err; cut -c1-400 gpurun_out/r2f_ref.json
