#!/bin/bash
# Round 2, GPU call D: CTA-pair kernel as the default of the single-pass precisions -- whole GPU suite, smoke, bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -s > gpurun_out/r2d_tests.log 2>&1
(timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -4) > gpurun_out/r2d_smoke.log 2>&1
(timeout 500 python bench.py 2> gpurun_out/r2d_bench.err | tail -1) > gpurun_out/r2d_bench.json
grep -E "passed|failed|FAILED|Error" gpurun_out/r2d_tests.log | tail -20
cat gpurun_out/r2d_smoke.log; cut -c1-1200 gpurun_out/r2d_bench.json; echo; tail -3 gpurun_out/r2d_bench.err
