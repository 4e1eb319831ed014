#!/bin/bash
# round-end validation on one B200: GPU tests, default bench (+ reference arm), smoke, secondary workloads
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/pytest_gpu_full.log
timeout 600 python bench.py 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/r1w_bench.json; cut -c1-300 gpurun_out/r1w_bench.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r1w_bench_reference.json; cut -c1-200 gpurun_out/r1w_bench_reference.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --workload wide --steps 5 --warmup 2 2>&1 | tail -1 > gpurun_out/r1w_bench_wide.json; cut -c1-200 gpurun_out/r1w_bench_wide.json
