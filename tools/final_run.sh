mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/pytest_gpu_full.log
timeout 600 python bench.py 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/r1w_bench.json; cut -c1-600 gpurun_out/r1w_bench.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r1w_bench_reference.json; cut -c1-400 gpurun_out/r1w_bench_reference.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1w_launches_wide.csv python bench.py --workload wide --steps 1 --warmup 1 > gpurun_out/ncu_list_wide.log 2>&1; tail -1 gpurun_out/ncu_list_wide.log | cut -c1-200
timeout 700 ncu --set full --clock-control none --import-source on -k regex:"gat_aggregate_tile5|proj_bf16" -s 12 -c 4 -o gpurun_out/prof_r1w_wide -f python bench.py --workload wide --steps 1 --warmup 1 > gpurun_out/ncu_wide.log 2>&1; tail -1 gpurun_out/ncu_wide.log | cut -c1-200
