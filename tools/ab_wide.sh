#!/bin/bash
for rep in 1 2; do for v in t5old t5new; do
  BGNN_LIB=$PWD/tools/libbgnn_$v.so timeout 300 python bench.py --workload wide --steps 5 --warmup 2 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('$v fp32', j['fp32']['ms_per_step'], 'bf16', j['bf16']['ms_per_step'], 'agg_l0', j['fp32']['kernels_ms']['agg_l0'], j['bf16']['kernels_ms']['agg_l0'], 'agg_l3', j['bf16']['kernels_ms']['agg_l3'])"
done; done
