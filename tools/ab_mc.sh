#!/bin/bash
# A/B of the 2-CTA cluster multicast (default on; BGNN_NO_MULTICAST=1 switches it off) for the N = 256 PM_SPLIT projections: parity tests first, then the bench
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_umma.py tests/test_gpu_gnn.py -q -x 2>&1 | tail -5 | cut -c1-300
rc=${PIPESTATUS[0]}
if [ "$rc" != "0" ]; then echo "multicast parity run rc=$rc (124 = hang)"; exit 0; fi
for v in base mc base mc; do
  if [ $v = base ]; then export BGNN_NO_MULTICAST=1; else unset BGNN_NO_MULTICAST; fi
  timeout 200 python bench.py --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); k=j['kernels']; print('$v', round(j['ms_per_step'],3), {n:k[n]['ms'] for n in ('proj_l0','proj_l1','proj_l2','proj_l3','agg_l1')})"
done
