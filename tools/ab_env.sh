#!/bin/bash
# A/B one environment switch of libbgnn.so inside ONE gpurun call (same GPU): tools/ab_env.sh VAR [reps]
for rep in $(seq 1 ${2:-2}); do for v in "" 1; do
  env ${v:+$1=1} timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); k=j['kernels']; print('$1=$v', round(j['ms_per_step'],3), {n:round(v['ms'],3) for n,v in k.items() if not n.startswith('gb_')})"
done; done
