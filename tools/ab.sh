#!/bin/bash
# A/B builds of libbgnn.so inside ONE gpurun call (same GPU): tools/ab.sh reps tag1 tag2 ...   (tools/libbgnn_<tag>.so)
reps=$1; shift
for rep in $(seq 1 $reps); do for v in "$@"; do
  BGNN_LIB=$PWD/tools/libbgnn_$v.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-survey --no-extra 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); k=j['kernels']; print('$v', round(j['ms_per_step'],3), {n:round(v['ms'],3) for n,v in k.items() if not n.startswith('gb_')})"
done; done
