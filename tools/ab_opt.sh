#!/bin/bash
# A/B of a model option through its environment default (BGNN_NO_<NAME>=1 switches it off): bench --quick twice each way
# usage: tools/ab_opt.sh BGNN_NO_AGG_PIPELINE [tag]
VAR=${1:-BGNN_NO_AGG_PIPELINE}; TAG=${2:-ab}
mkdir -p gpurun_out
for rep in 1 2; do
  for v in on off; do
    if [ $v = off ]; then export $VAR=1; else unset $VAR; fi
    python bench.py --no-survey --no-extra --no-cpu --steps 20 --warmup 5 > gpurun_out/${TAG}_${v}_${rep}.json 2> gpurun_out/${TAG}_${v}_${rep}.err
    python - <<PY
import json
j=json.load(open("gpurun_out/${TAG}_${v}_${rep}.json"))
k=j["kernels"]
print("$v $rep ms_per_step %.3f e2e %.3f | " % (j["ms_per_step"], j["e2e"]["ms_per_step"]) + " ".join("%s=%.3f" % (n, k[n]["ms"]) for n in sorted(k) if n.startswith(("agg","proj","heads","graph_build_total"))))
PY
  done
done
